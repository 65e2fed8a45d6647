import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from tests.helpers import make_gpu_walkers, make_oracle_walker
from oracle import orc
ov = {"eq_mc_cycles": 500}
g, up = make_gpu_walkers("ice1_sample", nwalkers=2048, overrides=ov)
g.set_rng_philox(20141211, 0, 1000000)
for _ in range(6):
    g.mc_run(1000); g.mc_monitor()
g.mc_run(500)
st = g.states()
inc = np.array([list(s.model_energy) for s in st])
fresh = g.compute_model_energy_all()
d = np.abs(inc - fresh) / np.abs(fresh)
print("walkers with drift > 1e-12:", int((d.max(1) > 1e-12).sum()), "of", len(d), " max", d.max())
worst = int(np.argmax(d.max(1)))
print("worst walker", worst, d[worst], "errors", st[worst].error)
# the same walker on the oracle (same stream, same schedule)
o, _ = make_oracle_walker("ice1_sample", rank=worst, size=2048, overrides=ov)
o.set_rng_philox(20141211, worst, 1000000)
for _ in range(6):
    assert o.mc_run(1000) == 0; o.mc_monitor()
assert o.mc_run(500) == 0
oinc = np.array(o.model_energy).copy()
ofresh = np.array([o.compute_model_energy(1), o.compute_model_energy(2)])
print("oracle drift same walker:", np.abs(oinc - ofresh) / np.abs(ofresh))
ljr, ref, hm = g.download(worst)
print("positions bit-exact vs oracle:", np.array_equal(ljr, o.ljr), " accepted", list(st[worst].accepted), [o.geti("acc_r"), o.geti("acc_v"), o.geti("acc_s")])
print("GPU inc", inc[worst], "oracle inc", oinc, "GPU fresh", fresh[worst], "oracle fresh", ofresh)
