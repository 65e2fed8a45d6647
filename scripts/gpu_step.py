"""Step GPU and oracle cycle by cycle and report the first divergence (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.helpers import make_oracle_walker, load_example
from mc_water_ls_mw_b200 import walkers as W

ex = sys.argv[1] if len(sys.argv) > 1 else "ice1_sample"
ncyc = int(sys.argv[2]) if len(sys.argv) > 2 else 20
o, up = make_oracle_walker(ex)
up2, h, r, wts, wl = load_example(ex)
g = W.WalkerBatch(up.nwater, up.num_lattices, 1)
g.upload(r, h); g.energy_init()
g.mc_init(W.params_from_user(up), 0, 1, wts, wl)
g.set_rng_philox(20141211, 0, 1000000); o.set_rng_philox(20141211, 0, 1000000)
for c in range(ncyc):
    try:
        g.mc_run(1)
    except Exception as e:
        print("cycle", c + 1, "GPU error:", e)
        s = g.state(); print(" state", list(s.accepted), list(s.attempted), s.rng_index, s.ls_mu)
        break
    o.mc_run(1)
    s = g.state()
    ljr, ref, hm = g.download()
    ok = (np.array_equal(ljr, o.ljr), np.array_equal(hm, o.hmatrix), list(s.accepted) == [o.geti("acc_r"), o.geti("acc_v"), o.geti("acc_s")], s.rng_index == o.geti("rng_index"))
    print("cycle", c + 1, ok, "acc", list(s.accepted), "att", list(s.attempted), "rng", s.rng_index, o.geti("rng_index"),
          "E", list(s.model_energy), list(o.model_energy), "mu", s.ls_mu, o.getd("ls_mu"))
    if not all(ok):
        print(" oracle", o.counters(), "max pos diff", np.max(np.abs(ljr - o.ljr)))
        break
