import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from mc_water_ls_mw_b200 import walkers as W
up, h, r, w, wl = bench._example()
nw = int(sys.argv[1])
g = W.WalkerBatch(up.nwater, up.num_lattices, nw)
g.upload(r, h); g.energy_init()
g.mc_init(W.params_from_user(up), 0, nw, w, wl)
g.set_rng_philox(20141211, 0, 1000000)
for _ in range(2):
    g.mc_run(500); g.mc_monitor()
ts = []
for i in range(16):
    g.timer_start(); g.mc_run_async(250); ms = g.timer_stop()
    if i % 2 == 1: g.comms_allreduce_bins()
    st = g.states()
    nb = 0
    ts.append(ms)
print(nw, "kernel ms per step:", " ".join(f"{t:.1f}" for t in ts))
st = g.states()
acc = np.array([s.accepted[0] / max(1, s.attempted[0]) for s in st]); vol = np.array([s.volume[0] for s in st])
print("acc ratio mean %.3f  vol mean %.1f min %.1f max %.1f  max_trans mean %.3f dv mean %.4f max %.3f" % (acc.mean(), vol.mean(), vol.min(), vol.max(), np.mean([s.mc_max_trans for s in st]), np.mean([s.mc_dv_max for s in st]), np.max([s.mc_dv_max for s in st])))
