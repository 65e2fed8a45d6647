"""What makes a walker slow: per-walker launch time against cell volume, energies, list lengths and acceptance.
Development aid.  usage: python scripts/diag/slow_walkers.py [walkers]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from mc_water_ls_mw_b200 import walkers as W
up, h, r, w, wl = bench._example()
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = W.WalkerBatch(up.nwater, up.num_lattices, nw)
g.upload(r, h); g.energy_init()
g.mc_init(W.params_from_user(up), 0, nw, w, wl)
g.set_kernel(2)
g.set_rng_philox(20141211, 0, 1000000)
for _ in range(2):
    g.mc_run(500); g.mc_monitor()
prev = None
for step in range(6):
    st0 = g.states()
    g.timer_start(); g.mc_run_async(250); ms = g.timer_stop()
    t = g.walker_times().astype(np.int64); d = (t[:, 1] - t[:, 0]) * 1e-6
    st = g.states()
    vol = np.array([s.volume for s in st]); E = np.array([s.model_energy for s in st])
    accv = np.array([s.attempted[1] - s0.attempted[1] for s, s0 in zip(st, st0)])
    accr = np.array([s.accepted[0] - s0.accepted[0] for s, s0 in zip(st, st0)])
    ls = np.array([s.ls for s in st])
    nn = []
    order = np.argsort(d)
    pick = list(order[:3]) + list(order[len(order) // 2 - 1:len(order) // 2 + 2]) + list(order[-5:])
    print(f"step {step}: {ms:.1f} ms; walker ms min {d.min():.1f} median {np.median(d):.1f} p90 {np.percentile(d, 90):.1f} p99 {np.percentile(d, 99):.1f} max {d.max():.1f}; "
          f"corr(d, vol1) {np.corrcoef(d, vol[:, 0])[0, 1]:.2f} corr(d, vol2) {np.corrcoef(d, vol[:, 1])[0, 1]:.2f} corr(d, nvol) {np.corrcoef(d, accv)[0, 1]:.2f} corr(d, acc) {np.corrcoef(d, accr)[0, 1]:.2f}"
          + (f" corr(d, d_prev) {np.corrcoef(d, prev)[0, 1]:.2f}" if prev is not None else ""))
    for wkr in pick:
        n1 = g.get_neighbours(1, int(wkr))[0]; n2 = g.get_neighbours(2, int(wkr))[0]
        print(f"   walker {wkr:4d}: {d[wkr]:6.1f} ms  vol {vol[wkr, 0]:7.1f} {vol[wkr, 1]:7.1f}  E {E[wkr, 0]:9.3f} {E[wkr, 1]:9.3f}  ls {ls[wkr]}  volume moves {accv[wkr]:3d}  accepted {accr[wkr]:5d}  list rows mean {n1.mean():.1f}/{n2.mean():.1f} max {n1.max()}/{n2.max()}")
    prev = d
