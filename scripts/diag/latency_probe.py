"""Step latency of small ensembles (the strong-scaling shape: 4096 walkers over 8 / 4 / 2 GPUs) with 2 and 4 warps per
walker.  Development aid.  usage: python scripts/diag/latency_probe.py [walkers ...]"""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from mc_water_ls_mw_b200 import walkers as W
up, h, r, w, wl = bench._example()
if os.environ.get('MAXTRANS'):
    up.mc_max_trans = float(os.environ['MAXTRANS']); up.eq_adjust_mc = False    # acceptance ratio experiment
import os
for nw in [int(x) for x in sys.argv[1:]] or [512, 1024]:
    for kernel in [int(k) for k in os.environ.get('KERNELS', '2,4,0').split(',')]:
        g = W.WalkerBatch(up.nwater, up.num_lattices, nw)
        g.upload(r, h); g.energy_init()
        g.mc_init(W.params_from_user(up), 0, nw, w, wl)
        g.set_kernel(kernel)
        g.set_rng_philox(20141211, 0, 1000000)
        for _ in range(2):
            g.mc_run(500); g.mc_monitor()
        ts = []
        for i in range(8):
            g.timer_start(); g.mc_run_async(250); ts.append(g.timer_stop())
        t = g.walker_times().astype(np.int64); d = (t[:, 1] - t[:, 0]) * 1e-6
        st = g.states(); acc = np.mean([s.accepted[0] / max(1, s.attempted[0]) for s in st]); sw = np.mean([s.accepted[2] / max(1, s.attempted[2]) for s in st])
        print(f"walkers {nw} kernel {kernel}: step ms " + " ".join(f"{x:.1f}" for x in ts) + f" | walker ms min {d.min():.1f} mean {d.mean():.1f} max {d.max():.1f} | acc {acc:.2f} switch {sw:.2f}")
        del g
