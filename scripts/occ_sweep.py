"""Throughput of k_mc_run vs resident walkers per SM for one example deck (development aid).
usage: occ_sweep.py <example> <k1,k2,...>   (walkers = 148*k)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.helpers import make_gpu_walkers
ex = sys.argv[1]
for k in [int(x) for x in sys.argv[2].split(",")]:
    nw = 148 * k
    g, up = make_gpu_walkers(ex, nwalkers=nw)
    g.set_rng_philox(20141211, 0, 1000000)
    for _ in range(4):
        g.mc_run(25); g.mc_monitor()
    ms = []
    for _ in range(8):
        g.mc_run(10); ms.append(g.last_kernel_ms())
    t = float(np.median(ms))
    print(f"{ex} walkers/SM {k}: {nw * up.nwater * 10 / (t * 1e-3):.4g} moves/s  ({t:.3f} ms)", flush=True)
    del g
