"""Batched full-energy kernel on thermalised walkers: evals/s of both generations (development aid).
usage: energy_bench.py [replicas of the 4096 decorrelated walkers, default 8]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mc_water_ls_mw_b200 import walkers as W
from tests.helpers import make_gpu_walkers

rep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
g, up = make_gpu_walkers("ice1_sample", nwalkers=4096)
g.set_rng_philox(20141211, 0, 1000000)
for _ in range(4):
    g.mc_run(25); g.mc_monitor()
ljr, ref, hm = g.download_all()
big = W.WalkerBatch(up.nwater, up.num_lattices, 4096 * rep)
big.upload_all(np.ascontiguousarray(np.tile(ljr, (rep, 1, 1, 1))), np.ascontiguousarray(np.tile(hm, (rep, 1, 1))))
big.energy_init()
out = np.empty((4096 * rep, up.num_lattices))
res = {}
for kern in (1, 0):
    big.set_kernel(kern)
    for _ in range(3):
        big.compute_model_energy_all(out)
    ms = []
    for _ in range(10):
        big.compute_model_energy_all(out); ms.append(big.last_kernel_ms())
    res[kern] = out.copy()
    n = 4096 * rep * up.num_lattices
    print(f"energy kernel generation {2 - kern if kern else 2}: {n} evals in {np.mean(ms):.3f} ms = {n / np.mean(ms) * 1e3:.4g} evals/s")
print("max rel diff between generations", float(np.max(np.abs(res[0] - res[1]) / np.abs(res[1]))))
