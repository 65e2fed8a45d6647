"""Which evaluation path do the trial moves of the bench workload take? (development aid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.helpers import make_gpu_walkers

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g, up = make_gpu_walkers("ice1_sample", nwalkers=nw)
g.set_rng_philox(20141211, 0, 1000000)
for blk in range(8):
    g.mc_run(25)
    c = np.zeros(5, dtype=np.int64)
    for w in range(nw):
        c += np.array(g.path_counts(w))
    s = g.state(0)
    print(f"block {blk}: counts fast/dup/close/guard/forced = {c.tolist()}  frac fast = {c[0] / max(1, c.sum()):.3f}  "
          f"max_trans[0] = {s.mc_max_trans * 0.5291772108:.3f} A  acc/att = {s.accepted[0]}/{s.attempted[0]}")
    if blk < 4:
        g.mc_monitor()
