"""Times the batched full-energy kernel on 4096 decorrelated ice1_sample walkers (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.helpers import make_gpu_walkers
g, up = make_gpu_walkers("ice1_sample", nwalkers=4096)
g.set_rng_philox(20141211, 0, 1000000)
for _ in range(4):
    g.mc_run(25); g.mc_monitor()
g.mc_run(int(os.environ.get("MW_EXTRA_CYCLES", "0")))
out = np.empty((4096, 2))
for _ in range(3):
    g.compute_model_energy_all(out)
ms = []
for _ in range(100):
    g.compute_model_energy_all(out); ms.append(g.last_kernel_ms())
t = float(np.mean(ms))
print(f"{sys.argv[1] if len(sys.argv) > 1 else ''}: {8192 / (t * 1e-3):.4g} evals/s ({t * 1e3:.1f} us)", flush=True)
