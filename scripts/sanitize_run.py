"""Workload for scripts/sanitize.sh: a short mc_run of one deck (walkers x cycles) through the C ABI.
usage: sanitize_run.py <deck> <walkers> <cycles> [kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import make_gpu_walkers

deck, nw, ncyc = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
kern = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ov = {"eq_mc_cycles": 2, "mc_vol_prob": 0.02}
g, up = make_gpu_walkers(deck, nwalkers=nw, overrides=ov)
g.set_kernel(kern)
g.set_rng_philox(20141211, 0, 1000000)
g.mc_run(ncyc)
if up.num_lattices == 2:
    g.comms_allreduce_bins()
    g.mc_chain_sync()
g.mc_monitor()
e = g.compute_model_energy_all()
print("ok", deck, nw, ncyc, "kernel", kern, "E0", float(e[0, 0]), "launches", g.kernel_launches())
