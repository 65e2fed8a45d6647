"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), the library's own NCCL communicator.
Walkers are sharded in contiguous blocks; every rank checks its walkers against a single-process oracle
run of ALL walkers (the oracle is test infrastructure)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from mc_water_ls_mw_b200 import comms
from oracle import orc
from tests.helpers import make_gpu_walkers, make_oracle_walkers

SEED = 20141211


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = 3
    total = per * world

    def gpu_batch(ex, ov):
        from mc_water_ls_mw_b200 import walkers as W
        from tests.helpers import load_example
        up, h, r, w, wl = load_example(ex, size=total)
        for k, v in ov.items():
            setattr(up, k, v)
        g = W.WalkerBatch(up.nwater, up.num_lattices, per, device=local)
        g.upload(r, h); g.energy_init()
        g.mc_init(W.params_from_user(up), rank * per, total, w, wl)
        g.set_rng_philox(SEED, rank * per, 1000000)
        comms.init_nccl(g, rank, world)
        return g, up

    def oracle_all(ex, ov):
        ws = make_oracle_walkers(ex, total, size=total, overrides=ov)
        for i, s in enumerate(ws):
            s.set_rng_philox(SEED, i, 1000000)
        return ws

    # ---- 'mw': delta all-reduce, flatness check and deltaG over NCCL
    ov = {"eq_mc_cycles": 2}
    g, up = gpu_batch("ice1_gen_weights", ov)
    ws = oracle_all("ice1_gen_weights", ov)
    for it in range(3):
        g.mc_run(8); assert orc.mc_run_many(ws, 8, 4) == 0
        g.comms_allreduce_bins(); orc.allreduce_bins(ws)
        a = g.mc_check_flatness(1, -1, up.wl_flattol, False)
        b = orc.mc_check_flatness(ws, 1, -1, up.wl_flattol, False)
        assert (a.checked, a.hist_reset, a.flat) == (b.checked, b.hist_reset, b.flat), (it, rank)
        assert abs(a.mean - b.mean) <= 1e-12 * max(1.0, abs(b.mean))
        for w in range(per):
            s = ws[rank * per + w]
            wt, h, _ = g.bins(w)
            np.testing.assert_allclose(wt, s.weight, rtol=1e-11, atol=1e-11)
            np.testing.assert_allclose(h, s.histogram, rtol=0, atol=1e-9)
            assert abs(g.state(w).wl_factor - s.getd("wl_factor")) <= 1e-15
    ov = {"eq_mc_cycles": 2}
    g, up = gpu_batch("ice1_sample", ov)
    ws = oracle_all("ice1_sample", ov)
    g.mc_run(20); assert orc.mc_run_many(ws, 20, 4) == 0
    dg, npg = g.mc_compute_deltaG_from_hist()
    do, npo = orc.mc_deltaG_from_hist(ws)
    assert abs(dg - do) < 1e-10 * max(1.0, abs(do)), (dg, do)
    np.testing.assert_allclose(npg, npo, rtol=1e-9, atol=1e-300)

    # ---- 'dd': windows spread over the GPUs, joined with an NCCL all-gather
    ov = {"eq_mc_cycles": 100000}
    g, up = gpu_batch("ice1_sample_dd", ov)
    ws = oracle_all("ice1_sample_dd", ov)
    nb = g.nbins
    rng = np.random.default_rng(5)
    x = np.linspace(-2, 2, nb)
    for i, s in enumerate(ws):                      # same synthetic window contents on every rank
        u = np.exp(-x * x) * (2.0 ** i) * (1 + 0.01 * rng.standard_normal(nb))
        wt = x * x + 3.0 * i + 0.01 * rng.standard_normal(nb)
        s.unbiased_hist[:] = u; s.weight[:] = wt
        if i // per == rank:
            g.set_bins(i % per, weight=wt, unbiased_hist=u)
    for ovl in (0, 2):
        np.testing.assert_allclose(g.comms_join_uhist(ovl), orc.join_uhist(ws, ovl), rtol=1e-13)
        np.testing.assert_allclose(g.comms_join_eta(ovl), orc.join_eta(ws, ovl), rtol=1e-13, atol=1e-13)
    dist.barrier()
    if rank == 0:
        print("multigpu worker ok: world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
