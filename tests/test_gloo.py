"""CPU, world_size 2, gloo: the N>1 host logic -- walker sharding and the delta all-reduce of
weights / histograms (comms_mpi.f90:244-277, :461-530) -- against a single-process oracle run."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from mc_water_ls_mw_b200 import comms


def test_shard_walkers_partition():
    for total, ws in [(4096, 8), (4096, 3), (7, 2), (5, 8)]:
        seen = []
        for r in range(ws):
            first, n = comms.shard_walkers(total, ws, r)
            seen += list(range(first, first + n))
        assert seen == list(range(total))
    with pytest.raises(ValueError):
        comms.shard_walkers(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, ncyc, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.helpers import make_oracle_walker
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = comms.shard_walkers(total, world, rank)
    ov = {"eq_mc_cycles": 2}
    ws = []
    for g in range(first, first + n):
        w, _ = make_oracle_walker("ice1_gen_weights", rank=g, size=total, overrides=ov)
        w.set_rng_philox(20141211, g, 1000000)
        ws.append(w)
    for _ in range(2):
        for w in ws:
            assert w.mc_run(ncyc) == 0
        comms.delta_merge_host([w.weight for w in ws], [w.arr_d("eta_last_sync", (101,)) for w in ws])
        comms.delta_merge_host([w.histogram for w in ws], [w.arr_d("hist_last_sync", (101,)) for w in ws])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), weight=np.array([w.weight for w in ws]),
             hist=np.array([w.histogram for w in ws]), ljr=np.array([w.ljr for w in ws]))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_delta_allreduce_matches_single_process(tmp_path):
    from oracle import orc
    from tests.helpers import make_oracle_walker
    total, ncyc, world = 5, 6, 2
    mp.spawn(_worker, args=(world, _free_port(), total, ncyc, str(tmp_path)), nprocs=world, join=True)
    ov = {"eq_mc_cycles": 2}
    ref = []
    for g in range(total):
        w, _ = make_oracle_walker("ice1_gen_weights", rank=g, size=total, overrides=ov)
        w.set_rng_philox(20141211, g, 1000000)
        ref.append(w)
    for _ in range(2):
        for w in ref:
            assert w.mc_run(ncyc) == 0
        orc.allreduce_bins(ref)
    got_w = np.concatenate([np.load(tmp_path / f"rank{r}.npz")["weight"] for r in range(world)])
    got_h = np.concatenate([np.load(tmp_path / f"rank{r}.npz")["hist"] for r in range(world)])
    got_x = np.concatenate([np.load(tmp_path / f"rank{r}.npz")["ljr"] for r in range(world)])
    for g in range(total):
        np.testing.assert_allclose(got_w[g], ref[g].weight, rtol=0, atol=1e-12)   # summation order differs
        np.testing.assert_allclose(got_h[g], ref[g].histogram, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(got_x[g], ref[g].ljr)                        # chains are untouched by sharding
    assert got_h.sum() > 0
    np.testing.assert_array_equal(got_w[0], got_w[-1])


def _join_worker(rank, world, port, per, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb = 101
    x = np.linspace(-2, 2, nb)
    rng = np.random.default_rng(100 + rank)
    mine_u = np.array([np.exp(-x * x) * 2.0 ** (rank * per + w) * (1 + 0.01 * rng.standard_normal(nb)) for w in range(per)])
    mine_w = np.array([x * x + 3.0 * (rank * per + w) + 0.01 * rng.standard_normal(nb) for w in range(per)])
    all_u = comms.gather_windows_host(mine_u)
    all_w = comms.gather_windows_host(mine_w)
    np.savez(os.path.join(out_dir, f"join{rank}.npz"), mine_u=mine_u, mine_w=mine_w,
             ju=comms.join_windows_host(all_u, 2, False), jw=comms.join_windows_host(all_w, 2, True))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_window_join_matches_oracle(tmp_path):
    """dd windows spread over two ranks: all-gather in rank order + the reference's stitch (comms_mpi.f90:299-459)
    against the oracle's restatement over all four windows in one process."""
    from oracle import orc
    from tests.helpers import make_oracle_walkers
    world, per = 2, 2
    mp.spawn(_join_worker, args=(world, _free_port(), per, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"join{r}.npz") for r in range(world)]
    ws = make_oracle_walkers("ice1_sample_dd", world * per, size=world * per, overrides={"eq_mc_cycles": 100000})
    for r in range(world):
        for w in range(per):
            ws[r * per + w].unbiased_hist[:] = parts[r]["mine_u"][w]
            ws[r * per + w].weight[:] = parts[r]["mine_w"][w]
    for r in range(world):                                           # every rank ends with the same joined arrays
        np.testing.assert_allclose(parts[r]["ju"], orc.join_uhist(ws, 2), rtol=1e-13)
        np.testing.assert_allclose(parts[r]["jw"], orc.join_eta(ws, 2), rtol=1e-13, atol=1e-13)
