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
