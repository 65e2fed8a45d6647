"""Multi-GPU plumbing: one process per GPU, walkers sharded across ranks.

The reference runs one walker per MPI rank and merges the lattice-switch
weights / histograms with three MPI_Allreduce calls of the increments since
the last synchronisation (comms_mpi.f90:244-277, :461-530), every
``mpi_sync_int`` cycles (mc_moves.F90:258-278).  Here a rank owns a contiguous
block of walkers on one B200; the increments of its walkers are summed on the
device (``mwgpu_comms_reduce_local``), the per-rank sums are all-reduced over
NVLink with NCCL, and every walker is re-based on the device
(``mwgpu_comms_apply``).  There is no other data-path exchange: walkers are
independent Markov chains.

Two transports for the one collective:
* ``init_nccl`` / ``WalkerBatch.comms_allreduce_bins``: the library's own NCCL
  communicator (what a Fortran host gets through the C ABI);
* ``allreduce_bins_torch``: ``torch.distributed.all_reduce`` on the device
  buffer (NCCL backend) -- the same code path runs over ``gloo`` on CPU tensors
  for the world_size-2 tests.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_walkers(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition: (first global walker id, number of walkers) of ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def allreduce_array(arr: np.ndarray, group=None) -> np.ndarray:
    """Sum a host array over all ranks (gloo/NCCL via torch.distributed); in place, returns arr."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(arr)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return arr


class _DevBuf:
    """Wraps a raw device pointer for torch.as_tensor via __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def allreduce_bins_torch(batch, group=None) -> None:
    """Delta all-reduce of one rank's WalkerBatch through torch.distributed (NCCL backend)."""
    import torch
    import torch.distributed as dist
    ptr, n = batch.comms_reduce_local()           # synchronises the context's stream
    t = torch.as_tensor(_DevBuf(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream().synchronize()
    batch.comms_apply()


def init_nccl(batch, rank: int, world_size: int) -> None:
    """Create the library-side NCCL communicator; the unique id travels over torch.distributed."""
    import torch.distributed as dist
    from . import walkers
    obj = [walkers.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    batch.comms_init_nccl(world_size, rank, obj[0])


def delta_merge_host(arrays, bases, group=None):
    """Reference semantics on host arrays (used by the CPU tests and by hosts that keep the bins
    on the CPU): arrays[w] <- base[w] + sum_over_all_walkers_of_all_ranks(arrays - bases); base <- arrays."""
    local = np.zeros_like(arrays[0])
    for a, b in zip(arrays, bases):
        local += a - b
    total = allreduce_array(local, group)
    for a, b in zip(arrays, bases):
        a[:] = total + b
        b[:] = a
    return total


def gather_windows_host(mine: np.ndarray, group=None) -> np.ndarray:
    """All ranks' [W][nbins] blocks in rank order = window order (what the device path does with one
    ncclAllGather before comms_join_*): returns [size][nbins].  Every rank must own the same number of walkers."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(mine, dtype=np.float64))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return np.concatenate([o.numpy() for o in out], axis=0)


def join_windows_host(arrs: np.ndarray, overlap: int, eta: bool) -> np.ndarray:
    """comms_join_eta (eta=True, comms_mpi.f90:377-459) / comms_join_uhist (eta=False, :299-375) on the
    gathered [size][nbins] windows, in the reference's order of operations (host mirror of k_join_windows)."""
    size, nb = arrs.shape
    bpw = nb // size
    joined = arrs[0].astype(np.float64).copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        for ir in range(1, size):
            e = ir * bpw
            myave = 0.0; nextav = 0.0
            for k in range(e - overlap, e + overlap + 1):
                myave = myave + (joined[k - 1] if eta else np.log(joined[k - 1]))
            myave = myave / float(2 * overlap + 1)
            for k in range(e - overlap, e + overlap + 1):
                nextav = nextav + (arrs[ir][k - 1] if eta else np.log(arrs[ir][k - 1]))
            nextav = nextav / float(2 * overlap + 1)
            shift = myave - nextav
            if eta:
                joined[e:] = arrs[ir][e:] + shift
            else:
                if np.isnan(shift):
                    shift = 0.0
                joined[e:] = arrs[ir][e:] * np.exp(shift)
    if eta:
        joined = joined - joined[nb // 2]
    return joined
