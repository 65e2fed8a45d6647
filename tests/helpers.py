"""Shared helpers for the tests (oracle side)."""
from __future__ import annotations

import os

import numpy as np

from mc_water_ls_mw_b200 import decks
from oracle import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXAMPLES = os.path.join(GOLDEN, "examples")


def example_dir(name: str) -> str:
    return os.path.join(EXAMPLES, name)


def load_example(name: str, size: int = 1):
    """(UserParams, hmatrix[nlat,9], ljr[nlat,N,3], weights-or-None, file_wl_factor)."""
    d = example_dir(name)
    up = decks.read_input(os.path.join(d, "ice.input"), size=size)
    h, r = decks.read_config(d, up)
    wpath = os.path.join(d, "eta_weights.dat")
    if up.num_lattices == 2 and os.path.exists(wpath):
        wl, _, w = decks.read_eta_weights(wpath)
    else:
        wl, w = 0.0, None
    return up, h, r, w, wl


def make_oracle_walker(name: str, rank: int = 0, size: int = 1, overrides: dict | None = None):
    up, h, r, w, wl = load_example(name, size=size)
    for k, v in (overrides or {}).items():
        setattr(up, k, v)
    s = orc.System(up.nwater, up.num_lattices)
    s.set_config(r, h)
    s.energy_init()
    for ils in range(1, up.num_lattices + 1):      # main.f90:125-128
        s.compute_model_energy(ils)
    rc = s.mc_init(orc.params_from_user(up), rank=rank, size=size, weights=w, file_wl_factor=wl)
    assert rc == 0
    return s, up


def make_gpu_walkers(name: str, nwalkers: int = 1, first_rank: int = 0, size: int | None = None,
                     overrides: dict | None = None, mc: bool = True):
    """A WalkerBatch initialised like the reference's start-up sequence (main.f90:98-175)."""
    from mc_water_ls_mw_b200 import walkers as W
    size = nwalkers if size is None else size
    up, h, r, w, wl = load_example(name, size=size)
    for k, v in (overrides or {}).items():
        setattr(up, k, v)
    g = W.WalkerBatch(up.nwater, up.num_lattices, nwalkers)
    g.upload(r, h)                      # the same xmol configuration for every walker
    g.energy_init()
    if mc:
        g.mc_init(W.params_from_user(up), first_rank, size, w, wl)
    return g, up


def make_oracle_walkers(name: str, nwalkers: int, first_rank: int = 0, size: int | None = None,
                        overrides: dict | None = None):
    size = nwalkers if size is None else size
    return [make_oracle_walker(name, rank=first_rank + w, size=size, overrides=overrides)[0] for w in range(nwalkers)]


def used_lists(nn, jn, vn):
    """Zero the unused tail of jn/vn rows (the reference leaves stale entries beyond nn)."""
    jn = np.array(jn, copy=True); vn = np.array(vn, copy=True)
    for i, n in enumerate(np.asarray(nn)):
        jn[i, n:] = 0; vn[i, n:] = 0
    return np.asarray(nn), jn, vn


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


class OracleSchedule:
    """The periodic section of mc_cycle (mc_moves.F90:257-316) over in-process oracle walkers
    (= MPI ranks): the checker's counterpart of mc_water_ls_mw_b200.schedule.CycleSchedule."""

    def __init__(self, walkers, up, nthreads: int = 0):
        self.ws, self.up, self.nthreads = walkers, up, nthreads
        self.cycle = 0
        self.flatness, self.deltaG = [], []

    def run(self, ncycles: int):
        up, ws = self.up, self.ws
        two, mw = ws[0].nlat == 2, up.parallel_strategy == "mw"
        for _ in range(int(ncycles)):
            assert orc.mc_run_many(ws, 1, self.nthreads) == 0
            self.cycle += 1
            c = self.cycle
            if two and mw and c % up.mpi_sync_int == 0:
                orc.allreduce_bins(ws)
            if c % up.monitor_int == 0:
                for s in ws:
                    s.mc_monitor()
                if two and mw:                               # mc_monitor_stats re-synchronises the bins (mc_moves.F90:1813-1821)
                    orc.allreduce_bins(ws)
            if not two:
                continue
            if c % up.flat_chk_int == 0:
                self.flatness.append((c, orc.mc_check_flatness(ws, up.wl_schedule, up.wl_minhist, up.wl_flattol, up.wl_useinvt)))
            if c % up.latt_sync_int == 0:
                for s in ws:
                    s.mc_chain_sync()
            if up.samplerun and c % up.deltaG_int == 0:
                self.deltaG.append((c,) + orc.mc_deltaG_from_hist(ws))
