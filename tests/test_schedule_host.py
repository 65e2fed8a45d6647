"""Host logic of mc_water_ls_mw_b200.schedule.CycleSchedule without a GPU: a recording stand-in for the WalkerBatch
shows which C-ABI calls the schedule would issue; they must be the periodic section of mc_cycle
(mc_moves.F90:257-316, and the second merge inside mc_monitor_stats, :1813-1821) replayed cycle by cycle -- same
events, same order, one mc_run per stretch between two events, every cycle accounted for."""
from types import SimpleNamespace

import pytest

from mc_water_ls_mw_b200.schedule import CycleSchedule


class RecordingBatch:
    def __init__(self, nlat):
        self.nlat = nlat
        self.calls = []

    def mc_run(self, n):
        assert n > 0
        self.calls.append(("run", int(n)))

    def comms_allreduce_bins(self):
        self.calls.append(("allreduce",))

    def mc_monitor(self):
        self.calls.append(("monitor",))

    def mc_check_flatness(self, sched, minhist, tol, useinvt):
        self.calls.append(("flatness", sched, minhist, tol, useinvt))
        return SimpleNamespace(checked=1)

    def mc_chain_sync(self):
        self.calls.append(("chain_sync",))

    def mc_compute_deltaG_from_hist(self):
        self.calls.append(("deltaG",))
        return 0.0, None


def reference_calls(up, nlat, ncycles, start=0):
    """mc_moves.F90:257-316 cycle by cycle; runs of event-free cycles merged into one ('run', n)."""
    out, pending = [], 0
    for cyc in range(start + 1, start + ncycles + 1):
        pending += 1
        ev = []
        if nlat == 2 and cyc % up.mpi_sync_int == 0 and up.parallel_strategy == "mw":          # :257-276
            ev.append(("allreduce",))
        if cyc % up.monitor_int == 0:                                                           # :280-285
            ev.append(("monitor",))
            if nlat == 2 and up.parallel_strategy == "mw":                                      # :1813-1821
                ev.append(("allreduce",))
        if nlat == 2:
            if cyc % up.flat_chk_int == 0:                                                      # :291-294
                ev.append(("flatness", up.wl_schedule, up.wl_minhist, up.wl_flattol, up.wl_useinvt))
            if cyc % up.latt_sync_int == 0:                                                     # :297-300
                ev.append(("chain_sync",))
            if cyc % up.deltaG_int == 0 and up.samplerun:                                       # :302-306
                ev.append(("deltaG",))
        if ev or cyc == start + ncycles:
            out.append(("run", pending)); pending = 0
            out += ev
    return out


def _up(**kw):
    d = dict(parallel_strategy="mw", samplerun=True, mpi_sync_int=250, monitor_int=1000, flat_chk_int=500,
             latt_sync_int=2500, deltaG_int=2000, wl_schedule=0, wl_minhist=20, wl_flattol=0.1, wl_useinvt=True)
    d.update(kw)
    return SimpleNamespace(**d)


@pytest.mark.parametrize("nlat,kw,ncycles", [
    (2, {}, 5000),                                                        # the decks' defaults
    (2, {"samplerun": False}, 3000),                                      # weight generation: no deltaG
    (2, {"mpi_sync_int": 8, "monitor_int": 12, "flat_chk_int": 20, "latt_sync_int": 30, "deltaG_int": 45}, 400),
    (2, {"parallel_strategy": "dd", "monitor_int": 16, "flat_chk_int": 20, "latt_sync_int": 30, "deltaG_int": 40}, 200),
    (1, {"monitor_int": 7}, 50),                                          # single box: monitor only
    (2, {"mpi_sync_int": 1, "monitor_int": 1, "flat_chk_int": 1, "latt_sync_int": 1, "deltaG_int": 1}, 5),
])
def test_schedule_issues_the_reference_events_in_the_reference_order(nlat, kw, ncycles):
    up = _up(**kw)
    g = RecordingBatch(nlat)
    sch = CycleSchedule(g, up)
    sch.run(ncycles)
    assert g.calls == reference_calls(up, nlat, ncycles)
    assert sum(c[1] for c in g.calls if c[0] == "run") == ncycles and sch.cycle == ncycles


def test_schedule_resumes_between_calls_and_reports_events():
    """Several run() calls continue one cycle count (a stretch may end between two events), and the on_event hook sees
    every event with its cycle."""
    up = _up(mpi_sync_int=8, monitor_int=12, flat_chk_int=20, latt_sync_int=30, deltaG_int=45)
    g = RecordingBatch(2)
    seen = []
    sch = CycleSchedule(g, up, on_event=lambda what, cyc: seen.append((what, cyc)))
    want = []
    start = 0
    for n in (5, 11, 1, 40, 63):
        sch.run(n)
        want += reference_calls(up, 2, n, start)
        start += n
    assert g.calls == want and sch.cycle == 120
    assert [c for w, c in seen if w == "monitor"] == list(range(12, 121, 12))
    assert [c for w, c in seen if w == "deltaG"] == [45, 90]
    assert len(sch.log.flatness) == 6 and len(sch.log.deltaG) == 2
