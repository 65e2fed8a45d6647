"""The periodic section of ``mc_cycle`` (mc_moves.F90:257-316) driven from the host.

The reference calls ``mc_cycle`` once per cycle (main.f90:181-198); after the move loop every
cycle checks five intervals, in this order:

    mpi_sync_int   -> comms_allreduce_eta / _hist / _uhist        (two lattices, 'mw')
    monitor_int    -> mc_monitor_stats (+ the same all-reduces again, mc_moves.F90:1813-1821)
    flat_chk_int   -> mc_check_flatness                            (two lattices)
    latt_sync_int  -> mc_check_chain_synchronisation               (two lattices)
    deltaG_int     -> mc_compute_deltaG_from_hist                  (two lattices, samplerun)

``CycleSchedule.run(ncycles)`` advances a ``WalkerBatch`` by ``ncycles`` cycles with ONE kernel
launch per stretch between two such events and the device-side routine of each event at its
cycle, in the reference's order, so a whole weight-generation or sampling run needs no download
of the walker state.  Nothing here computes; every call goes through the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

from .walkers import WalkerBatch


@dataclass
class ScheduleLog:
    """What the reference would have written to its log at the events (rank 0 / walker 0)."""

    flatness: List[tuple] = field(default_factory=list)     # (cycle, FlatReport)
    deltaG: List[tuple] = field(default_factory=list)       # (cycle, deltaG_kT, normP)


class CycleSchedule:
    def __init__(self, batch: WalkerBatch, up, on_event: Optional[Callable[[str, int], None]] = None):
        """``up``: decks.UserParams (intervals, Wang-Landau schedule parameters)."""
        self.g = batch
        self.up = up
        self.cycle = 0                      # mc_cycle_num of the batch (all walkers advance together)
        self.log = ScheduleLog()
        self.on_event = on_event
        self.two = batch.nlat == 2
        self.mw = up.parallel_strategy == "mw"

    def _intervals(self):
        up = self.up
        iv = [up.monitor_int]
        if self.two:
            if self.mw:
                iv.append(up.mpi_sync_int)
            iv += [up.flat_chk_int, up.latt_sync_int]
            if up.samplerun:
                iv.append(up.deltaG_int)
        return [i for i in iv if i > 0]

    def _events(self, cyc: int) -> None:
        up, g = self.up, self.g
        if self.two and self.mw and up.mpi_sync_int > 0 and cyc % up.mpi_sync_int == 0:      # :258-276
            g.comms_allreduce_bins()
            self._note("sync", cyc)
        if up.monitor_int > 0 and cyc % up.monitor_int == 0:                                # :280-285
            g.mc_monitor()
            if self.two and self.mw:
                # mc_monitor_stats synchronises histogram, weights and (sample runs) the unbiased histogram itself
                # (mc_moves.F90:1813-1821): a no-op right after the mpi_sync_int merge, a real merge whenever
                # monitor_int is not a multiple of mpi_sync_int
                g.comms_allreduce_bins()
            self._note("monitor", cyc)
        if not self.two:
            return
        if up.flat_chk_int > 0 and cyc % up.flat_chk_int == 0:                              # :291-294
            rep = g.mc_check_flatness(up.wl_schedule, up.wl_minhist, up.wl_flattol, up.wl_useinvt)
            self.log.flatness.append((cyc, rep))
            self._note("flatness", cyc)
        if up.latt_sync_int > 0 and cyc % up.latt_sync_int == 0:                            # :297-300
            g.mc_chain_sync()
            self._note("chain_sync", cyc)
        if up.samplerun and up.deltaG_int > 0 and cyc % up.deltaG_int == 0:                 # :302-306
            dG, normP = g.mc_compute_deltaG_from_hist()
            self.log.deltaG.append((cyc, dG, normP))
            self._note("deltaG", cyc)

    def _note(self, what: str, cyc: int) -> None:
        if self.on_event:
            self.on_event(what, cyc)

    def run(self, ncycles: int) -> ScheduleLog:
        end = self.cycle + int(ncycles)
        ivs = self._intervals()
        while self.cycle < end:
            nxt = min([end] + [(self.cycle // i + 1) * i for i in ivs])
            self.g.mc_run(nxt - self.cycle)
            self.cycle = nxt
            self._events(nxt)
        return self.log
