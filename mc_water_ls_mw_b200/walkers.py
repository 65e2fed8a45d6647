"""Host-side mirror of the reference interface for the hot path.

``WalkerBatch`` owns ``nwalkers`` independent walker boxes on one B200 (one
walker == one MPI rank of the reference) and exposes, with the reference's
names, argument meaning (1-based ``ils``/``imol``) and error behaviour:

* module ``energy`` (molint.F90:22-37): ``energy_init``, ``compute_ivects``,
  ``compute_neighbours``, ``compute_model_energy``, ``compute_local_real_energy``
* the move loop of ``mc_moves`` (mc_moves.F90:117-255): ``mc_init``, ``mc_cycle``
  (= ``mc_run(1)``), ``mc_run``, plus the state effects of the periodic
  bookkeeping (``mc_monitor``, ``mc_chain_sync``, ``comms_allreduce_bins``).

Everything goes through the C ABI in include/mwgpu.h; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import FlatParams, FlatReport, McParams, MwgpuError, ThermRow, WalkerState, check, lib


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


def params_from_user(up) -> McParams:
    """decks.UserParams (already in internal units, io.f90:165-186) -> mwgpu_mc_params."""
    p = McParams()
    p.temperature = up.temperature; p.pressure = up.pressure; p.npt = int(up.mc_ensemble == "npt")
    p.mc_max_trans = up.mc_max_trans; p.mc_dv_max = up.mc_dv_max; p.mc_target_ratio = up.mc_target_ratio
    p.wl_factor = up.wl_factor; p.wl_swetnam = int(up.wl_swetnam); p.wl_alpha = up.wl_alpha
    p.eta_interp = int(up.eta_interp); p.samplerun = int(up.samplerun); p.leshift = int(up.leshift)
    p.nbins = up.nbins; p.mu_min = up.mu_min; p.mu_max = up.mu_max
    p.allow_switch = int(up.allow_switch); p.allow_vol = int(up.allow_vol); p.allow_trans = int(up.allow_trans)
    p.mc_trans_prob = up.mc_trans_prob; p.mc_vol_prob = up.mc_vol_prob; p.mc_switch_prob = up.mc_switch_prob
    p.mc_always_switch = int(up.mc_always_switch); p.list_update_int = up.list_update_int
    p.eq_mc_cycles = up.eq_mc_cycles; p.max_mc_cycles = up.max_mc_cycles
    p.eq_adjust_mc = int(up.eq_adjust_mc); p.monitor_int = up.monitor_int
    p.dd = int(up.parallel_strategy == "dd"); p.window_overlap = up.window_overlap
    p.input_ref_enthalpy[0] = up.input_ref_enthalpy[0]; p.input_ref_enthalpy[1] = up.input_ref_enthalpy[1]
    p.ls = up.ls
    return p


class WalkerBatch:
    def __init__(self, nwater: int, num_lattices: int, nwalkers: int = 1, device: int = 0):
        self.L = lib()
        self.nwater, self.nlat, self.nwalkers = int(nwater), int(num_lattices), int(nwalkers)
        h = C.c_void_p()
        check(self.L.mwgpu_create(self.nwater, self.nlat, self.nwalkers, int(device), C.byref(h)))
        self.h = h
        self.nbins = 0
        self._keep = None

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.mwgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ model state
    def upload(self, ljr: np.ndarray, hmatrix: np.ndarray, ref_ljr: Optional[np.ndarray] = None, walker: int = -1) -> None:
        """ljr[nlat,nwater,3] (== Fortran ljr(3,1,nwater,nlat)), hmatrix[nlat,9] column-major, Bohr."""
        ljr = np.ascontiguousarray(ljr, dtype=np.float64).reshape(self.nlat, self.nwater, 3)
        hm = np.ascontiguousarray(hmatrix, dtype=np.float64).reshape(self.nlat, 9)
        ref = None if ref_ljr is None else np.ascontiguousarray(ref_ljr, dtype=np.float64).reshape(self.nlat, self.nwater, 3)
        check(self.L.mwgpu_upload(self.h, walker, _dp(ljr), _dp(ref), _dp(hm)))

    def upload_all(self, ljr: np.ndarray, hmatrix: np.ndarray, ref_ljr: Optional[np.ndarray] = None) -> None:
        """Per-walker arrays ljr[nwalkers,nlat,nwater,3], hmatrix[nwalkers,nlat,9] (may be pinned host memory)."""
        assert ljr.dtype == np.float64 and ljr.flags.c_contiguous and ljr.size == self.nwalkers * self.nlat * self.nwater * 3
        assert hmatrix.dtype == np.float64 and hmatrix.flags.c_contiguous and hmatrix.size == self.nwalkers * self.nlat * 9
        check(self.L.mwgpu_upload_all(self.h, _dp(ljr), _dp(ref_ljr), _dp(hmatrix)))

    def download(self, walker: int = 0):
        ljr = np.empty((self.nlat, self.nwater, 3)); ref = np.empty_like(ljr); hm = np.empty((self.nlat, 9))
        check(self.L.mwgpu_download(self.h, walker, _dp(ljr), _dp(ref), _dp(hm)))
        return ljr, ref, hm

    def download_all(self):
        ljr = np.empty((self.nwalkers, self.nlat, self.nwater, 3)); ref = np.empty_like(ljr)
        hm = np.empty((self.nwalkers, self.nlat, 9))
        check(self.L.mwgpu_download_all(self.h, _dp(ljr), _dp(ref), _dp(hm)))
        return ljr, ref, hm

    # ------------------------------------------------------------------ module energy
    def energy_init(self) -> None:
        check(self.L.mwgpu_energy_init(self.h))

    def compute_ivects(self, ils: int, walker: int = 0):
        n = C.c_int(0)
        iv = np.zeros((_lib.MAXIVECT, 3))
        check(self.L.mwgpu_compute_ivects(self.h, walker, ils, C.byref(n), _dp(iv)))
        return n.value, iv

    def _lists(self, fn, ils: int, walker: int):
        nn = np.zeros(self.nwater, dtype=np.int32)
        jn = np.zeros((self.nwater, _lib.MAXNEIGH), dtype=np.int32)
        vn = np.zeros((self.nwater, _lib.MAXNEIGH), dtype=np.int32)
        check(fn(self.h, walker, ils, _ip(nn), _ip(jn), _ip(vn)))
        return nn, jn, vn

    def compute_neighbours(self, ils: int, walker: int = 0):
        return self._lists(self.L.mwgpu_compute_neighbours, ils, walker)

    def get_neighbours(self, ils: int, walker: int = 0):
        return self._lists(self.L.mwgpu_get_neighbours, ils, walker)

    def compute_neighbours_all(self) -> None:
        check(self.L.mwgpu_compute_neighbours_all(self.h))

    def compute_model_energy(self, ils: int, walker: int = 0) -> float:
        e = C.c_double(0.0)
        check(self.L.mwgpu_compute_model_energy(self.h, walker, ils, C.byref(e)))
        return e.value

    def compute_model_energy_all(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.nwalkers, self.nlat))
        check(self.L.mwgpu_compute_model_energy_all(self.h, _dp(out)))
        return out

    def compute_local_real_energy(self, imol: int, ils: int, walker: int = 0) -> float:
        e = C.c_double(0.0)
        check(self.L.mwgpu_compute_local_real_energy(self.h, walker, imol, ils, C.byref(e)))
        return e.value

    def compute_local_real_energy_all(self, ils: int, walker: int = 0) -> np.ndarray:
        e = np.zeros(self.nwater)
        check(self.L.mwgpu_compute_local_real_energy_all(self.h, walker, ils, _dp(e)))
        return e

    # ------------------------------------------------------------------ mc_moves
    def mc_init(self, params: McParams, first_rank: int = 0, size: Optional[int] = None,
                weights: Optional[np.ndarray] = None, file_wl_factor: float = 0.0) -> None:
        size = self.nwalkers if size is None else size
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.float64)
            check(self.L.mwgpu_mc_init(self.h, C.byref(params), first_rank, size, _dp(w), len(w), file_wl_factor))
        else:
            check(self.L.mwgpu_mc_init(self.h, C.byref(params), first_rank, size, None, 0, 0.0))
        self.nbins = params.nbins + (1 - params.nbins % 2)

    def set_rng_philox(self, seed: int, first_stream: int = 0, start_index: int = 0) -> None:
        check(self.L.mwgpu_mc_set_rng_philox(self.h, seed, first_stream, start_index))

    def set_rng_index(self, index: int, walker: int = -1) -> None:
        """Next draw index of one walker (-1: all), e.g. after mc_restore (mwgpu_mc_set_rng_index)."""
        check(self.L.mwgpu_mc_set_rng_index(self.h, walker, int(index)))

    def set_rng_fifo(self, u: np.ndarray) -> None:
        u = np.ascontiguousarray(u, dtype=np.float64)
        check(self.L.mwgpu_mc_set_rng_fifo(self.h, _dp(u), len(u)))

    def mc_run(self, ncycles: int) -> None:
        check(self.L.mwgpu_mc_run(self.h, int(ncycles)))

    def mc_cycle(self) -> None:
        self.mc_run(1)

    def set_kernel(self, warps_per_walker: int) -> None:
        """0 = automatic, 1 = one warp per walker, 2 = one warp per lattice, 4 = two warps per lattice (mwgpu_mc_set_kernel)."""
        check(self.L.mwgpu_mc_set_kernel(self.h, int(warps_per_walker)))

    def mc_run_async(self, ncycles: int) -> None:
        check(self.L.mwgpu_mc_run_async(self.h, int(ncycles)))

    def synchronize(self) -> None:
        check(self.L.mwgpu_synchronize(self.h))

    def state(self, walker: int = 0) -> WalkerState:
        s = WalkerState()
        check(self.L.mwgpu_mc_get_state(self.h, walker, C.byref(s)))
        return s

    def states(self):
        arr = (WalkerState * self.nwalkers)()
        check(self.L.mwgpu_mc_get_states(self.h, arr))
        return arr

    def translations(self, walker: int = 0) -> np.ndarray:
        t = np.zeros(self.nwater, dtype=np.int32)
        check(self.L.mwgpu_mc_get_translations(self.h, walker, _ip(t)))
        return t

    def bins(self, walker: int = 0):
        w = np.zeros(self.nbins); h = np.zeros(self.nbins); u = np.zeros(self.nbins)
        check(self.L.mwgpu_mc_get_bins(self.h, walker, _dp(w), _dp(h), _dp(u)))
        return w, h, u

    def set_bins(self, walker: int, weight=None, histogram=None, unbiased_hist=None) -> None:
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (weight, histogram, unbiased_hist)]
        check(self.L.mwgpu_mc_set_bins(self.h, walker, _dp(arrs[0]), _dp(arrs[1]), _dp(arrs[2])))

    def grid(self):
        mu = np.zeros(self.nbins); bw = np.zeros(self.nbins); sc = np.zeros(4)
        check(self.L.mwgpu_mc_get_grid(self.h, _dp(mu), _dp(bw), _dp(sc)))
        return mu, bw, {"r_pos": sc[0], "r_neg": sc[1], "av_binwidth": sc[2], "log_unbiased_norm": sc[3]}

    def set_wl_factor(self, wl_factor: float, wl_invt_active: bool = False, walker: int = -1) -> None:
        check(self.L.mwgpu_mc_set_wl_factor(self.h, walker, float(wl_factor), int(wl_invt_active)))

    def set_active_lattice(self, ls: int, walker: int = -1) -> None:
        check(self.L.mwgpu_mc_set_active_lattice(self.h, walker, int(ls)))

    def mc_monitor(self) -> None:
        check(self.L.mwgpu_mc_monitor(self.h))

    def mc_chain_sync(self) -> None:
        check(self.L.mwgpu_mc_chain_sync(self.h))

    # ------------------------------------------------------------------ comms
    def comms_allreduce_bins(self) -> None:
        check(self.L.mwgpu_comms_allreduce_bins(self.h))

    def comms_set_hist_base(self, histogram=None, unbiased_hist=None) -> None:
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (histogram, unbiased_hist)]
        check(self.L.mwgpu_comms_set_hist_base(self.h, _dp(arrs[0]), _dp(arrs[1])))

    def comms_reduce_local(self):
        """(device pointer, count) of the summed increments, for an external all-reduce."""
        p = C.c_void_p(); n = C.c_int(0)
        check(self.L.mwgpu_comms_reduce_local(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def comms_apply(self) -> None:
        check(self.L.mwgpu_comms_apply(self.h))

    def comms_join_uhist(self, overlap: int) -> np.ndarray:
        """comms_join_uhist (comms_mpi.f90:299-375): unbiased histograms of the dd windows stitched."""
        out = np.zeros(self.nbins, dtype=np.float64)
        check(self.L.mwgpu_comms_join_uhist(self.h, int(overlap), _dp(out)))
        return out

    def comms_join_eta(self, overlap: int) -> np.ndarray:
        """comms_join_eta (comms_mpi.f90:377-459)."""
        out = np.zeros(self.nbins, dtype=np.float64)
        check(self.L.mwgpu_comms_join_eta(self.h, int(overlap), _dp(out)))
        return out

    # ------------------------------------------------------------------ bookkeeping on the reduced arrays
    def mc_check_flatness(self, wl_schedule: int = 0, wl_minhist: int = 20,
                          wl_flattol: float = float(np.float32(0.05)), wl_useinvt: bool = False) -> FlatReport:
        """mc_check_flatness (mc_moves.F90:1936-2185) for every walker; the report is walker 0's."""
        fp = FlatParams(int(wl_schedule), int(wl_minhist), float(wl_flattol), int(wl_useinvt))
        rep = FlatReport()
        check(self.L.mwgpu_mc_check_flatness(self.h, C.byref(fp), C.byref(rep)))
        return rep

    def mc_compute_deltaG_from_hist(self):
        """mc_compute_deltaG_from_hist (mc_moves.F90:2498-2621): (deltaG in kT for the box, normP)."""
        dG = C.c_double(0.0)
        normP = np.zeros(self.nbins, dtype=np.float64)
        check(self.L.mwgpu_mc_deltag_from_hist(self.h, C.byref(dG), _dp(normP)))
        return float(dG.value), normP

    # ------------------------------------------------------------------ restart, therm rows
    def mc_restore(self, walker: int, rec: dict) -> None:
        """mc_checkpoint_load (mc_moves.F90:403-501) + the refresh of mc_init (:842-862) for one walker of a batch
        that went through the normal start-up; ``rec`` = decks.read_checkpoint(...) / checkpoint_record()."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        h, w, u = f(rec["histogram"]), f(rec["weight"]), f(rec["unbiased_hist"])
        hm, ref, ljr = f(rec["hmatrix"]), f(rec["ref_ljr"]), f(rec["ljr"])
        check(self.L.mwgpu_mc_restore(self.h, int(walker), int(rec["mc_cycle_num"]), float(rec["mc_max_trans"]),
                                      float(rec["mc_dv_max"]), float(rec["wl_factor"]), int(rec["wl_invt_active"]),
                                      int(rec["ls"]), _dp(h), _dp(w), _dp(u), _dp(hm), _dp(ref), _dp(ljr)))

    def checkpoint_record(self, walker: int = 0) -> dict:
        """Everything mc_checkpoint_write (mc_moves.F90:324-388) puts into checkpointRRR.dat.N for one walker."""
        s = self.state(walker)
        w, h, u = self.bins(walker)
        ljr, ref, hm = self.download(walker)
        return dict(nwater=self.nwater, mc_cycle_num=s.mc_cycle_num, mc_max_trans=s.mc_max_trans, mc_dv_max=s.mc_dv_max,
                    wl_factor=s.wl_factor, histogram=h, weight=w, wl_invt_active=bool(s.wl_invt_active),
                    unbiased_hist=u, hmatrix=hm, ref_ljr=ref, ljr=ljr, ls=s.ls)

    def set_therm(self, file_output_int: int, capacity: int = 32) -> None:
        """Record the values of a therm row (main.f90:200-223) every ``file_output_int`` cycles in the kernel."""
        check(self.L.mwgpu_mc_set_therm(self.h, int(file_output_int), int(capacity)))
        self._therm_cap = int(capacity)

    def therm(self, walker: int = 0):
        """Drain the recorded rows of one walker: (list of ThermRow, number of rows dropped)."""
        cap = getattr(self, "_therm_cap", 0)
        rows = (ThermRow * max(cap, 1))()
        n = C.c_int(0); nd = C.c_int(0)
        check(self.L.mwgpu_mc_get_therm(self.h, int(walker), rows, max(cap, 1), C.byref(n), C.byref(nd)))
        return [rows[i] for i in range(n.value)], nd.value

    def comms_init_nccl(self, nranks: int, rank: int, unique_id: bytes) -> None:
        buf = C.create_string_buffer(unique_id, 128)
        check(self.L.mwgpu_comms_init(self.h, nranks, rank, C.cast(buf, C.c_void_p)))

    # ------------------------------------------------------------------ measurement
    def timer_start(self) -> None:
        check(self.L.mwgpu_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0.0)
        check(self.L.mwgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0.0)
        check(self.L.mwgpu_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def kernel_launches(self) -> int:
        n = C.c_int64(0)
        check(self.L.mwgpu_kernel_launches(self.h, C.byref(n)))
        return n.value

    def set_schedule(self, chunk_cycles: int = 0, max_blocks: int = 0) -> None:
        """Cycles per unit of work and number of persistent blocks of the walker kernel (0 = automatic);
        never changes a result."""
        check(self.L.mwgpu_mc_set_schedule(self.h, int(chunk_cycles), int(max_blocks)))

    def walker_times(self) -> np.ndarray:
        """[W, 2] uint64: %globaltimer ns at the start / end of every walker's part of the last mc_run launch."""
        t = np.zeros((self.nwalkers, 2), dtype=np.uint64)
        check(self.L.mwgpu_mc_get_walker_times(self.h, t.ctypes.data_as(C.POINTER(C.c_uint64))))
        return t


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().mwgpu_comms_get_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def measure_fp64_peak(device: int = 0) -> float:
    t = C.c_double(0.0)
    check(lib().mwgpu_measure_fp64_peak(device, C.byref(t)))
    return t.value


def device_count() -> int:
    return lib().mwgpu_device_count()
