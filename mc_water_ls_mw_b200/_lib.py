"""ctypes loader for libmwgpu.so (the C ABI declared in include/mwgpu.h).

There is no CPU fallback: if the shared library is missing the import fails
loudly, and every compute entry point fails loudly when no CUDA device exists.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
# MWGPU_LIB_PATH: development aid (builds with other -D knobs side by side); the default is the in-tree library
LIB_PATH = os.environ.get("MWGPU_LIB_PATH") or os.path.join(_PKG, "libmwgpu.so")
CSRC = os.path.join(_PKG, "csrc")

MAXNEIGH = 50
MAXIVECT = 32
LIST_SLOTS = 32


class MwgpuError(RuntimeError):
    pass


class McParams(C.Structure):
    """mwgpu_mc_params (include/mwgpu.h)."""

    _fields_ = [
        ("temperature", C.c_double), ("pressure", C.c_double), ("npt", C.c_int),
        ("mc_max_trans", C.c_double), ("mc_dv_max", C.c_double), ("mc_target_ratio", C.c_double),
        ("wl_factor", C.c_double), ("wl_swetnam", C.c_int), ("wl_alpha", C.c_double),
        ("eta_interp", C.c_int), ("samplerun", C.c_int), ("leshift", C.c_int), ("nbins", C.c_int),
        ("mu_min", C.c_double), ("mu_max", C.c_double),
        ("allow_switch", C.c_int), ("allow_vol", C.c_int), ("allow_trans", C.c_int),
        ("mc_trans_prob", C.c_double), ("mc_vol_prob", C.c_double), ("mc_switch_prob", C.c_double),
        ("mc_always_switch", C.c_int), ("list_update_int", C.c_int), ("eq_mc_cycles", C.c_int),
        ("max_mc_cycles", C.c_int), ("eq_adjust_mc", C.c_int), ("monitor_int", C.c_int),
        ("dd", C.c_int), ("window_overlap", C.c_int),
        ("input_ref_enthalpy", C.c_double * 2), ("ls", C.c_int),
    ]


class WalkerState(C.Structure):
    """mwgpu_walker_state (include/mwgpu.h)."""

    _fields_ = [
        ("model_energy", C.c_double * 2), ("volume", C.c_double * 2), ("ls_mu", C.c_double),
        ("mc_max_trans", C.c_double), ("mc_dv_max", C.c_double), ("wl_factor", C.c_double),
        ("my_mu_min", C.c_double), ("my_mu_max", C.c_double),
        ("average_energy", C.c_double * 2), ("min_dmu", C.c_double), ("max_dmu", C.c_double),
        ("ref_enthalpy", C.c_double * 2),
        ("rng_index", C.c_int64), ("ls", C.c_int), ("mc_cycle_num", C.c_int),
        ("accepted", C.c_int * 3), ("attempted", C.c_int * 3),
        ("my_start_bin", C.c_int), ("my_end_bin", C.c_int), ("walker_in_window", C.c_int), ("error", C.c_int),
        ("wl_invt_active", C.c_int),
    ]


class FlatParams(C.Structure):
    """mwgpu_flat_params (include/mwgpu.h): userparams.f90:33-36."""

    _fields_ = [("wl_schedule", C.c_int), ("wl_minhist", C.c_int), ("wl_flattol", C.c_double), ("wl_useinvt", C.c_int)]


class FlatReport(C.Structure):
    """mwgpu_flat_report (include/mwgpu.h)."""

    _fields_ = [("checked", C.c_int), ("hist_reset", C.c_int), ("flat", C.c_int), ("invt_switched", C.c_int),
                ("mean", C.c_double), ("max_pct", C.c_double), ("min_pct", C.c_double), ("wl_factor", C.c_double)]


class ThermRow(C.Structure):
    """mwgpu_therm_row (include/mwgpu.h): the values of one row of <seed>RRR_therm.dat."""

    _fields_ = [("icyc", C.c_int64), ("ls", C.c_int64), ("model_energy", C.c_double * 2), ("ls_mu", C.c_double),
                ("volume", C.c_double * 2), ("hmatrix1", C.c_double * 9)]


# every symbol include/mwgpu.h declares: name -> (restype, argtypes)
_vp, _i, _d, _dp, _ip = C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int)
SYMBOLS = {
    "mwgpu_create": (_i, [_i, _i, _i, _i, C.POINTER(_vp)]),
    "mwgpu_destroy": (None, [_vp]),
    "mwgpu_last_error": (C.c_char_p, []),
    "mwgpu_device_count": (_i, []),
    "mwgpu_num_walkers": (_i, [_vp]),
    "mwgpu_upload": (_i, [_vp, _i, _dp, _dp, _dp]),
    "mwgpu_download": (_i, [_vp, _i, _dp, _dp, _dp]),
    "mwgpu_upload_all": (_i, [_vp, _dp, _dp, _dp]),
    "mwgpu_download_all": (_i, [_vp, _dp, _dp, _dp]),
    "mwgpu_energy_init": (_i, [_vp]),
    "mwgpu_compute_ivects": (_i, [_vp, _i, _i, _ip, _dp]),
    "mwgpu_compute_neighbours": (_i, [_vp, _i, _i, _ip, _ip, _ip]),
    "mwgpu_compute_model_energy": (_i, [_vp, _i, _i, _dp]),
    "mwgpu_compute_local_real_energy": (_i, [_vp, _i, _i, _i, _dp]),
    "mwgpu_compute_local_real_energy_all": (_i, [_vp, _i, _i, _dp]),
    "mwgpu_compute_neighbours_all": (_i, [_vp]),
    "mwgpu_compute_model_energy_all": (_i, [_vp, _dp]),
    "mwgpu_get_neighbours": (_i, [_vp, _i, _i, _ip, _ip, _ip]),
    "mwgpu_mc_init": (_i, [_vp, C.POINTER(McParams), _i, _i, _dp, _i, _d]),
    "mwgpu_mc_set_rng_philox": (_i, [_vp, C.c_uint64, C.c_uint32, C.c_uint64]),
    "mwgpu_mc_set_rng_fifo": (_i, [_vp, _dp, C.c_int64]),
    "mwgpu_mc_set_rng_index": (_i, [_vp, _i, C.c_uint64]),
    "mwgpu_mc_run": (_i, [_vp, _i]),
    "mwgpu_mc_run_async": (_i, [_vp, _i]),
    "mwgpu_mc_set_kernel": (_i, [_vp, _i]),
    "mwgpu_synchronize": (_i, [_vp]),
    "mwgpu_mc_get_state": (_i, [_vp, _i, C.POINTER(WalkerState)]),
    "mwgpu_mc_get_states": (_i, [_vp, C.POINTER(WalkerState)]),
    "mwgpu_mc_get_translations": (_i, [_vp, _i, _ip]),
    "mwgpu_mc_get_bins": (_i, [_vp, _i, _dp, _dp, _dp]),
    "mwgpu_mc_set_bins": (_i, [_vp, _i, _dp, _dp, _dp]),
    "mwgpu_mc_get_grid": (_i, [_vp, _dp, _dp, _dp]),
    "mwgpu_mc_set_wl_factor": (_i, [_vp, _i, _d, _i]),
    "mwgpu_mc_set_active_lattice": (_i, [_vp, _i, _i]),
    "mwgpu_mc_monitor": (_i, [_vp]),
    "mwgpu_mc_chain_sync": (_i, [_vp]),
    "mwgpu_comms_allreduce_bins": (_i, [_vp]),
    "mwgpu_comms_set_hist_base": (_i, [_vp, _dp, _dp]),
    "mwgpu_comms_get_unique_id": (_i, [_vp]),
    "mwgpu_comms_init": (_i, [_vp, _i, _i, _vp]),
    "mwgpu_comms_reduce_local": (_i, [_vp, C.POINTER(_vp), _ip]),
    "mwgpu_comms_apply": (_i, [_vp]),
    "mwgpu_mc_check_flatness": (_i, [_vp, C.POINTER(FlatParams), C.POINTER(FlatReport)]),
    "mwgpu_mc_deltag_from_hist": (_i, [_vp, _dp, _dp]),
    "mwgpu_comms_join_uhist": (_i, [_vp, _i, _dp]),
    "mwgpu_comms_join_eta": (_i, [_vp, _i, _dp]),
    "mwgpu_mc_restore": (_i, [_vp, _i, _i, _d, _d, _d, _i, _i, _dp, _dp, _dp, _dp, _dp, _dp]),
    "mwgpu_mc_set_therm": (_i, [_vp, _i, _i]),
    "mwgpu_mc_get_therm": (_i, [_vp, _i, C.POINTER(ThermRow), _i, _ip, _ip]),
    "mwgpu_timer_start": (_i, [_vp]),
    "mwgpu_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "mwgpu_last_kernel_ms": (_i, [_vp, C.POINTER(C.c_float)]),
    "mwgpu_measure_fp64_peak": (_i, [_i, _dp]),
    "mwgpu_kernel_launches": (_i, [_vp, C.POINTER(C.c_int64)]),
    "mwgpu_mc_set_schedule": (_i, [_vp, _i, _i]),
    "mwgpu_mc_get_walker_times": (_i, [_vp, C.POINTER(C.c_uint64)]),
}


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("mwgpu.cu", "mw2.cuh", "mw2_energy.cuh", "mw_mc.cuh", "mw_device.cuh", "Makefile")]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "mwgpu.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise MwgpuError("building libmwgpu.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MwgpuError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C mc_water_ls_mw_b200/csrc). "
            "There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)          # AttributeError == symbol missing from the library
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise MwgpuError(lib().mwgpu_last_error().decode(errors="replace") + f" [code {rc}]")
