"""ctypes binding of the CPU ORACLE (oracle/libmw_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(mc_water_ls_mw_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmw_oracle.so")
MAXNEIGH = 50
MAXIVECT = 125


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("mw_oracle.c", "mw_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


class FlatParams(C.Structure):
    _fields_ = [("wl_schedule", C.c_int), ("wl_minhist", C.c_int), ("wl_flattol", C.c_double), ("wl_useinvt", C.c_int)]


class FlatReport(C.Structure):
    _fields_ = [("checked", C.c_int), ("hist_reset", C.c_int), ("flat", C.c_int), ("invt_switched", C.c_int),
                ("mean", C.c_double), ("max_pct", C.c_double), ("min_pct", C.c_double), ("wl_factor", C.c_double)]


class Params(C.Structure):
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


_lib: Optional[C.CDLL] = None
_lib_fast: Optional[C.CDLL] = None
_FAST_DIR = os.path.join(_HERE, "_fast")
FAST_FLAGS = "-O3 -march=native -ffp-contract=fast -fno-math-errno -fno-trapping-math"
STRICT_FLAGS = "-O2 -ffp-contract=off -fno-fast-math"


def _cpu_tag() -> str:
    """Identifies the host CPU: the -march=native build is redone when the library travels to another machine."""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        model = next((l for l in txt.splitlines() if l.startswith("model name")), "")
        flags = next((l for l in txt.splitlines() if l.startswith("flags")), "")
        return hashlib.sha1((model + flags).encode()).hexdigest()[:16]
    except Exception:
        return "unknown"


def build_fast(force: bool = False) -> str:
    """The TIMING build of the same source (-O3 -march=native, FMA contraction allowed, as an
    optimising Fortran compiler would build the reference).  Never used for parity: the strict build is."""
    os.makedirs(_FAST_DIR, exist_ok=True)
    so = os.path.join(_FAST_DIR, "libmw_oracle_fast.so")
    tag_file = os.path.join(_FAST_DIR, "cpu.tag")
    src = os.path.join(_HERE, "mw_oracle.c")
    tag = _cpu_tag()
    have = open(tag_file).read().strip() if os.path.exists(tag_file) else ""
    stale = (not os.path.exists(so)) or have != tag or os.path.getmtime(src) > os.path.getmtime(so)
    if force or stale:
        cmd = ["gcc"] + FAST_FLAGS.split() + ["-std=gnu99", "-fPIC", "-pthread", "-shared", "-o", so, src, "-lm"]
        subprocess.run(cmd, check=True, capture_output=True)
        open(tag_file, "w").write(tag)
    return so


def lib(fast: bool = False) -> C.CDLL:
    global _lib, _lib_fast
    if fast:
        if _lib_fast is None:
            _lib_fast = _bind(C.CDLL(build_fast()))
        return _lib_fast
    if _lib is not None:
        return _lib
    build()
    _lib = _bind(C.CDLL(_LIB_PATH))
    return _lib


def _bind(L: C.CDLL) -> C.CDLL:
    vp, cp, d, i, i64 = C.c_void_p, C.c_char_p, C.c_double, C.c_int, C.c_int64
    dp = C.POINTER(C.c_double)
    L.orc_const.restype = d; L.orc_const.argtypes = [cp]
    L.orc_create.restype = vp; L.orc_create.argtypes = [i, i]
    L.orc_destroy.argtypes = [vp]
    L.orc_set_config.argtypes = [vp, dp, dp]
    L.orc_determinant.restype = d; L.orc_determinant.argtypes = [dp]
    L.orc_recipmatrix.argtypes = [dp, dp]
    for name in ("orc_energy_init", "orc_mc_water_translation", "orc_mc_volume", "orc_mc_lattice_switch",
                 "orc_mc_update_wl_bins", "orc_mc_monitor", "orc_mc_chain_sync"):
        getattr(L, name).argtypes = [vp]; getattr(L, name).restype = None
    for name in ("orc_compute_ivects", "orc_compute_neighbours", "orc_compute_model_energy"):
        getattr(L, name).argtypes = [vp, i]; getattr(L, name).restype = None
    L.orc_compute_local_real_energy.restype = d; L.orc_compute_local_real_energy.argtypes = [vp, i, i]
    L.orc_params_default.argtypes = [C.POINTER(Params)]
    L.orc_mc_init.restype = i; L.orc_mc_init.argtypes = [vp, C.POINTER(Params), i, i, dp, i, d]
    L.orc_eta_weight.restype = d; L.orc_eta_weight.argtypes = [vp, d]
    L.orc_mu_to_bin.restype = i; L.orc_mu_to_bin.argtypes = [vp, d]
    L.orc_mc_cycle.restype = i; L.orc_mc_cycle.argtypes = [vp]
    L.orc_mc_run.restype = i; L.orc_mc_run.argtypes = [vp, i]
    L.orc_allreduce_bins.argtypes = [C.POINTER(vp), i]
    L.orc_mc_restore.argtypes = [vp, i, d, d, d, i, i, dp, dp, dp, dp, dp, dp]
    L.orc_mc_check_flatness.restype = i
    L.orc_mc_check_flatness.argtypes = [C.POINTER(vp), i, C.POINTER(FlatParams), C.POINTER(FlatReport)]
    L.orc_mc_deltaG_from_hist.restype = d; L.orc_mc_deltaG_from_hist.argtypes = [C.POINTER(vp), i, dp]
    L.orc_join_uhist.argtypes = [C.POINTER(vp), i, i, dp]
    L.orc_join_eta.argtypes = [C.POINTER(vp), i, i, dp]
    L.orc_philox_block.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, dp]
    L.orc_philox_raw.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_mc_run_many.restype = i; L.orc_mc_run_many.argtypes = [C.POINTER(vp), i, i, i]
    L.orc_model_energy_many.argtypes = [C.POINTER(vp), i, i, dp]
    L.orc_max_threads.restype = i
    L.orc_ptr_d.restype = dp; L.orc_ptr_d.argtypes = [vp, cp]
    L.orc_ptr_i.restype = C.POINTER(C.c_int); L.orc_ptr_i.argtypes = [vp, cp]
    L.orc_get_d.restype = d; L.orc_get_d.argtypes = [vp, cp]
    L.orc_get_i.restype = i64; L.orc_get_i.argtypes = [vp, cp]
    L.orc_set_d.argtypes = [vp, cp, d]
    L.orc_set_i.argtypes = [vp, cp, i64]
    L.orc_set_rng_philox.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint64]
    L.orc_set_rng_fifo.argtypes = [vp, dp, i64]
    return L


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def const(name: str) -> float:
    return lib().orc_const(name.encode())


def params_from_user(up) -> Params:
    """Map a mc_water_ls_mw_b200.decks.UserParams onto the oracle's orc_params."""
    p = Params()
    lib().orc_params_default(C.byref(p))
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


class System:
    """One walker (= one MPI rank of the reference)."""

    def __init__(self, nwater: int, nlat: int, fast: bool = False):
        self.L = lib(fast)
        self.nwater, self.nlat = nwater, nlat
        self.h = self.L.orc_create(nwater, nlat)
        self._keep = []

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- configuration ---------------------------------------------------
    def set_config(self, ljr: np.ndarray, hmatrix: np.ndarray) -> None:
        ljr = np.ascontiguousarray(ljr, dtype=np.float64).reshape(self.nlat, self.nwater, 3)
        hm = np.ascontiguousarray(hmatrix, dtype=np.float64).reshape(self.nlat, 9)
        self.L.orc_set_config(self.h, _dp(ljr), _dp(hm))

    def energy_init(self) -> None:
        self.L.orc_energy_init(self.h)

    # -- views -----------------------------------------------------------
    def arr_d(self, name: str, shape) -> np.ndarray:
        p = self.L.orc_ptr_d(self.h, name.encode())
        return np.ctypeslib.as_array(p, shape=tuple(shape))

    def arr_i(self, name: str, shape) -> np.ndarray:
        p = self.L.orc_ptr_i(self.h, name.encode())
        return np.ctypeslib.as_array(p, shape=tuple(shape))

    @property
    def ljr(self): return self.arr_d("ljr", (self.nlat, self.nwater, 3))
    @property
    def ref_ljr(self): return self.arr_d("ref_ljr", (self.nlat, self.nwater, 3))
    @property
    def hmatrix(self): return self.arr_d("h", (2, 9))[: self.nlat]
    @property
    def recip(self): return self.arr_d("recip", (2, 9))[: self.nlat]
    @property
    def volume(self): return self.arr_d("volume", (2,))[: self.nlat]
    @property
    def model_energy(self): return self.arr_d("model_energy", (2,))[: self.nlat]
    @property
    def nivect(self): return self.arr_i("nivect", (2,))[: self.nlat]
    @property
    def ivect(self): return self.arr_d("ivect", (self.nlat, MAXIVECT, 3))
    @property
    def nn(self): return self.arr_i("nn", (self.nlat, self.nwater))
    @property
    def jn(self): return self.arr_i("jn", (self.nlat, self.nwater, MAXNEIGH))
    @property
    def vn(self): return self.arr_i("vn", (self.nlat, self.nwater, MAXNEIGH))
    @property
    def nbins(self): return int(self.geti("nbins"))
    @property
    def weight(self): return self.arr_d("weight", (self.nbins,))
    @property
    def histogram(self): return self.arr_d("histogram", (self.nbins,))
    @property
    def unbiased_hist(self): return self.arr_d("unbiased_hist", (self.nbins,))
    @property
    def mu_bin(self): return self.arr_d("mu_bin", (self.nbins,))
    @property
    def binwidth(self): return self.arr_d("binwidth", (self.nbins,))
    @property
    def mc_translations(self): return self.arr_i("mc_translations", (self.nwater,))

    def getd(self, name: str) -> float: return self.L.orc_get_d(self.h, name.encode())
    def geti(self, name: str) -> int: return self.L.orc_get_i(self.h, name.encode())
    def setd(self, name: str, v: float) -> None: self.L.orc_set_d(self.h, name.encode(), float(v))
    def seti(self, name: str, v: int) -> None: self.L.orc_set_i(self.h, name.encode(), int(v))

    # -- energy module ---------------------------------------------------
    def compute_ivects(self, ils: int) -> None: self.L.orc_compute_ivects(self.h, ils - 1)
    def compute_neighbours(self, ils: int) -> None: self.L.orc_compute_neighbours(self.h, ils - 1)
    def compute_model_energy(self, ils: int) -> float:
        self.L.orc_compute_model_energy(self.h, ils - 1)
        return float(self.model_energy[ils - 1])
    def compute_local_real_energy(self, imol: int, ils: int) -> float:
        return self.L.orc_compute_local_real_energy(self.h, imol - 1, ils - 1)

    # -- mc_moves --------------------------------------------------------
    def mc_init(self, params: Params, rank: int = 0, size: int = 1,
                weights: Optional[np.ndarray] = None, file_wl_factor: float = 0.0) -> int:
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.float64)
            return self.L.orc_mc_init(self.h, C.byref(params), rank, size, _dp(w), len(w), file_wl_factor)
        return self.L.orc_mc_init(self.h, C.byref(params), rank, size, None, 0, 0.0)

    def set_rng_philox(self, seed: int, stream: int, start_index: int = 0) -> None:
        self.L.orc_set_rng_philox(self.h, seed, stream, start_index)

    def set_rng_fifo(self, u: np.ndarray) -> None:
        u = np.ascontiguousarray(u, dtype=np.float64)
        self._keep = [u]
        self.L.orc_set_rng_fifo(self.h, _dp(u), len(u))

    def eta_weight(self, mu: float) -> float: return self.L.orc_eta_weight(self.h, mu)
    def mu_to_bin(self, mu: float) -> int: return self.L.orc_mu_to_bin(self.h, mu)
    def mc_cycle(self) -> int: return self.L.orc_mc_cycle(self.h)
    def mc_run(self, ncycles: int) -> int: return self.L.orc_mc_run(self.h, ncycles)
    def mc_water_translation(self) -> None: self.L.orc_mc_water_translation(self.h)
    def mc_volume(self) -> None: self.L.orc_mc_volume(self.h)
    def mc_lattice_switch(self) -> None: self.L.orc_mc_lattice_switch(self.h)
    def mc_update_wl_bins(self) -> None: self.L.orc_mc_update_wl_bins(self.h)
    def mc_monitor(self) -> None: self.L.orc_mc_monitor(self.h)
    def mc_chain_sync(self) -> None: self.L.orc_mc_chain_sync(self.h)

    def mc_restore(self, rec: dict) -> None:
        """mc_checkpoint_load + restart refresh (mc_moves.F90:403-501, :842-862); rec as decks.read_checkpoint gives it."""
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = [f(rec[k]) for k in ("histogram", "weight", "unbiased_hist", "hmatrix", "ref_ljr", "ljr")]
        self.L.orc_mc_restore(self.h, int(rec["mc_cycle_num"]), float(rec["mc_max_trans"]), float(rec["mc_dv_max"]),
                              float(rec["wl_factor"]), int(rec["wl_invt_active"]), int(rec["ls"]), *[_dp(a) for a in arrs])

    def counters(self) -> dict:
        return {k: int(self.geti(k)) for k in ("acc_r", "acc_v", "acc_s", "att_r", "att_v", "att_s")}


def _handles(walkers: Sequence[System]):
    arr = (C.c_void_p * len(walkers))(*[w.h for w in walkers])
    return arr


def allreduce_bins(walkers: Sequence[System]) -> None:
    walkers[0].L.orc_allreduce_bins(_handles(walkers), len(walkers))


def mc_check_flatness(walkers: Sequence[System], wl_schedule: int = 0, wl_minhist: int = 20,
                      wl_flattol: float = float(np.float32(0.05)), wl_useinvt: bool = False) -> "FlatReport":
    """mc_check_flatness (mc_moves.F90:1936-2185) over in-process walkers (= MPI ranks)."""
    fp = FlatParams(wl_schedule, wl_minhist, wl_flattol, int(wl_useinvt))
    rep = FlatReport()
    rc = lib().orc_mc_check_flatness(_handles(walkers), len(walkers), C.byref(fp), C.byref(rep))
    if rc:
        raise RuntimeError(f"oracle: reference stop {rc} in mc_check_flatness")
    return rep


def mc_deltaG_from_hist(walkers: Sequence[System]):
    """mc_compute_deltaG_from_hist (mc_moves.F90:2498-2621): (deltaG in kT, normP)."""
    normP = np.zeros(walkers[0].nbins, dtype=np.float64)
    dG = lib().orc_mc_deltaG_from_hist(_handles(walkers), len(walkers), _dp(normP))
    return float(dG), normP


def join_uhist(walkers: Sequence[System], overlap: int) -> np.ndarray:
    out = np.zeros(walkers[0].nbins, dtype=np.float64)
    lib().orc_join_uhist(_handles(walkers), len(walkers), overlap, _dp(out))
    return out


def join_eta(walkers: Sequence[System], overlap: int) -> np.ndarray:
    out = np.zeros(walkers[0].nbins, dtype=np.float64)
    lib().orc_join_eta(_handles(walkers), len(walkers), overlap, _dp(out))
    return out


def mc_run_many(walkers: Sequence[System], ncycles: int, nthreads: int = 0) -> int:
    return walkers[0].L.orc_mc_run_many(_handles(walkers), len(walkers), ncycles, nthreads)


def model_energy_many(walkers: Sequence[System], nthreads: int = 0) -> np.ndarray:
    out = np.zeros((len(walkers), 2), dtype=np.float64)
    walkers[0].L.orc_model_energy_many(_handles(walkers), len(walkers), nthreads, _dp(out))
    return out


def max_threads() -> int:
    return lib().orc_max_threads()


def philox_block(seed: int, stream: int, block: int) -> np.ndarray:
    out = np.zeros(2, dtype=np.float64)
    lib().orc_philox_block(seed, stream, block, _dp(out))
    return out


def philox_raw(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox_raw(c, k, o)
    return [int(x) for x in o]
