"""Readers for the reference's UNCHANGED input files (host side of the drop-in).

The hot path keeps the reference's program entry, input decks and output
formats (BASELINE.json north_star), so the host driver has to understand:

* the namelist input deck        -- io.f90:58-327  (``io_read_input``)
* ``inputNNN.xmol``              -- init.f90:38-125 (``read_xmol``)
* ``eta_weights.dat``            -- mc_moves.F90:748-768 (read), :1827-1840 (write)

Everything is converted to the reference's internal units (Bohr, Hartree,
atomic-unit pressure) exactly where the reference converts it, including the
``mc_vol_prob = 1/768`` quirk (io.f90:172 runs before ``nwater`` is read at
io.f90:191, so the default ``nwater = 768`` of userparams.f90:17 is used).
"""
from __future__ import annotations

import dataclasses
import os
import re
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

# constants.f90:23-24,39,43,59
PI = 3.141592653589793238462643383279502884197
INV_PI = 1.0 / 3.141592653589793238462643383279502884197
KB = 1.0 / 3.1577465e5
BOHR_TO_ANG = 0.5291772108
ANG_TO_BOHR = 1.0 / 0.5291772108
HART_TO_EV = 27.211396181
AUP_TO_ATM = 2.90363081e8
WATER_MASS = 18.0158
AUD_TO_KGM3 = 1.120587168e4


def f32(x: float) -> float:
    """A Fortran default-real literal promoted to double (userparams.f90:32, molint.F90:74)."""
    return struct.unpack("f", struct.pack("f", x))[0]


@dataclasses.dataclass
class UserParams:
    """userparams.f90:14-79 -- names and defaults are the reference's."""

    model_type: str = "mW"
    nwater: int = 768
    num_lattices: int = 2
    method: str = "xmol"
    r_overlap: float = 1.7 * ANG_TO_BOHR
    pressure: float = 1.0 / AUP_TO_ATM
    temperature: float = 240.0
    mc_ensemble: str = "npt"
    mc_max_trans: float = 0.6
    mc_target_ratio: float = 0.50
    mc_dv_max: float = 0.1
    wl_factor: float = f32(0.05)
    wl_schedule: int = 0
    wl_minhist: int = 20
    wl_flattol: float = f32(0.05)
    wl_useinvt: bool = False
    wl_swetnam: bool = False
    wl_alpha: float = 1.0
    eta_interp: bool = True
    samplerun: bool = False
    leshift: bool = False
    nbins: int = 201
    mu_min: float = -8000.0
    mu_max: float = 8000.0
    allow_switch: bool = True
    allow_vol: bool = True
    allow_trans: bool = True
    mc_trans_prob: float = 0.5
    mc_vol_prob: float = 0.01
    mc_switch_prob: float = 0.0
    mc_always_switch: bool = True
    input_ref_enthalpy: Tuple[float, float] = (0.0, 0.0)
    list_update_int: int = 50
    traj_output_int: int = 5000000
    file_output_int: int = 5
    latt_sync_int: int = 10000
    mpi_sync_int: int = 250
    chkpt_dump_int: int = 1000
    monitor_int: int = 1000
    flat_chk_int: int = 10000
    invt_dump_int: int = 500000
    eq_adjust_mc: bool = False
    deltaG_int: int = 100000
    max_mc_cycles: int = 1000
    eq_mc_cycles: int = 25000
    parallel_strategy: str = "mw"
    window_overlap: int = 2
    timer_qtime: float = 0.0
    timer_closetime: float = 0.0
    ls: int = 1  # model.ls (data_structures.f90:51), read in &config
    seedname: str = ""


_GROUPS = ("potential", "thermal", "montecarlo", "config", "bookkeeping", "parallelisation")


def _parse_value(text: str):
    t = text.strip().rstrip(",").strip()
    if not t:
        raise ValueError("empty namelist value")
    if t[0] in "'\"":
        return t[1:t.index(t[0], 1)]
    low = t.lower()
    if low in (".true.", "t", ".t."):
        return True
    if low in (".false.", "f", ".f."):
        return False
    parts = [p for p in re.split(r"[,\s]+", t) if p]
    vals = []
    for p in parts:
        q = p.lower().replace("d", "e")
        vals.append(int(q) if re.fullmatch(r"[+-]?\d+", q) else float(q))
    return vals[0] if len(vals) == 1 else tuple(vals)


def parse_namelists(text: str) -> Dict[str, Dict[str, object]]:
    """Minimal Fortran-namelist reader: ``&group key = value ... /`` with ``!`` comments."""
    groups: Dict[str, Dict[str, object]] = {}
    current: Optional[str] = None
    for raw in text.splitlines():
        # strip comments (quotes in these decks never contain '!')
        line = raw.split("!", 1)[0].strip()
        if not line:
            continue
        if line.startswith("&"):
            current = line[1:].split()[0].lower()
            groups[current] = {}
            line = line[1 + len(current):].strip()
            if not line:
                continue
        if line == "/" or line.lower() == "&end":
            current = None
            continue
        if current is None:
            continue
        if line.endswith("/"):
            line, closing = line[:-1], True
        else:
            closing = False
        for m in re.finditer(r"(\w+(?:\(\d+\))?)\s*=\s*([^=]+?)(?=(?:,?\s*\w+(?:\(\d+\))?\s*=)|$)", line):
            groups[current][m.group(1).lower()] = _parse_value(m.group(2))
        if closing:
            current = None
    return groups


def read_input(path: str, size: int = 1) -> UserParams:
    """io.f90:58-327.  ``size`` is the number of walkers/ranks (io.f90:249)."""
    with open(path, "r") as fh:
        groups = parse_namelists(fh.read())
    p = UserParams()
    base = os.path.basename(path)
    p.seedname = base[: base.rindex(".")] if "." in base else base  # io.f90:127-132

    for g in ("potential", "thermal", "montecarlo", "config", "bookkeeping"):
        if g not in groups:
            raise ValueError(f"Error reading {g} namelist")  # io.f90:152-229 `stop`

    def apply(group: str) -> None:
        for k, v in groups.get(group, {}).items():
            if k == "input_ref_enthalpy":
                p.input_ref_enthalpy = tuple(float(x) for x in v)  # type: ignore[arg-type]
                continue
            if not hasattr(p, k):
                raise ValueError(f"Error reading {group} namelist: unknown variable {k}")
            cur = getattr(p, k)
            if isinstance(cur, bool):
                setattr(p, k, bool(v))
            elif isinstance(cur, int):
                setattr(p, k, int(v))  # type: ignore[arg-type]
            elif isinstance(cur, float):
                setattr(p, k, float(v))  # type: ignore[arg-type]
            else:
                setattr(p, k, v)

    apply("potential")
    apply("thermal")
    if p.temperature < 0.0:
        raise ValueError("Error temperature must be positive")
    p.pressure = p.pressure / AUP_TO_ATM  # io.f90:165

    p.mc_switch_prob = 0.1  # io.f90:171
    p.mc_vol_prob = 1.0 / float(p.nwater)  # io.f90:172 -- nwater is still the default 768 here
    apply("montecarlo")
    if p.mc_ensemble not in ("nvt", "npt"):
        raise ValueError("Error - unrecognised ensemble.")
    p.mc_max_trans = p.mc_max_trans * ANG_TO_BOHR  # io.f90:185-186
    p.mc_dv_max = p.mc_dv_max * ANG_TO_BOHR

    apply("config")
    if p.nwater < 1:
        raise ValueError("Error - invalid number of waters")
    if p.r_overlap < 0.0:
        raise ValueError("Error - invalid r_overlap")
    if p.method.strip() != "xmol":
        raise ValueError("Invalid initialisation option in config namelist")
    p.r_overlap = p.r_overlap * ANG_TO_BOHR
    if p.num_lattices == 1:  # io.f90:208-214
        p.allow_switch = False
        p.mc_switch_prob = 0.0
        p.mc_always_switch = False
        p.ls = 1
    elif p.num_lattices != 2:
        raise ValueError("Error num_lattices must equal 1 or 2!")

    apply("bookkeeping")
    for name in ("list_update_int", "traj_output_int", "file_output_int", "max_mc_cycles", "eq_mc_cycles"):
        if getattr(p, name) < 1:
            raise ValueError(f"Error - {name} must be > 0")
    apply("parallelisation")
    if size == 1:
        p.window_overlap = 0  # io.f90:249
    if p.parallel_strategy not in ("mw", "dd"):
        raise ValueError("Unknown parallel strategy")  # mc_moves.F90:721
    return p


def read_xmol(path: str, nwater: int) -> Tuple[np.ndarray, np.ndarray]:
    """init.f90:75-106.  Returns (hmatrix[3,3] column-major as h[col,row], pos[nwater,3]) in Bohr.

    ``hmatrix`` is returned as a flat Fortran-ordered array of 9 numbers
    (h(1,1),h(2,1),h(3,1),h(1,2),...), i.e. the order of the file's second line.
    """
    with open(path, "r") as fh:
        lines = [ln for ln in fh.read().splitlines()]
    n = int(lines[0].split()[0])
    if n != nwater:
        raise ValueError("Error wrong number of atoms in input.xmol")
    h = np.array([float(x) for x in lines[1].split()[:9]], dtype=np.float64)
    pos = np.empty((nwater, 3), dtype=np.float64)
    for i in range(nwater):
        f = lines[2 + i].split()
        pos[i] = [float(f[1]), float(f[2]), float(f[3])]
    return h * ANG_TO_BOHR, pos * ANG_TO_BOHR


def read_config(directory: str, p: UserParams) -> Tuple[np.ndarray, np.ndarray]:
    """All lattices of a run directory: (hmatrix[nlat,9], ljr[nlat,nwater,3]) in Bohr."""
    hs: List[np.ndarray] = []
    rs: List[np.ndarray] = []
    for ils in range(1, p.num_lattices + 1):
        h, r = read_xmol(os.path.join(directory, f"input{ils:03d}.xmol"), p.nwater)
        hs.append(h)
        rs.append(r)
    return np.ascontiguousarray(np.stack(hs)), np.ascontiguousarray(np.stack(rs))


def read_eta_weights(path: str) -> Tuple[float, np.ndarray, np.ndarray]:
    """mc_moves.F90:751-766: header ``'(A29,E20.12)'`` then ``mu_bin weight`` rows."""
    with open(path, "r") as fh:
        lines = fh.read().splitlines()
    wl = float(lines[0][29:49].strip().lower().replace("d", "e"))
    mu, w = [], []
    for ln in lines[1:]:
        f = ln.split()
        if len(f) >= 2:
            mu.append(float(f[0]))
            w.append(float(f[1]))
    return wl, np.array(mu), np.array(w)


def write_eta_weights(path: str, wl_factor: float, mu_bin: np.ndarray, weight: np.ndarray) -> None:
    """mc_moves.F90:1829,1840: header '("#Current energy increment = ",E20.12)' + list-directed rows."""
    mant, exp = f"{wl_factor:.11E}".split("E")
    # Fortran E20.12 prints 0.dddddddddddd E+xx (leading zero form)
    val = float(wl_factor)
    if val == 0.0:
        s = "0.000000000000E+00"
    else:
        e = int(np.floor(np.log10(abs(val)))) + 1
        m = val / 10.0 ** e
        s = f"{m:.12f}E{e:+03d}"
    with open(path, "w") as fh:
        fh.write("#Current energy increment = " + s.rjust(20) + "\n")
        for a, b in zip(mu_bin, weight):
            fh.write(f"  {a: .17g}       {b: .17g}     \n")


# --------------------------------------------------------------------------------------------------
# output formats of the host driver that depend on hot-path state
# --------------------------------------------------------------------------------------------------
HART_TO_EV = 27.211396181          # constants.f90:47
WATER_MASS = 18.0158               # constants.f90:52
AUD_TO_KGM3 = 1.120587168e4        # constants.f90:63


def fortran_e(x: float, w: int, d: int) -> str:
    """Fortran ``Ew.d`` edit descriptor (0.ddddddE+xx form, as gfortran / ifort print it)."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return ("NaN" if x != x else ("Infinity" if x > 0 else "-Infinity")).rjust(w)
    if x == 0.0:
        body = "0." + "0" * d + "E+00"
    else:
        m, e = f"{abs(x):.{d - 1}E}".split("E")       # d significant digits: D.ddd E exp
        e = int(e) + 1
        digits = m.replace(".", "")
        body = "0." + digits + "E" + ("+" if e >= 0 else "-") + f"{abs(e):02d}"
    s = ("-" if x < 0 else "") + body
    if len(s) > w and s.startswith("0."):
        s = s[1:]
    elif len(s) > w and s.startswith("-0."):
        s = "-" + s[2:]
    return s.rjust(w) if len(s) <= w else "*" * w


def fortran_f(x: float, w: int, d: int) -> str:
    s = f"{float(x):.{d}f}"
    return s.rjust(w) if len(s) <= w else "*" * w


def hmatrix_to_abc(h: np.ndarray):
    """util_hmatrix_to_abc (util.f90:79-106): h = hmatrix(:,:,1) flattened column-major (Bohr)."""
    a, b, c = np.asarray(h[0:3], float), np.asarray(h[3:6], float), np.asarray(h[6:9], float)
    la, lb, lc = np.sqrt(a @ a), np.sqrt(b @ b), np.sqrt(c @ c)
    alpha = np.degrees(np.arccos((a @ c) / (la * lc)))
    beta = np.degrees(np.arccos((b @ c) / (lb * lc)))
    gamma = np.degrees(np.arccos((a @ b) / (la * lb)))
    return la, lb, lc, alpha, beta, gamma


def format_therm_row(row, p: UserParams) -> str:
    """One line of ``<seed>RRR_therm.dat`` (main.f90:200-223) from the values the walker kernel recorded
    (``mwgpu_therm_row`` / any object with icyc, ls, model_energy, ls_mu, volume, hmatrix1)."""
    icyc, ls = int(row.icyc), int(row.ls)
    E = [float(row.model_energy[0]), float(row.model_energy[1])]
    V = [float(row.volume[0]), float(row.volume[1])]
    b3 = BOHR_TO_ANG ** 3
    if p.num_lattices == 1:                                         # '(I8,E15.6,5x,F15.6,6F15.6)'
        la, lb, lc, al, be, ga = hmatrix_to_abc(np.array(list(row.hmatrix1)))
        return (f"{icyc:8d}" + fortran_e(E[0] * HART_TO_EV, 15, 6) + " " * 5 + fortran_f(V[0] * b3, 15, 6)
                + "".join(fortran_f(v, 15, 6) for v in (la * BOHR_TO_ANG, lb * BOHR_TO_ANG, lc * BOHR_TO_ANG, al, be, ga)))
    head = f"{icyc:8d}" + fortran_e(E[ls - 1] * HART_TO_EV, 15, 6) + " " * 5
    if p.wl_factor < np.finfo(np.float64).tiny or p.samplerun:       # '(I8,E15.6,5x,3F15.6,1x,I1)'
        return head + fortran_f(float(row.ls_mu), 15, 6) + fortran_f(V[0] * b3, 15, 6) + fortran_f(V[1] * b3, 15, 6) + f" {ls:1d}"
    density = p.nwater * WATER_MASS / V[ls - 1]                      # '(I8,E15.6,5x,2F15.6,1x,I1)'
    return head + fortran_f(float(row.ls_mu), 15, 6) + fortran_f(density * AUD_TO_KGM3, 15, 6) + f" {ls:1d}"


def _rec(payload: bytes) -> bytes:
    """One record of a Fortran sequential unformatted file (4-byte little-endian length markers: the
    gfortran / ifort default; the reference leaves the record format to the compiler)."""
    n = struct.pack("<i", len(payload))
    return n + payload + n


def write_checkpoint(path: str, rec: dict, nbins: int, samplerun: bool) -> None:
    """``checkpointRRR.dat.N`` exactly as mc_checkpoint_write (mc_moves.F90:324-388) writes it: records
    nwater | mc_cycle_num | mc_max_trans,mc_dv_max | wl_factor | histogram | weight | wl_invt_active |
    [unbiased_hist] | hmatrix | ref_ljr | ljr | ls.  Arrays are in the reference's column-major layouts, which is
    what ``WalkerBatch.checkpoint_record`` returns (hmatrix[nlat,9], ljr[nlat,nwater,3] C-order == Fortran
    (3,3,nlat) / (3,1,nwater,nlat))."""
    f8 = lambda a: np.ascontiguousarray(a, dtype="<f8").tobytes()
    out = [_rec(struct.pack("<i", int(rec["nwater"]))), _rec(struct.pack("<i", int(rec["mc_cycle_num"]))),
           _rec(struct.pack("<dd", float(rec["mc_max_trans"]), float(rec["mc_dv_max"]))),
           _rec(struct.pack("<d", float(rec["wl_factor"]))),
           _rec(f8(np.asarray(rec["histogram"])[:nbins])), _rec(f8(np.asarray(rec["weight"])[:nbins])),
           _rec(struct.pack("<i", 1 if rec["wl_invt_active"] else 0))]          # default LOGICAL: 4 bytes
    if samplerun:
        out.append(_rec(f8(np.asarray(rec["unbiased_hist"])[:nbins])))
    out += [_rec(f8(rec["hmatrix"])), _rec(f8(rec["ref_ljr"])), _rec(f8(rec["ljr"])), _rec(struct.pack("<i", int(rec["ls"])))]
    with open(path, "wb") as fh:
        fh.write(b"".join(out))


def read_checkpoint(path: str, nbins: int, num_lattices: int, samplerun: bool) -> dict:
    """Inverse of write_checkpoint = what mc_checkpoint_load (mc_moves.F90:390-501) reads."""
    data = open(path, "rb").read()
    pos = 0

    def rec():
        nonlocal pos
        (n,) = struct.unpack_from("<i", data, pos)
        payload = data[pos + 4: pos + 4 + n]
        (m,) = struct.unpack_from("<i", data, pos + 4 + n)
        if m != n:
            raise ValueError("corrupt checkpoint record")
        pos += 8 + n
        return payload

    out = {}
    (out["nwater"],) = struct.unpack("<i", rec())
    (out["mc_cycle_num"],) = struct.unpack("<i", rec())
    out["mc_max_trans"], out["mc_dv_max"] = struct.unpack("<dd", rec())
    (out["wl_factor"],) = struct.unpack("<d", rec())
    out["histogram"] = np.frombuffer(rec(), dtype="<f8").copy()
    out["weight"] = np.frombuffer(rec(), dtype="<f8").copy()
    out["wl_invt_active"] = bool(struct.unpack("<i", rec())[0])
    out["unbiased_hist"] = np.frombuffer(rec(), dtype="<f8").copy() if samplerun else np.zeros(nbins)
    n = out["nwater"]
    out["hmatrix"] = np.frombuffer(rec(), dtype="<f8").reshape(num_lattices, 9).copy()
    out["ref_ljr"] = np.frombuffer(rec(), dtype="<f8").reshape(num_lattices, n, 3).copy()
    out["ljr"] = np.frombuffer(rec(), dtype="<f8").reshape(num_lattices, n, 3).copy()
    out["ls"] = struct.unpack("<i", rec())[0] if pos < len(data) else 1         # 'read(...,end=10) ls'
    return out
