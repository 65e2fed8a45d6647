"""Host-side output formats that carry hot-path state: Fortran edit descriptors,
therm rows (main.f90:200-223), checkpoint records (mc_moves.F90:324-501).  No GPU."""
import os
import struct

import numpy as np

from mc_water_ls_mw_b200 import decks
from tests.helpers import example_dir, load_example


class _Row:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def test_fortran_e_matches_the_reference_own_output():
    # the header of examples/ice1_sample/eta_weights.dat was written by the reference with
    # '("#Current energy increment = ",E20.12)' (mc_moves.F90:1829): 28 characters of text + the E20.12 field
    line = open(os.path.join(example_dir("ice1_sample"), "eta_weights.dat")).readline().rstrip("\n")
    assert len(line) == 48
    wl = float(line[28:48])
    assert line[28:48] == decks.fortran_e(wl, 20, 12)
    assert decks.fortran_e(1.0, 15, 6) == "   0.100000E+01"
    assert decks.fortran_e(-123456.789, 15, 6) == "  -0.123457E+06"
    assert decks.fortran_e(0.0, 15, 6) == "   0.000000E+00"
    assert decks.fortran_e(9.9999996e5, 15, 6) == "   0.100000E+07"        # rounding carries into the exponent
    assert decks.fortran_f(-0.5, 15, 6) == "      -0.500000"


def test_therm_row_formats():
    up2, *_ = load_example("ice1_sample")
    up1, h1, *_ = load_example("single_box")
    row = _Row(icyc=125, ls=2, model_energy=[-0.7, -0.71], ls_mu=-12.3456789, volume=[9000.0, 9100.0], hmatrix1=list(h1[0]))
    s = decks.format_therm_row(row, up2)                                   # sample run: '(I8,E15.6,5x,3F15.6,1x,I1)'
    assert len(s) == 8 + 15 + 5 + 45 + 2 and s.endswith(" 2")
    assert s[:8] == "     125" and s[8:23] == decks.fortran_e(-0.71 * decks.HART_TO_EV, 15, 6)
    assert float(s[28:43]) == -12.345679 and abs(float(s[58:73]) - 9100.0 * decks.BOHR_TO_ANG ** 3) < 1e-6
    upg, *_ = load_example("ice1_gen_weights")
    s = decks.format_therm_row(row, upg)                                   # weight generation: density column
    assert len(s) == 8 + 15 + 5 + 30 + 2
    dens = upg.nwater * decks.WATER_MASS / 9100.0 * decks.AUD_TO_KGM3
    assert abs(float(s[43:58]) - dens) < 1e-6 and 800 < dens < 1100       # kg/m^3 of ice
    s = decks.format_therm_row(row, up1)                                   # single box: '(I8,E15.6,5x,F15.6,6F15.6)'
    assert len(s) == 8 + 15 + 5 + 15 * 7
    la, lb, lc, al, be, ga = decks.hmatrix_to_abc(h1[0])
    assert abs(float(s[43:58]) - la * decks.BOHR_TO_ANG) < 1e-6 and abs(float(s[-15:]) - ga) < 1e-6


def test_checkpoint_roundtrip_and_record_structure(tmp_path):
    rng = np.random.default_rng(1)
    nb, nl, n = 101, 2, 48
    rec = dict(nwater=n, mc_cycle_num=1234, mc_max_trans=0.37, mc_dv_max=0.011, wl_factor=0.0025,
               histogram=rng.random(nb), weight=rng.random(nb), wl_invt_active=True, unbiased_hist=rng.random(nb),
               hmatrix=rng.random((nl, 9)), ref_ljr=rng.random((nl, n, 3)), ljr=rng.random((nl, n, 3)), ls=2)
    for samplerun in (True, False):
        p = str(tmp_path / f"checkpoint000.dat.{int(samplerun) + 1}")
        decks.write_checkpoint(p, rec, nb, samplerun)
        raw = open(p, "rb").read()
        # sequential unformatted: every record is framed by its byte length; first two records = nwater, cycle
        assert struct.unpack_from("<iii", raw, 0) == (4, n, 4) and struct.unpack_from("<iii", raw, 12) == (4, 1234, 4)
        nrec = 0; pos = 0
        while pos < len(raw):
            (m,) = struct.unpack_from("<i", raw, pos); pos += 8 + m; nrec += 1
        assert pos == len(raw) and nrec == (12 if samplerun else 11)          # mc_moves.F90:353-384
        back = decks.read_checkpoint(p, nb, nl, samplerun)
        for k, v in rec.items():
            if k == "unbiased_hist" and not samplerun:
                continue
            assert np.array_equal(np.asarray(back[k]), np.asarray(v)), k
