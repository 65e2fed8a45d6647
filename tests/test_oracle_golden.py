"""CPU tests: pin the oracle against every golden datum the reference ships
and against its own committed vectors; host-side deck handling."""
import os

import numpy as np
import pytest

from mc_water_ls_mw_b200 import decks
from oracle import orc
from tests.helpers import GOLDEN, example_dir, load_example, make_oracle_walker, make_oracle_walkers


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox_raw([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox_raw([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox_raw([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    d = orc.philox_block(0, 0, 0)
    assert d[0] == (((0xE169C58D << 32) | 0x6627E8D5) >> 11) * 2.0 ** -53
    assert d[1] == (((0x9B00DBD8 << 32) | 0xBC57AC4C) >> 11) * 2.0 ** -53
    assert 0.0 <= d[0] < 1.0


def test_literal_promotion_and_constants():
    # molint.F90:74: cos0 is a single-precision literal; userparams.f90:32 likewise for wl_factor,
    # the latter proven by the header of examples/ice1_sample/eta_weights.dat
    assert orc.const("cos0") == float(np.float32(-0.33331324756))
    assert orc.const("cos0").hex() == "-0x1.5550120000000p-2"
    wl, _, _ = decks.read_eta_weights(os.path.join(example_dir("ice1_sample"), "eta_weights.dat"))
    assert abs(wl - float(np.float32(0.05))) < 1e-13
    assert orc.const("wl_factor_default") == float(np.float32(0.05))
    assert orc.const("ang_to_bohr") == 1.0 / 0.5291772108
    assert decks.ANG_TO_BOHR == orc.const("ang_to_bohr") and decks.KB == orc.const("kb")


def test_bin_grid_matches_reference_golden():
    # column 1 of eta_weights.dat was written by the reference itself (mc_moves.F90:1840)
    s, up = make_oracle_walker("ice1_sample")
    _, mu, _ = decks.read_eta_weights(os.path.join(example_dir("ice1_sample"), "eta_weights.dat"))
    assert len(mu) == 101 == s.nbins
    np.testing.assert_allclose(s.mu_bin, mu, rtol=1e-14, atol=0)
    assert abs(s.getd("r_pos") - 1.0694879976881435) < 1e-14
    assert abs(s.getd("av_binwidth") - 7.920792079208) < 1e-11
    assert abs(s.binwidth.sum() - 800.0) < 1e-9
    # mu_to_bin inverts the grid
    for k in range(101):
        assert s.mu_to_bin(float(s.mu_bin[k])) == k + 1
    # eta at the bin centres is the tabulated weight (interpolation anchors)
    for k in range(1, 100):
        assert abs(s.eta_weight(float(s.mu_bin[k])) - s.weight[k]) < 1e-12


def test_neighbour_list_known_answer():
    # molint.F90:79 "only expect 16/17 entries"
    s, _ = make_oracle_walker("ice1_sample")
    assert set(s.nn[0].tolist()) == {16} and set(s.nn[1].tolist()) == {17}
    assert s.geti("nn_warnings") == 0
    assert list(s.nivect) == [27, 27]
    # symmetry: j in list(i) with image k  <=>  i in list(j) with the inverse image
    for l in range(2):
        pairs = {(i, int(s.jn[l, i, t]) - 1, tuple(np.round(s.ivect[l, s.vn[l, i, t] - 1], 9)))
                 for i in range(48) for t in range(s.nn[l, i])}
        for (i, j, v) in pairs:
            assert (j, i, tuple(-x + 0.0 for x in v)) in pairs


def test_energies_survey_values():
    s, _ = make_oracle_walker("ice1_sample")
    # survey-time numpy restatement ("provisional"), independent of this C code
    assert abs(s.model_energy[0] - (-0.9391903897229951)) < 1e-13
    assert abs(s.model_energy[1] - (-0.9403038656656721)) < 1e-13
    assert abs(s.getd("ls_mu") - 1.645457170) < 1e-8


def test_delta_full_equals_delta_local_invariant():
    # the reference's own DEBUG check, threshold 1d-10 Ha (mc_moves.F90:1094-1102)
    s, _ = make_oracle_walker("ice1_sample")
    rng = np.random.default_rng(7)
    for _ in range(25):
        i = int(rng.integers(48)); l = int(rng.integers(2))
        e_full0 = s.compute_model_energy(l + 1)
        e_loc0 = s.compute_local_real_energy(i + 1, l + 1)
        s.ljr[l, i] += rng.uniform(-0.6, 0.6, 3)
        e_full1 = s.compute_model_energy(l + 1)
        e_loc1 = s.compute_local_real_energy(i + 1, l + 1)
        assert abs((e_full1 - e_full0) - (e_loc1 - e_loc0)) < 1e-10


def test_energy_drift_check():
    # "Checking accumulated energies" (mc_moves.F90:1781-1792)
    s, _ = make_oracle_walker("ice1_sample")
    s.set_rng_philox(20141211, 0, 1000000)
    assert s.mc_run(100) == 0
    stored = np.array(s.model_energy)
    fresh = np.array([s.compute_model_energy(1), s.compute_model_energy(2)])
    assert np.max(np.abs(stored - fresh)) < 1e-10


def test_oracle_vectors_regression():
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    for ex in ("ice1_sample", "single_box", "ice1_gen_weights"):
        s, _ = make_oracle_walker(ex)
        np.testing.assert_array_equal(s.nn, v[f"{ex}/nn"])
        np.testing.assert_array_equal(s.jn, v[f"{ex}/jn"])
        np.testing.assert_array_equal(s.vn, v[f"{ex}/vn"])
        np.testing.assert_allclose(s.model_energy, v[f"{ex}/energy0"], rtol=1e-13)
        s.set_rng_philox(20141211, 0, 1000000)
        assert s.mc_run(30) == 0
        np.testing.assert_array_equal(s.ljr, v[f"{ex}/ljr30"])
        np.testing.assert_array_equal(s.hmatrix, v[f"{ex}/h30"])
        np.testing.assert_allclose(s.model_energy, v[f"{ex}/energy30"], rtol=1e-12)
        got = [s.geti(k) for k in ("acc_r", "acc_v", "acc_s", "att_r", "att_v", "att_s", "ls")]
        assert got == v[f"{ex}/counters30"].tolist()
        assert s.geti("rng_index") == int(v[f"{ex}/rng_index30"][0])


def test_deck_reader_quirks():
    up = decks.read_input(os.path.join(example_dir("ice1_sample"), "ice.input"))
    assert up.nwater == 48 and up.num_lattices == 2 and up.samplerun and up.mc_always_switch
    assert up.mc_vol_prob == 1.0 / 768.0            # io.f90:172 runs before nwater is read
    assert up.mc_switch_prob == 0.1                 # zeroed later by mc_always_switch (mc_moves.F90:158)
    assert up.mc_max_trans == 1.1 * decks.ANG_TO_BOHR and up.mc_dv_max == 0.924 * decks.ANG_TO_BOHR
    assert up.pressure == 1.0 / decks.AUP_TO_ATM and up.temperature == 200.0
    assert up.nbins == 101 and up.mu_max == 400.0 and up.list_update_int == 10
    assert up.window_overlap == 0                   # size == 1 (io.f90:249)
    sb = decks.read_input(os.path.join(example_dir("single_box"), "ice.input"))
    assert sb.num_lattices == 1 and not sb.allow_switch and not sb.mc_always_switch and sb.mc_switch_prob == 0.0
    assert sb.nbins == 201 and sb.wl_factor == float(np.float32(0.05))
    dd = decks.read_input(os.path.join(example_dir("ice1_sample_dd"), "ice.input"), size=4)
    assert dd.parallel_strategy == "dd" and dd.window_overlap == 2
    s, _ = make_oracle_walker("ice1_sample")
    s.mc_cycle()
    assert abs(s.getd("transP") - 0.5 / (0.5 + 1.0 / 768.0)) < 1e-15 and abs(s.getd("volP") - 1.0) < 2e-16


def test_xmol_reader():
    up, h, r, w, wl = load_example("ice1_sample")
    assert h.shape == (2, 9) and r.shape == (2, 48, 3) and len(w) == 101
    assert abs(h[0, 0] - 13.352018 * decks.ANG_TO_BOHR) < 1e-12
    assert abs(r[0, 0, 0] - 1.183841 * decks.ANG_TO_BOHR) < 1e-12
    vol = abs(np.linalg.det(h[0].reshape(3, 3))) * decks.BOHR_TO_ANG ** 3
    assert abs(vol - 1504.43) < 0.01


def test_dd_windows_cover_range():
    ws = make_oracle_walkers("ice1_sample_dd", 4)
    lo = [w.getd("my_mu_min") for w in ws]; hi = [w.getd("my_mu_max") for w in ws]
    assert lo[0] == -400.0 and hi[-1] == 400.0
    for a in range(3):
        assert lo[a + 1] < hi[a]                     # overlapping windows
    assert [w.geti("my_start_bin") for w in ws] == [1, 23, 48, 73]
    assert [w.geti("my_end_bin") for w in ws] == [27, 52, 77, 101]
    assert ws[0].geti("ls") == 1 and ws[3].geti("ls") == 2     # mc_moves.F90:702-703


def test_oracle_allreduce_semantics():
    # comms_mpi.f90:256-270: weight_after = base + sum_over_walkers(weight - base)
    ov = {"eq_mc_cycles": 1, "samplerun": False, "wl_factor": 0.005}
    ws = make_oracle_walkers("ice1_gen_weights", 3, overrides=ov)
    for i, w in enumerate(ws):
        w.set_rng_philox(20141211, i, 1000000)
        assert w.mc_run(5) == 0
    deltas_w = sum(np.array(w.weight) - np.array(w.arr_d("eta_last_sync", (101,))) for w in ws)
    deltas_h = sum(np.array(w.histogram) for w in ws)
    base = np.array(ws[0].arr_d("eta_last_sync", (101,)))
    orc.allreduce_bins(ws)
    for w in ws:
        np.testing.assert_allclose(w.weight, base + deltas_w, rtol=0, atol=1e-12)
        np.testing.assert_allclose(w.histogram, deltas_h, rtol=0, atol=1e-12)
    assert deltas_h.sum() > 0
