"""The C oracle against a SECOND, independent restatement of the reference (tests/golden/ref_numpy.py: plain
Python / numpy written directly from molint.F90 and mc_moves.F90, vectors committed in tests/golden/numpy_vectors.npz
by tests/golden/make_fixtures_numpy.py).  The reference ships no energies or accept / reject counts and cannot be
compiled here; two restatements by different routes that agree -- lists, positions, counters, random-number
consumption bit for bit, energies to 1e-12 -- are what pins the oracle as far as this image allows.  (Unpinnable by
construction: the reference's own random stream and its libm / compiler)"""
import os

import numpy as np
import pytest

from oracle import orc
from tests.helpers import GOLDEN, load_example, used_lists

V = np.load(os.path.join(GOLDEN, "numpy_vectors.npz"))
CASES = {
    # key: (deck, overrides, rank, size) -- tests/golden/make_fixtures_numpy.py
    "ice1_sample": ("ice1_sample", {"eq_mc_cycles": 1, "mc_vol_prob": 0.04, "list_update_int": 2}, 0, 1),
    "single_box": ("single_box", {"eq_mc_cycles": 1, "mc_vol_prob": 0.04, "list_update_int": 2}, 0, 1),
    "ice1_gen_weights": ("ice1_gen_weights", {"eq_mc_cycles": 1, "list_update_int": 2}, 0, 1),
    # single windows of the domain-decomposed decks (mc_moves.F90:660-703, :181-208, :237, :244, :808-812, :1682-1685)
    "ice1_sample_dd@0of4": ("ice1_sample_dd", {"eq_mc_cycles": 100, "mc_vol_prob": 0.04, "list_update_int": 2}, 0, 4),
    "ice1_sample_dd@2of4": ("ice1_sample_dd", {"eq_mc_cycles": 2, "mc_vol_prob": 0.04, "list_update_int": 2}, 2, 4),
    "ice1_gen_weights_dd@2of4": ("ice1_gen_weights_dd", {"eq_mc_cycles": 2, "list_update_int": 2}, 2, 4),
}


# branches the decks do not take by default; CPU side only (the CUDA path is held to the oracle on the same branches in
# tests/test_gpu_mc.py::test_single_walker_chain_bit_exact and tests/test_gpu_kernels.py::test_chain_bit_exact_on_both_kernels)
VARIANT_CASES = {
    "ice1_sample/nvt": ("ice1_sample", {"eq_mc_cycles": 1, "mc_ensemble": "nvt", "list_update_int": 2}, 0, 1),
    "ice1_sample/leshift": ("ice1_sample", {"eq_mc_cycles": 1, "leshift": True, "mc_vol_prob": 0.04}, 0, 1),
    "ice1_sample/no_interp": ("ice1_sample", {"eq_mc_cycles": 1, "eta_interp": False, "mc_vol_prob": 0.04}, 0, 1),
    "ice1_gen_weights/switch_prob": ("ice1_gen_weights", {"eq_mc_cycles": 1, "mc_always_switch": False, "mc_switch_prob": 0.3,
                                                          "mc_vol_prob": 0.04}, 0, 1),
    "ice1_gen_weights/swetnam": ("ice1_gen_weights", {"eq_mc_cycles": 1, "wl_swetnam": True}, 0, 1),
}
ALL_CASES = {**CASES, **VARIANT_CASES}


def _oracle(name):
    return _make(*ALL_CASES[name])


def _make(deck, ov, rank=0, size=1):
    up, h, r, w, wl = load_example(deck, size=size)
    for k, v in ov.items():
        setattr(up, k, v)
    s = orc.System(up.nwater, up.num_lattices)
    s.set_config(r, h)
    s.energy_init()
    for ils in range(1, up.num_lattices + 1):
        s.compute_model_energy(ils)
    assert s.mc_init(orc.params_from_user(up), rank=rank, size=size, weights=w, file_wl_factor=wl) == 0
    return s, up


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_lists_and_energies_of_the_input_configuration(name):
    s, up = _oracle(name)
    nl = up.num_lattices
    for l in range(nl):
        nn, jn, vn = used_lists(s.nn[l], s.jn[l], s.vn[l])
        gnn, gjn, gvn = used_lists(V[f"{name}/nn"][l], V[f"{name}/jn"][l], V[f"{name}/vn"][l])
        np.testing.assert_array_equal(nn, gnn); np.testing.assert_array_equal(jn, gjn); np.testing.assert_array_equal(vn, gvn)
        loc = np.array([s.compute_local_real_energy(i + 1, l + 1) for i in range(up.nwater)])
        np.testing.assert_allclose(loc, V[f"{name}/local0"][l], rtol=1e-12, atol=0)
    e = np.array([s.compute_model_energy(l + 1) for l in range(nl)])
    np.testing.assert_allclose(e, V[f"{name}/energy0"], rtol=1e-12, atol=0)
    # grid and normalisation (the integer powers r**k are compiler-defined in the reference: ulp-level freedom)
    np.testing.assert_allclose(np.array(s.mu_bin), V[f"{name}/mu_bin"], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(np.array(s.binwidth), V[f"{name}/binwidth"], rtol=1e-12)
    sc = V[f"{name}/scalars"]
    assert abs(s.getd("r_pos") - sc[0]) < 1e-14 and abs(s.getd("r_neg") - sc[1]) < 1e-14
    assert abs(s.getd("av_binwidth") - sc[2]) < 1e-12
    if nl == 2:
        assert abs(s.getd("log_unbiased_norm") - sc[3]) < 1e-10
        assert abs(s.getd("ls_mu") - V[f"{name}/mu0"][0]) < 1e-9
        # the rank's window of the order parameter, the lattice it forces, the weights it keeps
        win = V[f"{name}/window"]
        assert (s.geti("my_start_bin"), s.geti("my_end_bin"), s.geti("ls")) == (int(win[0]), int(win[1]), int(win[4]))
        assert abs(s.getd("my_mu_min") - win[2]) < 1e-11 and abs(s.getd("my_mu_max") - win[3]) < 1e-11
        np.testing.assert_array_equal(np.array(s.weight), V[f"{name}/weight0"])


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_chain_under_the_same_fifo(name):
    """3 cycles (2 for weight generation) with volume moves and a list refresh inside: every accept / reject decision,
    every position and every counter of the C oracle equals the numpy restatement's."""
    s, up = _oracle(name)
    nl = up.num_lattices
    s.set_rng_fifo(V[f"{name}/fifo"])
    ncyc = int(V[f"{name}/ncycles"][0])
    assert s.mc_run(ncyc) == 0
    c = V[f"{name}/counters"]
    assert [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s")] == list(c[0:3])
    assert [s.geti("att_r"), s.geti("att_v"), s.geti("att_s")] == list(c[3:6])
    assert s.geti("ls") == c[6] and s.geti("rng_fifo_pos") == c[7] and s.geti("mc_cycle_num") == c[8]
    np.testing.assert_array_equal(np.array(s.ljr), V[f"{name}/ljr"])                 # bit for bit
    np.testing.assert_array_equal(np.array(s.ref_ljr), V[f"{name}/ref_ljr"])
    np.testing.assert_array_equal(np.array(s.hmatrix), V[f"{name}/hmatrix"])
    np.testing.assert_array_equal(np.array(s.mc_translations), V[f"{name}/mc_translations"])
    np.testing.assert_array_equal(np.array(s.nn), V[f"{name}/nn_end"])
    np.testing.assert_allclose(np.array(s.model_energy), V[f"{name}/energy"], rtol=1e-12)
    np.testing.assert_allclose(np.array(s.volume), V[f"{name}/volume"], rtol=1e-15)
    np.testing.assert_allclose(np.array(s.arr_d("average_energy", (2,)))[:nl], V[f"{name}/average_energy"], rtol=1e-12)
    if nl == 2:
        assert abs(s.getd("ls_mu") - V[f"{name}/mu"][0]) < 1e-9
        np.testing.assert_allclose(np.array(s.histogram), V[f"{name}/histogram"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(np.array(s.weight), V[f"{name}/weight"], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(np.array(s.unbiased_hist), V[f"{name}/unbiased_hist"], rtol=1e-9, atol=1e-300)
    tr = V[f"{name}/trace"]
    assert (tr[:, 0] == 1).sum() == c[4] and tr[tr[:, 0] == 0][:, 1].sum() == c[0]


EVENT_CASES = {
    # key: (deck, overrides, event) -- tests/golden/make_fixtures_numpy.py
    "ice1_sample+monitor": ("ice1_sample", {"eq_mc_cycles": 100, "mc_vol_prob": 0.06, "list_update_int": 2}, "monitor"),
    "single_box+monitor": ("single_box", {"eq_mc_cycles": 100, "mc_vol_prob": 0.06, "list_update_int": 2}, "monitor"),
    "ice1_sample+chain_sync": ("ice1_sample", {"eq_mc_cycles": 1, "mc_vol_prob": 0.06, "list_update_int": 2}, "chain_sync"),
}


@pytest.mark.parametrize("name", list(EVENT_CASES))
def test_monitor_and_chain_synchronisation_between_two_stretches(name):
    """mc_monitor_stats (mc_moves.F90:1722-1732, :1786-1810: step sizes re-tuned, stored energies replaced, counters
    reset) and mc_check_chain_synchronisation (:2217-2416: lattice 2 forced onto lattice 1's displacements) between
    two stretches of cycles: the state right after the event and the chain that runs on it."""
    deck, ov, event = EVENT_CASES[name]
    s, up = _make(deck, ov)
    nl = up.num_lattices
    s.set_rng_fifo(V[f"{name}/fifo"])
    n1, n2 = [int(x) for x in V[f"{name}/ncycles"]]
    assert s.mc_run(n1) == 0
    pre = V[f"{name}/pre_counters"]
    assert [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s"), s.geti("att_r"), s.geti("att_v"), s.geti("att_s")] == list(pre)
    if event == "monitor":
        s.mc_monitor()
        st = V[f"{name}/mid_steps"]
        assert s.getd("mc_max_trans") == st[0] and s.getd("mc_dv_max") == st[1]
        assert [s.geti("acc_r"), s.geti("att_r"), s.geti("att_v")] == [0, 0, 0]
    else:
        s.mc_chain_sync()
    np.testing.assert_array_equal(np.array(s.ljr), V[f"{name}/mid_ljr"])
    np.testing.assert_array_equal(np.array(s.hmatrix), V[f"{name}/mid_hmatrix"])
    np.testing.assert_allclose(np.array(s.model_energy), V[f"{name}/mid_energy"], rtol=1e-12)
    if nl == 2:
        assert abs(s.getd("ls_mu") - V[f"{name}/mid_mu"][0]) < 1e-9
    assert s.mc_run(n2) == 0
    c = V[f"{name}/counters"]
    assert [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s")] == list(c[0:3])
    assert [s.geti("att_r"), s.geti("att_v"), s.geti("att_s")] == list(c[3:6])
    assert s.geti("ls") == c[6] and s.geti("rng_fifo_pos") == c[7] and s.geti("mc_cycle_num") == c[8]
    np.testing.assert_array_equal(np.array(s.ljr), V[f"{name}/ljr"])
    np.testing.assert_array_equal(np.array(s.ref_ljr), V[f"{name}/ref_ljr"])
    np.testing.assert_array_equal(np.array(s.hmatrix), V[f"{name}/hmatrix"])
    np.testing.assert_array_equal(np.array(s.mc_translations), V[f"{name}/mc_translations"])
    np.testing.assert_allclose(np.array(s.model_energy), V[f"{name}/energy"], rtol=1e-12)
    np.testing.assert_allclose(np.array(s.volume), V[f"{name}/volume"], rtol=1e-15)
    np.testing.assert_allclose(np.array(s.arr_d("average_energy", (2,)))[:nl], V[f"{name}/average_energy"], rtol=1e-12)
    if nl == 2:
        assert abs(s.getd("ls_mu") - V[f"{name}/mu"][0]) < 1e-9


def test_restatement_is_independent_of_the_oracle():
    """ref_numpy.py must not lean on the oracle (or on the product): it is the second opinion."""
    src = open(os.path.join(GOLDEN, "ref_numpy.py")).read()
    assert "oracle" not in src.split('"""', 2)[2] and "mwgpu" not in src and "ctypes" not in src
