"""GPU parity against the SECOND restatement directly: the vectors of tests/golden/ref_numpy.py (plain Python written
from the Fortran, tests/golden/numpy_vectors.npz) without the C oracle in between -- lists, energies, and the chain
under the same host FIFO of random numbers: positions / cell / counters / random-number consumption bit for bit,
energies to 1e-11 relative (the tolerance of BASELINE.json's north_star; summation order in DESIGN.md section 2).
Covers the three multi-walker decks and single windows of the two domain-decomposed decks (mc_moves.F90:660-703)."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, make_gpu_walkers, rel_err, used_lists
from tests.test_oracle_numpy import CASES, EVENT_CASES

pytestmark = pytest.mark.gpu
TOL = 1e-11
V = np.load(os.path.join(GOLDEN, "numpy_vectors.npz"))


def _walker(name):
    deck, ov, rank, size = CASES[name]
    return make_gpu_walkers(deck, nwalkers=1, first_rank=rank, size=size, overrides=ov)


@pytest.mark.parametrize("name", list(CASES))
def test_input_configuration_against_the_numpy_restatement(name):
    g, up = _walker(name)
    nl = up.num_lattices
    for l in range(nl):
        nn, jn, vn = g.get_neighbours(l + 1, 0)
        gnn, gjn, gvn = used_lists(V[f"{name}/nn"][l], V[f"{name}/jn"][l], V[f"{name}/vn"][l])
        np.testing.assert_array_equal(nn, gnn); np.testing.assert_array_equal(jn, gjn); np.testing.assert_array_equal(vn, gvn)
        loc = np.array([g.compute_local_real_energy(i + 1, l + 1, 0) for i in range(up.nwater)])
        assert rel_err(loc, V[f"{name}/local0"][l]) < TOL
    s = g.state(0)
    assert rel_err(list(s.model_energy)[:nl], V[f"{name}/energy0"]) < TOL
    if nl == 2:
        win = V[f"{name}/window"]
        assert (s.my_start_bin, s.my_end_bin, s.ls) == (int(win[0]), int(win[1]), int(win[4]))
        assert abs(s.my_mu_min - win[2]) < 1e-11 and abs(s.my_mu_max - win[3]) < 1e-11
        assert abs(s.ls_mu - V[f"{name}/mu0"][0]) < 1e-9
        np.testing.assert_array_equal(g.bins(0)[0], V[f"{name}/weight0"])


@pytest.mark.parametrize("name", list(CASES))
def test_chain_against_the_numpy_restatement(name):
    g, up = _walker(name)
    nl = up.num_lattices
    g.set_rng_fifo(V[f"{name}/fifo"])
    g.mc_run(int(V[f"{name}/ncycles"][0]))
    s = g.state(0)
    c = V[f"{name}/counters"]
    assert list(s.accepted) == list(c[0:3]) and list(s.attempted) == list(c[3:6])
    assert s.ls == c[6] and s.rng_index == c[7] and s.mc_cycle_num == c[8]
    ljr, ref, hm = g.download(0)
    np.testing.assert_array_equal(ljr, V[f"{name}/ljr"])                 # bit for bit
    np.testing.assert_array_equal(ref, V[f"{name}/ref_ljr"])
    np.testing.assert_array_equal(hm, V[f"{name}/hmatrix"])
    np.testing.assert_array_equal(g.translations(0), V[f"{name}/mc_translations"])
    assert rel_err(list(s.model_energy)[:nl], V[f"{name}/energy"]) < TOL
    assert rel_err(list(s.volume)[:nl], V[f"{name}/volume"]) < 1e-15
    assert rel_err(list(s.average_energy)[:nl], V[f"{name}/average_energy"]) < TOL
    for l in range(nl):
        nn, _, _ = g.get_neighbours(l + 1, 0)
        np.testing.assert_array_equal(nn, V[f"{name}/nn_end"][l])
    if nl == 2:
        assert abs(s.ls_mu - V[f"{name}/mu"][0]) < 1e-9
        w, h, u = g.bins(0)
        np.testing.assert_allclose(h, V[f"{name}/histogram"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(w, V[f"{name}/weight"], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(u, V[f"{name}/unbiased_hist"], rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("name", list(EVENT_CASES))
def test_monitor_and_chain_synchronisation_against_the_numpy_restatement(name):
    """mc_monitor_stats / mc_check_chain_synchronisation between two stretches of cycles (mc_moves.F90:1722-1810,
    :2217-2416): the state right after the event and the chain that runs on it, against the numpy vectors."""
    deck, ov, event = EVENT_CASES[name]
    g, up = make_gpu_walkers(deck, nwalkers=1, overrides=ov)
    nl = up.num_lattices
    g.set_rng_fifo(V[f"{name}/fifo"])
    n1, n2 = [int(x) for x in V[f"{name}/ncycles"]]
    g.mc_run(n1)
    s = g.state(0)
    assert list(s.accepted) + list(s.attempted) == list(V[f"{name}/pre_counters"])
    if event == "monitor":
        g.mc_monitor()
        s = g.state(0)
        st = V[f"{name}/mid_steps"]
        assert s.mc_max_trans == st[0] and s.mc_dv_max == st[1]                  # bit-exact
        assert list(s.attempted) == [0, 0, 0]
    else:
        g.mc_chain_sync()
        s = g.state(0)
    ljr, ref, hm = g.download(0)
    np.testing.assert_array_equal(ljr, V[f"{name}/mid_ljr"])
    np.testing.assert_array_equal(hm, V[f"{name}/mid_hmatrix"])
    assert rel_err(list(s.model_energy)[:nl], V[f"{name}/mid_energy"]) < TOL
    if nl == 2:
        assert abs(s.ls_mu - V[f"{name}/mid_mu"][0]) < 1e-9
    g.mc_run(n2)
    s = g.state(0)
    c = V[f"{name}/counters"]
    assert list(s.accepted) == list(c[0:3]) and list(s.attempted) == list(c[3:6])
    assert s.ls == c[6] and s.rng_index == c[7] and s.mc_cycle_num == c[8]
    ljr, ref, hm = g.download(0)
    np.testing.assert_array_equal(ljr, V[f"{name}/ljr"])
    np.testing.assert_array_equal(ref, V[f"{name}/ref_ljr"])
    np.testing.assert_array_equal(hm, V[f"{name}/hmatrix"])
    np.testing.assert_array_equal(g.translations(0), V[f"{name}/mc_translations"])
    assert rel_err(list(s.model_energy)[:nl], V[f"{name}/energy"]) < TOL
    assert rel_err(list(s.volume)[:nl], V[f"{name}/volume"]) < 1e-15
    assert rel_err(list(s.average_energy)[:nl], V[f"{name}/average_energy"]) < TOL
    if nl == 2:
        assert abs(s.ls_mu - V[f"{name}/mu"][0]) < 1e-9
