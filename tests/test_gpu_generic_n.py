"""Boxes that are not the 48-molecule boxes of the reference decks: the walker kernel then runs its
run-time-N instantiation (k_mc_run<NLAT, 0>) instead of the N = 48 one.  A 2x1x1 supercell (96 molecules) of
the example lattices, same parity bar as everywhere: lists / positions / counters bit-exact, energies 1e-11."""
import numpy as np
import pytest

from mc_water_ls_mw_b200 import walkers as W
from oracle import orc
from tests.helpers import load_example, rel_err, used_lists

pytestmark = pytest.mark.gpu
SEED = 20141211


def _supercell(ex, nwalkers, ov):
    up, h, r, w, wl = load_example(ex, size=nwalkers)
    for k, v in ov.items():
        setattr(up, k, v)
    nl, n = up.num_lattices, up.nwater
    h2 = h.copy(); r2 = np.zeros((nl, 2 * n, 3))
    for l in range(nl):
        a = h[l][0:3].copy()                         # first cell vector (column 1 of hmatrix)
        h2[l][0:3] = 2.0 * a
        r2[l, :n] = r[l]; r2[l, n:] = r[l] + a
    up.nwater = 2 * n
    g = W.WalkerBatch(up.nwater, nl, nwalkers)
    g.upload(r2, h2); g.energy_init()
    g.mc_init(W.params_from_user(up), 0, nwalkers, w, wl)
    ws = []
    for i in range(nwalkers):
        s = orc.System(up.nwater, nl)
        s.set_config(r2, h2); s.energy_init()
        assert s.mc_init(orc.params_from_user(up), rank=i, size=nwalkers, weights=w, file_wl_factor=wl) == 0
        ws.append(s)
    return g, ws, up


@pytest.mark.parametrize("ex", ["ice1_sample", "single_box"])
def test_supercell_96_molecules(ex):
    g, ws, up = _supercell(ex, 2, {"eq_mc_cycles": 3})
    o = ws[0]
    for l in range(1, up.num_lattices + 1):
        nn, jn, vn = g.get_neighbours(l)
        onn, ojn, ovn = used_lists(o.nn[l - 1], o.jn[l - 1], o.vn[l - 1])
        np.testing.assert_array_equal(nn, onn); np.testing.assert_array_equal(jn, ojn); np.testing.assert_array_equal(vn, ovn)
        assert rel_err(g.compute_model_energy(l), o.model_energy[l - 1]) < 1e-11
        loc = g.compute_local_real_energy_all(l)
        assert rel_err(loc, [o.compute_local_real_energy(i + 1, l) for i in range(up.nwater)]) < 1e-11
    g.set_rng_philox(SEED, 0, 1000000)
    for i, s in enumerate(ws):
        s.set_rng_philox(SEED, i, 1000000)
    for chunk in (3, 9):
        g.mc_run(chunk)
        for s in ws:
            assert s.mc_run(chunk) == 0
    for w, s in enumerate(ws):
        st = g.state(w)
        ljr, ref, hm = g.download(w)
        np.testing.assert_array_equal(ljr, s.ljr); np.testing.assert_array_equal(hm, s.hmatrix)
        assert list(st.accepted) == [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s")]
        assert list(st.attempted) == [s.geti("att_r"), s.geti("att_v"), s.geti("att_s")]
        assert st.attempted[0] > 0 and st.accepted[0] > 0
        assert st.rng_index == s.geti("rng_index") and st.ls == s.geti("ls")
        assert rel_err(list(st.model_energy)[: up.num_lattices], s.model_energy) < 1e-11
        np.testing.assert_array_equal(g.translations(w), s.mc_translations)
    e_all = g.compute_model_energy_all()
    for w, s in enumerate(ws):
        assert rel_err(e_all[w][: up.num_lattices], [s.compute_model_energy(l) for l in range(1, up.num_lattices + 1)]) < 1e-11
