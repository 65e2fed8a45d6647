"""GPU parity: module `energy` (molint.F90) through the C ABI vs the CPU oracle.

Bar: neighbour lists and image vectors bit-exact; fp64 energies within 1e-11 relative
(BASELINE.json north_star).  Summation order on the GPU: pairwise shuffle tree over
per-lane partial sums."""
import os

import numpy as np
import pytest

from tests.helpers import (GOLDEN, load_example, make_gpu_walkers, make_oracle_walker, rel_err, used_lists)

pytestmark = pytest.mark.gpu
TOL = 1e-11
EXAMPLES = ["ice1_sample", "single_box", "ice1_gen_weights", "ice1_sample_dd"]


@pytest.mark.parametrize("ex", EXAMPLES)
def test_energy_init_parity(ex):
    o, up = make_oracle_walker(ex)
    g, _ = make_gpu_walkers(ex, mc=False)
    for l in range(1, up.num_lattices + 1):
        n, iv = g.compute_ivects(l)
        assert n == o.nivect[l - 1] == 27
        np.testing.assert_array_equal(iv[:n], o.ivect[l - 1][:n])           # bit-exact
        nn, jn, vn = g.compute_neighbours(l)
        onn, ojn, ovn = used_lists(o.nn[l - 1], o.jn[l - 1], o.vn[l - 1])
        np.testing.assert_array_equal(nn, onn)
        np.testing.assert_array_equal(jn, ojn)                               # contents AND order
        np.testing.assert_array_equal(vn, ovn)
        assert rel_err(g.compute_model_energy(l), o.model_energy[l - 1]) < TOL
        loc = g.compute_local_real_energy_all(l)
        oloc = [o.compute_local_real_energy(i + 1, l) for i in range(up.nwater)]
        assert rel_err(loc, oloc) < TOL
        assert rel_err(g.compute_local_real_energy(7, l), oloc[6]) < TOL     # the 1:1 fine-grained call


def test_golden_vectors():
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    for ex in ("ice1_sample", "single_box"):
        g, up = make_gpu_walkers(ex, mc=False)
        for l in range(1, up.num_lattices + 1):
            nn, jn, vn = g.get_neighbours(l)
            np.testing.assert_array_equal(nn, v[f"{ex}/nn"][l - 1])
            np.testing.assert_array_equal(jn, v[f"{ex}/jn"][l - 1])
            np.testing.assert_array_equal(vn, v[f"{ex}/vn"][l - 1])
            assert rel_err(g.compute_model_energy(l), v[f"{ex}/energy0"][l - 1]) < TOL
            assert rel_err(g.compute_local_real_energy_all(l), v[f"{ex}/local"][l - 1]) < TOL


def _thermalised(ex, ncycles, seed_stream=0):
    o, up = make_oracle_walker(ex)
    o.set_rng_philox(20141211, seed_stream, 1000000)
    assert o.mc_run(ncycles) == 0
    return o, up


@pytest.mark.parametrize("ex,ncyc", [("ice1_sample", 60), ("ice1_sample", 400), ("single_box", 150)])
def test_thermalised_configurations(ex, ncyc):
    """Disordered configurations (up to ~10 molecules inside the cut-off, lists of 17..23 entries)."""
    from mc_water_ls_mw_b200 import walkers as W
    o, up = _thermalised(ex, ncyc)
    g = W.WalkerBatch(up.nwater, up.num_lattices, 1)
    g.upload(np.array(o.ljr), np.array(o.hmatrix), np.array(o.ref_ljr))
    g.energy_init()
    for l in range(1, up.num_lattices + 1):
        o.compute_neighbours(l)
        nn, jn, vn = g.get_neighbours(l)
        onn, ojn, ovn = used_lists(o.nn[l - 1], o.jn[l - 1], o.vn[l - 1])
        np.testing.assert_array_equal(nn, onn)
        np.testing.assert_array_equal(jn, ojn)
        np.testing.assert_array_equal(vn, ovn)
        assert rel_err(g.compute_model_energy(l), o.compute_model_energy(l)) < TOL
        loc = g.compute_local_real_energy_all(l)
        oloc = [o.compute_local_real_energy(i + 1, l) for i in range(up.nwater)]
        assert rel_err(loc, oloc) < TOL
        # the local energies count every pair twice and every triplet three times
        assert abs(sum(loc)) > abs(g.compute_model_energy(l))


def test_upload_download_roundtrip_and_layout():
    from mc_water_ls_mw_b200 import walkers as W
    up, h, r, _, _ = load_example("ice1_sample")
    rng = np.random.default_rng(3)
    nw = 5
    ljr = np.ascontiguousarray(r[None] + rng.normal(0, 0.05, (nw, 2, 48, 3)))
    hm = np.ascontiguousarray(np.repeat(h[None], nw, 0))
    g = W.WalkerBatch(48, 2, nw)
    g.upload_all(ljr, hm)
    l2, r2, h2 = g.download_all()
    np.testing.assert_array_equal(l2, ljr); np.testing.assert_array_equal(r2, ljr); np.testing.assert_array_equal(h2, hm)
    a, b, c = g.download(3)
    np.testing.assert_array_equal(a, ljr[3]); np.testing.assert_array_equal(c, hm[3])


def test_batched_model_energy_matches_per_walker():
    """The "full mW energy evals/s" kernel: every (walker, lattice) in one launch."""
    from mc_water_ls_mw_b200 import walkers as W
    from oracle import orc
    up, h, r, _, _ = load_example("ice1_sample")
    rng = np.random.default_rng(11)
    nw = 37                                                              # ragged vs any tile size
    ljr = np.ascontiguousarray(r[None] + rng.normal(0, 0.08, (nw, 2, 48, 3)))
    hm = np.ascontiguousarray(np.repeat(h[None], nw, 0) * (1.0 + rng.normal(0, 0.004, (nw, 1, 1))))
    g = W.WalkerBatch(48, 2, nw)
    g.upload_all(ljr, hm)
    g.energy_init()
    e = g.compute_model_energy_all()
    for w in range(nw):
        o = orc.System(48, 2); o.set_config(ljr[w], hm[w]); o.energy_init()
        assert rel_err(e[w], o.model_energy) < TOL
        nn, jn, vn = g.get_neighbours(2, walker=w)
        onn, ojn, ovn = used_lists(o.nn[1], o.jn[1], o.vn[1])
        np.testing.assert_array_equal(nn, onn); np.testing.assert_array_equal(jn, ojn); np.testing.assert_array_equal(vn, ovn)
    # idempotence: a second evaluation returns bit-identical energies
    np.testing.assert_array_equal(e, g.compute_model_energy_all())


def test_translation_invariance_property():
    """Size-independent property: shifting every molecule by a lattice vector changes neither lists
    nor energy (the positions are not wrapped, data_structures.f90:39)."""
    from mc_water_ls_mw_b200 import walkers as W
    up, h, r, _, _ = load_example("ice1_sample")
    g = W.WalkerBatch(48, 2, 2)
    shifted = r.copy()
    for l in range(2):
        shifted[l] += h[l].reshape(3, 3)[1]          # second cell vector (column-major storage)
    g.upload_all(np.ascontiguousarray(np.stack([r, shifted])), np.ascontiguousarray(np.stack([h, h])))
    g.energy_init()
    e = g.compute_model_energy_all()
    assert rel_err(e[1], e[0]) < 1e-12
    for l in (1, 2):
        a = g.get_neighbours(l, 0); b = g.get_neighbours(l, 1)
        np.testing.assert_array_equal(a[0], b[0]); np.testing.assert_array_equal(a[1], b[1])


def test_error_behaviour():
    from mc_water_ls_mw_b200 import walkers as W
    from mc_water_ls_mw_b200._lib import MwgpuError
    g, up = make_gpu_walkers("single_box", mc=False)
    with pytest.raises(MwgpuError):
        g.compute_model_energy(2)                    # only one lattice
    with pytest.raises(MwgpuError):
        g.compute_local_real_energy(49, 1)
    with pytest.raises(MwgpuError):
        g.mc_run(1)                                  # mc_init not called
    with pytest.raises(MwgpuError):
        W.WalkerBatch(48, 3, 1)                      # 'Error num_lattices must equal 1 or 2!' (io.f90:216)
    # a cell narrower than the list radius makes a molecule its own neighbour: reported, not ignored
    up2, h, r, _, _ = load_example("single_box")
    g2 = W.WalkerBatch(48, 1, 1)
    g2.upload(r * 0.3, h * 0.3)
    with pytest.raises(MwgpuError):
        g2.energy_init()


@pytest.mark.parametrize("scale", [0.94, 0.92])
def test_compressed_lattices_crowded_neighbourhoods(scale):
    """Both example lattices compressed until the second shell (12 molecules) is inside the cut-off:
    16 bonds per molecule, 28 list entries, > 256 (centre, bond) candidates per local energy -- the
    kernel then walks the triplet centres in groups of 8 (mw_device.cuh, stages 4+5)."""
    from mc_water_ls_mw_b200 import walkers as W
    from oracle import orc
    up, h, r, _, _ = load_example("ice1_sample")
    rng = np.random.default_rng(11)
    ljr = np.asarray(r) * scale + rng.normal(0.0, 0.02, np.asarray(r).shape)
    hm = np.asarray(h) * scale
    o = orc.System(up.nwater, up.num_lattices)
    o.set_config(ljr, hm)
    o.energy_init()
    g = W.WalkerBatch(up.nwater, up.num_lattices, 1)
    g.upload(ljr, hm)
    g.energy_init()
    for l in (1, 2):
        nn, jn, vn = g.get_neighbours(l)
        onn, ojn, ovn = used_lists(o.nn[l - 1], o.jn[l - 1], o.vn[l - 1])
        np.testing.assert_array_equal(nn, onn); np.testing.assert_array_equal(jn, ojn); np.testing.assert_array_equal(vn, ovn)
        assert nn.max() >= 26
        assert rel_err(g.compute_model_energy(l), o.compute_model_energy(l)) < TOL
        loc = g.compute_local_real_energy_all(l)
        oloc = [o.compute_local_real_energy(i + 1, l) for i in range(up.nwater)]
        assert rel_err(loc, oloc) < TOL
