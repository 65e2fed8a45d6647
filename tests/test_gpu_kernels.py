"""The walker kernels against the oracle and against each other.

mwgpu_mc_set_kernel(1): one warp per walker (first generation, mw_mc.cuh -- what boxes of more than 64 molecules
run); (2): one warp per LATTICE on a per-lattice shared-memory block (mw2.cuh, the default for the reference's
48-molecule boxes; mc_moves.F90:1007-1018, :1076-1090 are the per-lattice loops it runs side by side); (4): two warps
per lattice (the two warps of a lattice split the item passes and the old / new pair sums of
compute_local_real_energy, molint.F90:276-404; on request only).  Same bar for
all: positions / cell / lists / counters / random-number consumption bit-exact, energies 1e-11 relative."""
import numpy as np
import pytest

from tests.helpers import make_gpu_walkers, make_oracle_walker, make_oracle_walkers, rel_err, used_lists

pytestmark = pytest.mark.gpu
SEED = 20141211
TOL = 1e-11


def _same_state(g, o, up, w=0):
    s = g.state(w)
    ljr, ref, hm = g.download(w)
    np.testing.assert_array_equal(ljr, o.ljr); np.testing.assert_array_equal(ref, o.ref_ljr)
    np.testing.assert_array_equal(hm, o.hmatrix)
    assert list(s.accepted) == [o.geti("acc_r"), o.geti("acc_v"), o.geti("acc_s")]
    assert list(s.attempted) == [o.geti("att_r"), o.geti("att_v"), o.geti("att_s")]
    assert s.rng_index == o.geti("rng_index") and s.ls == o.geti("ls")
    nl = up.num_lattices
    assert rel_err(list(s.model_energy)[:nl], o.model_energy) < TOL
    assert abs(s.ls_mu - o.getd("ls_mu")) < 1e-9 * max(1.0, abs(s.ls_mu))
    for l in range(1, nl + 1):
        nn, jn, vn = g.get_neighbours(l, w)
        onn, ojn, ovn = used_lists(o.nn[l - 1], o.jn[l - 1], o.vn[l - 1])
        np.testing.assert_array_equal(nn, onn); np.testing.assert_array_equal(jn, ojn); np.testing.assert_array_equal(vn, ovn)
    wg, hg, ug = g.bins(w)
    np.testing.assert_allclose(hg, o.histogram, rtol=0, atol=1e-9)
    np.testing.assert_allclose(wg, o.weight, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(ug, o.unbiased_hist, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("kernel", [1, 2, 4])
@pytest.mark.parametrize("ex,ncyc,ov", [
    ("ice1_sample", 45, {"eq_mc_cycles": 5}),
    ("ice1_gen_weights", 45, {"eq_mc_cycles": 5}),
    ("single_box", 45, {"eq_mc_cycles": 5}),
    ("ice1_sample", 25, {"mc_vol_prob": 0.08, "eq_mc_cycles": 3}),          # volume moves next to cycle ends / list rebuilds
    ("ice1_sample", 25, {"mc_always_switch": False, "mc_switch_prob": 0.2, "eq_mc_cycles": 3}),
    # Swetnam's increment (mc_moves.F90:1636-1653): log(rms) < 0 makes it negative, a visit can then create a new
    # window minimum that the reference subtracts from every bin (:1682-1685)
    ("ice1_gen_weights", 30, {"wl_swetnam": True, "eq_mc_cycles": 3}),
])
def test_chain_bit_exact_on_both_kernels(kernel, ex, ncyc, ov):
    o, up = make_oracle_walker(ex, overrides=ov)
    g, _ = make_gpu_walkers(ex, overrides=ov)
    if kernel == 4 and up.num_lattices == 1:
        pytest.skip("four warps per walker = two per lattice of a lattice-switch box")
    g.set_kernel(kernel)
    o.set_rng_philox(SEED, 7, 1000000); g.set_rng_philox(SEED, 7, 1000000)
    done = 0
    for chunk in (1, 9, ncyc):
        n = min(chunk, ncyc - done)
        if n <= 0:
            break
        g.mc_run(n); assert o.mc_run(n) == 0
        done += n
        _same_state(g, o, up)


def test_kernels_agree_on_a_batch_and_are_deterministic():
    """256 walkers x 40 cycles with many volume moves: kernel 1 == kernel 2 == kernel 2 again, bit for bit in the
    state arithmetic (the two warps of a walker meet at barriers around every rare move: a missing one shows up
    here as run-to-run differences)."""
    nw, ncyc = 256, 40
    ov = {"eq_mc_cycles": 2, "mc_vol_prob": 0.03}
    out = []
    for kernel in (1, 2, 2, 4, 4):
        g, up = make_gpu_walkers("ice1_sample", nwalkers=nw, overrides=ov)
        g.set_kernel(kernel)
        g.set_rng_philox(SEED, 0, 1000000)
        for _ in range(4):
            g.mc_run(ncyc // 4)
        st = g.states()
        assert not any(s.error for s in st)
        out.append((g.download_all(), np.array([[s.accepted[0], s.accepted[1], s.accepted[2], s.rng_index, s.ls] for s in st]),
                    np.array([list(s.model_energy) for s in st])))
    for a, b in ((out[0], out[1]), (out[1], out[2]), (out[2], out[3]), (out[3], out[4])):
        for x, y in zip(a[0], b[0]):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(a[1], b[1])
    assert rel_err(out[0][2], out[1][2]) < TOL
    assert rel_err(out[2][2], out[3][2]) < TOL
    np.testing.assert_array_equal(out[3][2], out[4][2])
    np.testing.assert_array_equal(out[1][2], out[2][2])                     # same kernel: same bits in the energies too


def test_kernel_selection_errors():
    from mc_water_ls_mw_b200._lib import MwgpuError
    g, _ = make_gpu_walkers("ice1_sample")
    with pytest.raises(MwgpuError):
        g.set_kernel(3)
    g.set_kernel(4)
    g.set_kernel(0)
    g1, _ = make_gpu_walkers("single_box")
    with pytest.raises(MwgpuError):
        g1.set_kernel(4)


def test_crowded_cells_split_the_variants():
    """Compressed cells (16 bonds per molecule): old and new bonds of a move no longer share the record table of a
    lattice, the warp-per-lattice kernel then evaluates the two variants one after the other (mw2.cuh) -- same chain."""
    from mc_water_ls_mw_b200 import walkers as W
    from oracle import orc
    from tests.helpers import load_example
    up, h, r, w, wl = load_example("ice1_sample")
    up.eq_mc_cycles = 2
    rng = np.random.default_rng(11)
    ljr = np.asarray(r) * 0.93 + rng.normal(0.0, 0.02, np.asarray(r).shape)
    hm = np.asarray(h) * 0.93
    o = orc.System(up.nwater, up.num_lattices); o.set_config(ljr, hm); o.energy_init()
    assert o.mc_init(orc.params_from_user(up), rank=0, size=1, weights=w, file_wl_factor=wl) == 0
    from mc_water_ls_mw_b200._lib import MwgpuError
    o.set_rng_philox(SEED, 0, 1000000)
    assert o.mc_run(6) == 0
    # 0 = automatic: one walker on the GPU runs the instantiation for at most four walkers per SM, whose item loop takes
    # up to three passes per turn -- these 90-odd items per round are what drives it through the three-pass form
    for kernel in (2, 4, 0, 1):
        g = W.WalkerBatch(up.nwater, up.num_lattices, 1)
        g.upload(ljr, hm); g.energy_init()
        g.mc_init(W.params_from_user(up), 0, 1, w, wl)
        g.set_kernel(kernel)
        g.set_rng_philox(SEED, 0, 1000000)
        if kernel == 1:
            # the one-warp kernel keeps the bonds of both lattices and both variants in one 64-record table:
            # it reports the overflow instead of a result (envelope documented in include/mwgpu.h)
            with pytest.raises(MwgpuError, match="bonds"):
                g.mc_run(6)
            continue
        g.mc_run(6)
        assert g.state().error == 0
        np.testing.assert_array_equal(g.download()[0], o.ljr)
        assert rel_err(list(g.state().model_energy), o.model_energy) < TOL
