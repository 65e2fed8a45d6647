"""GPU parity: the move loop of mc_moves.F90 through the C ABI vs the CPU oracle.

Bar (BASELINE.json north_star): under the same random stream, positions / cell / neighbour lists /
accept-reject counts bit-exact in single-walker mode; energies and the order parameter within 1e-11
relative (the GPU evaluates exp/log with CUDA's libdevice and sums in a different order)."""
import numpy as np
import pytest

from tests.helpers import load_example, make_gpu_walkers, make_oracle_walker, make_oracle_walkers, rel_err, used_lists

pytestmark = pytest.mark.gpu
TOL = 1e-11
SEED = 20141211


def _compare(g, o, up, walker=0, bins=True):
    s = g.state(walker)
    ljr, ref, hm = g.download(walker)
    np.testing.assert_array_equal(ljr, o.ljr)                       # bit-exact
    np.testing.assert_array_equal(ref, o.ref_ljr)
    np.testing.assert_array_equal(hm, o.hmatrix)
    assert list(s.accepted) == [o.geti("acc_r"), o.geti("acc_v"), o.geti("acc_s")]
    assert list(s.attempted) == [o.geti("att_r"), o.geti("att_v"), o.geti("att_s")]
    assert s.ls == o.geti("ls") and s.mc_cycle_num == o.geti("mc_cycle_num")
    assert s.rng_index == o.geti("rng_index")
    assert s.walker_in_window == o.geti("walker_in_window")
    nl = up.num_lattices
    assert rel_err(list(s.model_energy)[:nl], o.model_energy) < TOL
    assert rel_err(list(s.volume)[:nl], o.volume) < 1e-15
    assert abs(s.ls_mu - o.getd("ls_mu")) < 1e-9 * max(1.0, abs(s.ls_mu))
    assert rel_err(list(s.average_energy)[:nl], np.array(o.arr_d("average_energy", (2,)))[:nl]) < TOL
    np.testing.assert_array_equal(g.translations(walker), o.mc_translations)
    for l in range(1, nl + 1):
        nn, jn, vn = g.get_neighbours(l, walker)
        onn, ojn, ovn = used_lists(o.nn[l - 1], o.jn[l - 1], o.vn[l - 1])
        np.testing.assert_array_equal(nn, onn); np.testing.assert_array_equal(jn, ojn); np.testing.assert_array_equal(vn, ovn)
    if bins:
        w, h, u = g.bins(walker)
        np.testing.assert_allclose(h, o.histogram, rtol=0, atol=1e-9)
        np.testing.assert_allclose(w, o.weight, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(u, o.unbiased_hist, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("ex,ncyc,ov", [
    ("ice1_sample", 120, {}),                                           # configs[1]: fixed weights (equilibration phase)
    ("ice1_sample", 60, {"eq_mc_cycles": 5}),                           # ... with histogram + unbiased histogram updates
    ("ice1_gen_weights", 60, {"eq_mc_cycles": 5}),                      # configs[2]: Wang-Landau weight updates
    ("single_box", 120, {"eq_mc_cycles": 5}),                           # configs[0]: one lattice, no switch
    ("ice1_sample", 40, {"mc_ensemble": "nvt", "eq_mc_cycles": 5}),     # nvt branch of the switch
    ("ice1_sample", 40, {"leshift": True, "eq_mc_cycles": 5}),
    ("ice1_sample", 40, {"eta_interp": False, "eq_mc_cycles": 5}),
    ("ice1_sample", 30, {"mc_vol_prob": 0.05, "eq_mc_cycles": 5}),     # many volume moves (accepted and rejected)
    ("ice1_gen_weights", 30, {"mc_always_switch": False, "mc_switch_prob": 0.3, "eq_mc_cycles": 5}),
])
def test_single_walker_chain_bit_exact(ex, ncyc, ov):
    o, up = make_oracle_walker(ex, overrides=ov)
    g, _ = make_gpu_walkers(ex, overrides=ov)
    mu, bw, sc = g.grid()
    np.testing.assert_array_equal(mu, o.mu_bin); np.testing.assert_array_equal(bw, o.binwidth)
    assert sc["log_unbiased_norm"] == o.getd("log_unbiased_norm") and sc["r_pos"] == o.getd("r_pos")
    assert abs(g.state().ls_mu - o.getd("ls_mu")) < 1e-9
    o.set_rng_philox(SEED, 0, 1000000); g.set_rng_philox(SEED, 0, 1000000)
    # uneven launch sizes: the chain must not depend on how the cycles are cut into launches
    done = 0
    for chunk in (1, 7, 13, ncyc):
        n = min(chunk, ncyc - done)
        if n <= 0:
            break
        g.mc_run(n); assert o.mc_run(n) == 0
        done += n
        _compare(g, o, up)
    assert done == ncyc
    a = g.state().accepted
    assert a[0] > 0                                                  # the chain actually moved
    if "mc_vol_prob" in ov:
        assert g.state().attempted[1] > 20


def test_host_fifo_rng_mode_matches_serial_stream():
    """Single-walker mode with the HOST's random numbers: the device consumes
    them in the reference's draw order and reports how many it used."""
    rng = np.random.default_rng(5)
    u = rng.random(8 * 48 * 25)
    o, up = make_oracle_walker("ice1_sample", overrides={"eq_mc_cycles": 3})
    g, _ = make_gpu_walkers("ice1_sample", overrides={"eq_mc_cycles": 3})
    o.set_rng_fifo(u); g.set_rng_fifo(u)
    g.mc_run(12); assert o.mc_run(12) == 0
    _compare(g, o, up)
    used = g.state().rng_index
    assert used == o.geti("rng_fifo_pos") and 6 * 48 * 12 <= used <= 8 * 48 * 12
    # unconsumed numbers stay queued: continue with the remainder of the same FIFO
    g.mc_run(10); assert o.mc_run(10) == 0
    assert g.state().rng_index == o.geti("rng_fifo_pos")
    np.testing.assert_array_equal(g.download()[0], o.ljr)
    # running dry is an error, not silent reuse
    from mc_water_ls_mw_b200._lib import MwgpuError
    with pytest.raises(MwgpuError):
        g.mc_run(10)


def test_monitor_adjusts_steps_and_resyncs_energy():
    ov = {"monitor_int": 20, "eq_mc_cycles": 1000}                    # eq_adjust_mc = .true. in the deck
    o, up = make_oracle_walker("ice1_sample", overrides=ov)
    g, _ = make_gpu_walkers("ice1_sample", overrides=ov)
    o.set_rng_philox(SEED, 3, 1000000); g.set_rng_philox(SEED, 3, 1000000)
    for _ in range(3):
        g.mc_run(20); assert o.mc_run(20) == 0
        g.mc_monitor(); o.mc_monitor()
        s = g.state()
        assert s.mc_max_trans == o.getd("mc_max_trans") and s.mc_dv_max == o.getd("mc_dv_max")   # bit-exact
        assert list(s.attempted) == [0, 0, 0]
    assert g.state().mc_max_trans < up.mc_max_trans                   # 1.1 Ang is far too large at 200 K
    g.mc_run(15); assert o.mc_run(15) == 0
    _compare(g, o, up)


def test_chain_synchronisation():
    o, up = make_oracle_walker("ice1_sample")
    g, _ = make_gpu_walkers("ice1_sample")
    o.set_rng_philox(SEED, 1, 1000000); g.set_rng_philox(SEED, 1, 1000000)
    g.mc_run(40); assert o.mc_run(40) == 0
    g.mc_chain_sync(); o.mc_chain_sync()
    ljr, ref, hm = g.download()
    np.testing.assert_array_equal(ljr, o.ljr); np.testing.assert_array_equal(hm, o.hmatrix)
    s = g.state()
    assert rel_err(list(s.model_energy), o.model_energy) < TOL and abs(s.ls_mu - o.getd("ls_mu")) < 1e-9
    g.mc_run(10); assert o.mc_run(10) == 0
    _compare(g, o, up)


def test_batch_walkers_equal_independent_single_walkers():
    """W walkers in one launch == W serial runs: walker w uses Philox stream first_stream + w."""
    nw = 6
    ov = {"eq_mc_cycles": 4}
    g, up = make_gpu_walkers("ice1_gen_weights", nwalkers=nw, overrides=ov)
    os_ = make_oracle_walkers("ice1_gen_weights", nw, overrides=ov)
    g.set_rng_philox(SEED, 100, 1000000)
    for w, o in enumerate(os_):
        o.set_rng_philox(SEED, 100 + w, 1000000)
        assert o.mc_run(25) == 0
    g.mc_run(25)
    for w, o in enumerate(os_):
        _compare(g, o, up, walker=w)
    mus = [g.state(w).ls_mu for w in range(nw)]
    assert len(set(mus)) == nw                                        # the walkers really are independent


@pytest.mark.parametrize("ex,chunk,blocks", [
    ("ice1_gen_weights", 3, 5),      # 12 walkers through 5 persistent blocks in units of 3 cycles (Wang-Landau updates)
    ("ice1_sample", 1, 3),           # every cycle a unit: the queue turns over 12 x 20 times
    ("ice1_sample", 7, 0),           # units, but a block per walker
    ("single_box", 2, 4),            # one-lattice instantiation
])
def test_unit_scheduler_never_changes_a_result(ex, chunk, blocks):
    """A launch cut into (walker, chunk of cycles) units taken from the device queue by a few persistent blocks
    (the form every batch larger than the GPU runs in, mw2.cuh) == the serial chains of the oracle, bit for bit,
    including a walker that stops early and the volume moves of the deck."""
    nw, ncyc = 12, 20
    ov = {"eq_mc_cycles": 4, "mc_vol_prob": 0.02}
    g, up = make_gpu_walkers(ex, nwalkers=nw, overrides=ov)
    os_ = make_oracle_walkers(ex, nw, overrides=ov)
    g.set_schedule(chunk, blocks)
    g.set_rng_philox(SEED, 7, 1000000)
    for w, o in enumerate(os_):
        o.set_rng_philox(SEED, 7 + w, 1000000)
        assert o.mc_run(ncyc) == 0
    g.mc_run(9); g.mc_run(ncyc - 9)                                  # two launches, the second one starts mid-chain
    for w, o in enumerate(os_):
        _compare(g, o, up, walker=w)
    t = g.walker_times()
    assert (t[:, 1] > t[:, 0]).all()
    # the automatic schedule gives the same chains
    g2, _ = make_gpu_walkers(ex, nwalkers=nw, overrides=ov)
    g2.set_rng_philox(SEED, 7, 1000000)
    g2.mc_run(ncyc)
    for w in range(nw):
        np.testing.assert_array_equal(g2.download(w)[0], g.download(w)[0])
        assert list(g2.state(w).accepted) == list(g.state(w).accepted)
        assert g2.state(w).model_energy[0] == g.state(w).model_energy[0]


def test_delta_allreduce_of_bins_matches_reference_semantics():
    """comms_allreduce_eta/hist (comms_mpi.f90:244-277,461-493) over the walkers of one context."""
    from oracle import orc
    nw = 5
    ov = {"eq_mc_cycles": 2}
    g, up = make_gpu_walkers("ice1_gen_weights", nwalkers=nw, overrides=ov)
    os_ = make_oracle_walkers("ice1_gen_weights", nw, overrides=ov)
    g.set_rng_philox(SEED, 0, 1000000)
    for w, o in enumerate(os_):
        o.set_rng_philox(SEED, w, 1000000)
    for _ in range(2):
        g.mc_run(10)
        for o in os_:
            assert o.mc_run(10) == 0
        g.comms_allreduce_bins(); orc.allreduce_bins(os_)
        for w, o in enumerate(os_):
            wg, hg, ug = g.bins(w)
            np.testing.assert_allclose(wg, o.weight, rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(hg, o.histogram, rtol=0, atol=1e-9)
        # after a sync every walker holds the same arrays
        np.testing.assert_array_equal(g.bins(0)[0], g.bins(nw - 1)[0])
    for w, o in enumerate(os_):
        _compare(g, o, up, walker=w)


def test_dd_window_walkers():
    """ice1_sample_dd: one walker per mu-window (mc_moves.F90:660-709,915-922); the windows, the
    forced starting lattice and the in-window bookkeeping follow the rank."""
    size = 4
    ov = {"eq_mc_cycles": 100000}
    g, up = make_gpu_walkers("ice1_sample_dd", nwalkers=size, size=size, overrides=ov)
    os_ = make_oracle_walkers("ice1_sample_dd", size, size=size, overrides=ov)
    for w, o in enumerate(os_):
        s = g.state(w)
        assert (s.my_start_bin, s.my_end_bin) == (o.geti("my_start_bin"), o.geti("my_end_bin"))
        assert s.my_mu_min == o.getd("my_mu_min") and s.my_mu_max == o.getd("my_mu_max")
        assert s.ls == o.geti("ls")
        np.testing.assert_array_equal(g.bins(w)[0], o.weight)         # weights outside the window zeroed
    g.set_rng_philox(SEED, 0, 1000000)
    for w, o in enumerate(os_):
        o.set_rng_philox(SEED, w, 1000000)
        assert o.mc_run(30) == 0
    g.mc_run(30)
    for w, o in enumerate(os_):
        _compare(g, o, up, walker=w)
        assert g.state(w).attempted[2] == 0                           # no switches during dd equilibration


def test_dd_window_error_is_reported():
    """mc_moves.F90:187-201: a walker outside its window at eq_mc_cycles stops the run."""
    from mc_water_ls_mw_b200._lib import MwgpuError
    size = 4
    g, up = make_gpu_walkers("ice1_sample_dd", nwalkers=size, size=size, overrides={"eq_mc_cycles": 3})
    g.set_rng_philox(SEED, 0, 1000000)
    with pytest.raises(MwgpuError, match="window"):
        g.mc_run(5)


def test_full_size_properties_4096_walkers():
    """BASELINE configs[4] at full size: properties that need no oracle run.
    (i) incremental energies equal a fresh full evaluation (the reference's drift check,
    mc_moves.F90:1781-1792, 1e-10 Ha); (ii) attempted-move accounting; (iii) two identical
    launches from identical states are bit-identical (determinism)."""
    nw = 4096
    g, up = make_gpu_walkers("ice1_sample", nwalkers=nw, overrides={"eq_mc_cycles": 2})
    g.set_rng_philox(SEED, 0, 1000000)
    g.mc_run(20)
    st = g.states()
    stored = np.array([[s.model_energy[0], s.model_energy[1]] for s in st])
    att = np.array([[s.attempted[0], s.attempted[1], s.attempted[2]] for s in st])
    assert np.all(att[:, 0] + att[:, 1] == 20 * 48) and np.all(att[:, 2] == 20 * 48)
    assert all(s.error == 0 for s in st)
    fresh = g.compute_model_energy_all()
    assert np.max(np.abs(stored - fresh)) < 1e-10
    assert len(np.unique(np.array([s.ls_mu for s in st]))) > nw // 2
    g2, _ = make_gpu_walkers("ice1_sample", nwalkers=nw, overrides={"eq_mc_cycles": 2})
    g2.set_rng_philox(SEED, 0, 1000000)
    g2.mc_run(20)
    np.testing.assert_array_equal(g2.download_all()[0], g.download_all()[0])
    # spot-check one walker of the big batch against the oracle
    o, _ = make_oracle_walker("ice1_sample", overrides={"eq_mc_cycles": 2})
    o.set_rng_philox(SEED, 1234, 1000000)
    assert o.mc_run(20) == 0
    np.testing.assert_array_equal(g.download(1234)[0], o.ljr)


def test_long_run_reproduces_the_reference_energy_drift():
    """The incrementally maintained model_energy drifts from a fresh evaluation when a neighbour enters the
    cut-off before the next list refresh (the reference prints this drift at every monitor, mc_moves.F90:1781-1792).
    That is the algorithm, not an error: the walker with the LARGEST drift out of 768 after 2 500 cycles must be
    reproduced by the oracle -- positions and counters bit for bit, and the same drift."""
    ov = {"eq_mc_cycles": 500}
    nw = 768
    g, up = make_gpu_walkers("ice1_sample", nwalkers=nw, overrides=ov)
    g.set_rng_philox(SEED, 0, 1000000)
    for _ in range(2):
        g.mc_run(1000); g.mc_monitor()
    g.mc_run(500)
    st = g.states()
    assert not any(s.error for s in st)
    inc = np.array([list(s.model_energy) for s in st])
    fresh = g.compute_model_energy_all()
    d = (np.abs(inc - fresh) / np.abs(fresh)).max(1)
    worst = int(np.argmax(d))
    o, _ = make_oracle_walker("ice1_sample", rank=worst, size=nw, overrides=ov)
    o.set_rng_philox(SEED, worst, 1000000)
    for _ in range(2):
        assert o.mc_run(1000) == 0; o.mc_monitor()
    assert o.mc_run(500) == 0
    ljr, ref, hm = g.download(worst)
    np.testing.assert_array_equal(ljr, o.ljr); np.testing.assert_array_equal(hm, o.hmatrix)
    assert list(st[worst].accepted) == [o.geti("acc_r"), o.geti("acc_v"), o.geti("acc_s")]
    assert st[worst].rng_index == o.geti("rng_index")
    oinc = np.array(o.model_energy).copy()
    ofresh = np.array([o.compute_model_energy(1), o.compute_model_energy(2)])
    assert rel_err(inc[worst], oinc) < TOL and rel_err(fresh[worst], ofresh) < TOL
    od = (np.abs(oinc - ofresh) / np.abs(ofresh)).max()
    assert abs(d[worst] - od) <= 1e-9 * max(od, 1e-12) + 1e-13
