"""GPU parity of the device-side periodic bookkeeping against the oracle:
mc_check_flatness, mc_compute_deltaG_from_hist, comms_join_uhist / comms_join_eta, and whole runs under
the reference's event schedule (mc_moves.F90:257-316) with nothing downloaded in between.

Bar: decisions, counters, positions, histograms bit-exact (the sums run in the reference's order with
uncontracted arithmetic); weights / energies / deltaG within 1e-11 (exp/log differ by an ulp)."""
import numpy as np
import pytest

from mc_water_ls_mw_b200.schedule import CycleSchedule
from oracle import orc
from tests.helpers import OracleSchedule, make_gpu_walkers, make_oracle_walkers, rel_err

pytestmark = pytest.mark.gpu
SEED = 20141211


def _pair(ex, n, ov=None):
    g, up = make_gpu_walkers(ex, nwalkers=n, overrides=ov or {})
    ws = make_oracle_walkers(ex, n, overrides=ov or {})
    g.set_rng_philox(SEED, 0, 1000000)
    for i, s in enumerate(ws):
        s.set_rng_philox(SEED, i, 1000000)
    return g, ws, up


def _same_report(a, b):
    assert (a.checked, a.hist_reset, a.flat, a.invt_switched) == (b.checked, b.hist_reset, b.flat, b.invt_switched)
    assert a.mean == b.mean and a.max_pct == b.max_pct and a.min_pct == b.min_pct
    assert abs(a.wl_factor - b.wl_factor) <= 1e-15 * abs(b.wl_factor)


def _same_bins(g, ws, wtol=1e-11):
    for w, s in enumerate(ws):
        wt, h, u = g.bins(w)
        np.testing.assert_allclose(h, s.histogram, rtol=0, atol=1e-9)
        np.testing.assert_allclose(wt, s.weight, rtol=wtol, atol=wtol)
        np.testing.assert_allclose(u, s.unbiased_hist, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("sched,minhist,useinvt", [(0, 20, False), (1, -1, False), (2, 20, True), (1, 0, True)])
def test_check_flatness_matches_oracle(sched, minhist, useinvt):
    g, ws, up = _pair("ice1_gen_weights", 5, {"eq_mc_cycles": 2})
    reports = []
    for it in range(4):                                             # guard -> first reset -> flat / not flat
        g.mc_run(12); assert orc.mc_run_many(ws, 12, 4) == 0
        a = g.mc_check_flatness(sched, minhist, up.wl_flattol, useinvt)
        b = orc.mc_check_flatness(ws, sched, minhist, up.wl_flattol, useinvt)
        _same_report(a, b)
        reports.append((a.hist_reset, a.flat, a.invt_switched))
        _same_bins(g, ws)
        for w, s in enumerate(ws):
            assert abs(g.state(w).wl_factor - s.getd("wl_factor")) <= 1e-15 * s.getd("wl_factor")
    if (sched, minhist) == (1, -1):                                 # one-off reset, then a halving at every check
        assert reports == [(1, 0, 0), (0, 1, 0), (0, 1, 0), (0, 1, 0)]
    if useinvt:
        assert reports[0][2] == 1                                   # switched to the 1/t increment
    g.mc_run(5); assert orc.mc_run_many(ws, 5, 4) == 0              # and the chain continues identically
    for w, s in enumerate(ws):
        st = g.state(w)
        assert list(st.accepted) == [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s")]
        assert st.rng_index == s.geti("rng_index")


def test_check_flatness_is_a_noop_for_sample_runs_and_single_boxes():
    g, ws, up = _pair("ice1_sample", 2, {"eq_mc_cycles": 2})
    g.mc_run(10)
    before = g.bins(0)
    rep = g.mc_check_flatness()
    assert rep.checked == 0
    for x, y in zip(before, g.bins(0)):
        np.testing.assert_array_equal(x, y)
    g1, _ = make_gpu_walkers("single_box", nwalkers=1)
    assert g1.mc_check_flatness().checked == 0


def test_deltaG_from_hist_matches_oracle_mw():
    g, ws, up = _pair("ice1_sample", 6, {"eq_mc_cycles": 3})
    g.mc_run(40); assert orc.mc_run_many(ws, 40, 4) == 0
    dg, npg = g.mc_compute_deltaG_from_hist()
    do, npo = orc.mc_deltaG_from_hist(ws)
    assert abs(dg - do) < 1e-11 * max(1.0, abs(do))
    np.testing.assert_allclose(npg, npo, rtol=1e-9, atol=1e-300)
    _same_bins(g, ws)                                               # the uhist all-reduce re-based every walker
    with pytest.raises(Exception):
        make_gpu_walkers("ice1_gen_weights", nwalkers=1)[0].mc_compute_deltaG_from_hist()   # not a sample run (:305)


def test_window_joins_and_deltaG_dd():
    g, ws, up = _pair("ice1_sample_dd", 4, {"eq_mc_cycles": 3})
    nb = g.nbins
    rng = np.random.default_rng(11)
    x = np.linspace(-2, 2, nb)
    for w, s in enumerate(ws):                                      # synthetic window contents, same on both sides
        u = np.exp(-x * x) * (2.0 ** w) * (1 + 0.01 * rng.standard_normal(nb))
        wt = x * x + 3.0 * w + 0.01 * rng.standard_normal(nb)
        s.unbiased_hist[:] = u; s.weight[:] = wt
        g.set_bins(w, weight=wt, unbiased_hist=u)
    for ov in (0, 2, 4):
        np.testing.assert_allclose(g.comms_join_uhist(ov), orc.join_uhist(ws, ov), rtol=1e-13)
        np.testing.assert_allclose(g.comms_join_eta(ov), orc.join_eta(ws, ov), rtol=1e-13, atol=1e-13)
    dg, npg = g.mc_compute_deltaG_from_hist()
    do, npo = orc.mc_deltaG_from_hist(ws)
    assert abs(dg - do) < 1e-11 * max(1.0, abs(do))
    np.testing.assert_allclose(npg, npo, rtol=1e-12)
    with pytest.raises(Exception):
        g.comms_join_eta(40)                                        # windows narrower than the overlap


@pytest.mark.parametrize("ex,n,ncyc,ov", [
    # weight generation: syncs, monitors, flatness checks (one-off reset, then halvings), chain syncs
    ("ice1_gen_weights", 6, 64, {"eq_mc_cycles": 4, "mpi_sync_int": 8, "monitor_int": 16, "flat_chk_int": 16,
                                 "latt_sync_int": 24, "wl_schedule": 1, "wl_minhist": 0}),
    # monitor_int not a multiple of mpi_sync_int: the monitor's own all-reduce (mc_moves.F90:1813-1821) is a real merge
    ("ice1_gen_weights", 5, 45, {"eq_mc_cycles": 4, "mpi_sync_int": 8, "monitor_int": 12, "flat_chk_int": 20,
                                 "latt_sync_int": 30, "wl_minhist": 2}),
    # sampling with fixed weights: syncs + deltaG estimates
    ("ice1_sample", 4, 48, {"eq_mc_cycles": 4, "mpi_sync_int": 8, "monitor_int": 16, "flat_chk_int": 16,
                            "latt_sync_int": 24, "deltaG_int": 24}),
    # dd windows: no all-reduce, per-window flatness
    ("ice1_sample_dd", 4, 90, {"eq_mc_cycles": 50, "monitor_int": 16, "flat_chk_int": 20, "latt_sync_int": 30,
                               "deltaG_int": 45}),
    # dd weight generation: every window checks its own histogram and halves its own wl_factor
    ("ice1_gen_weights_dd", 4, 90, {"eq_mc_cycles": 50, "monitor_int": 16, "flat_chk_int": 10, "latt_sync_int": 30,
                                    "wl_schedule": 1, "wl_minhist": -1}),
])
def test_whole_run_under_the_reference_schedule(ex, n, ncyc, ov):
    g, ws, up = _pair(ex, n, ov)
    sch = CycleSchedule(g, up)
    osch = OracleSchedule(ws, up, nthreads=4)
    for chunk in (ncyc // 2, ncyc - ncyc // 2):                     # cutting the run must not matter
        sch.run(chunk); osch.run(chunk)
    assert [c for c, _ in sch.log.flatness] == [c for c, _ in osch.flatness]
    for (_, a), (_, b) in zip(sch.log.flatness, osch.flatness):
        _same_report(a, b)
    assert len(sch.log.deltaG) == len(osch.deltaG)
    for (c1, d1, p1), (c2, d2, p2) in zip(sch.log.deltaG, osch.deltaG):
        assert c1 == c2
        # (sparsely filled windows give log(0) seams: inf / nan must come out the same way on both sides)
        np.testing.assert_allclose(d1, d2, rtol=1e-10, atol=1e-10, equal_nan=True)
        np.testing.assert_allclose(p1, p2, rtol=1e-9, atol=1e-300, equal_nan=True)
    if ex.startswith("ice1_gen_weights") and ov.get("wl_schedule") == 1:
        assert any(r.flat for _, r in sch.log.flatness)             # wl_factor was halved on the device
    for w, s in enumerate(ws):
        st = g.state(w)
        ljr, ref, hm = g.download(w)
        np.testing.assert_array_equal(ljr, s.ljr); np.testing.assert_array_equal(hm, s.hmatrix)
        assert list(st.accepted) == [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s")]
        assert st.rng_index == s.geti("rng_index") and st.ls == s.geti("ls")
        assert rel_err(list(st.model_energy), s.model_energy) < 1e-11
        assert abs(st.wl_factor - s.getd("wl_factor")) <= 1e-15 * max(s.getd("wl_factor"), 1e-300)
    _same_bins(g, ws)


def test_statistical_consistency_of_histograms_and_deltaG():
    """north_star: 'overlap histograms and free-energy differences statistically consistent'.
    Different random streams on the two sides: visit frequencies of the order parameter, acceptance
    ratios and the deltaG estimate must agree within their statistical errors."""
    ov = {"eq_mc_cycles": 20}
    ncyc, n_o, n_g = 220, 16, 512
    ws = make_oracle_walkers("ice1_sample", n_o, overrides=ov)
    for i, s in enumerate(ws):
        s.set_rng_philox(SEED, 100000 + i, 1000000)
    assert orc.mc_run_many(ws, ncyc, 0) == 0
    g, up = make_gpu_walkers("ice1_sample", nwalkers=2 * n_g, overrides=ov)
    g.set_rng_philox(SEED, 0, 1000000)
    g.mc_run(ncyc)
    st = g.states()
    acc_g = np.array([s.accepted[0] / s.attempted[0] for s in st])
    acc_o = np.array([s.geti("acc_r") / s.geti("att_r") for s in ws])
    # acceptance ratio: walker-to-walker scatter gives the error of the two means
    err = np.hypot(acc_g.std() / np.sqrt(len(acc_g)), acc_o.std() / np.sqrt(len(acc_o)))
    assert abs(acc_g.mean() - acc_o.mean()) < 5 * err + 1e-4
    # visit histogram of mu (per-walker increments, no sync was called): mean bin index and spread
    nb = g.nbins
    hg = np.array([g.bins(w)[1] for w in range(0, 2 * n_g, 8)])
    ho = np.array([s.histogram.copy() for s in ws])
    k = np.arange(nb)
    mg = (hg * k).sum(1) / hg.sum(1); mo = (ho * k).sum(1) / ho.sum(1)
    err = np.hypot(mg.std() / np.sqrt(len(mg)), mo.std() / np.sqrt(len(mo)))
    assert abs(mg.mean() - mo.mean()) < 5 * err + 0.05
    # overlap histograms: fraction of visits in coarse groups of bins, walker-to-walker scatter as the error
    lo, hi = np.nonzero(ho.sum(0))[0][[0, -1]]
    edges = np.linspace(lo, hi + 1, 6).astype(int)
    for a, b in zip(edges[:-1], edges[1:]):
        fg = hg[:, a:b].sum(1) / hg.sum(1); fo = ho[:, a:b].sum(1) / ho.sum(1)
        err = np.hypot(fg.std() / np.sqrt(len(fg)), fo.std() / np.sqrt(len(fo)))
        assert abs(fg.mean() - fo.mean()) < 5 * err + 2e-3, (a, b, fg.mean(), fo.mean(), err)
    # free-energy path: the unbiased-histogram estimate from two disjoint halves of the GPU walkers
    # (independent streams) agrees within the walker-to-walker scatter of its log-ratio of the two basins
    ug = np.array([g.bins(w)[2] for w in range(0, 2 * n_g, 2)])
    mid = nb // 2
    def basin_logratio(u):
        tot = u.sum(0)
        return np.log(tot[:mid].sum() / tot[mid:].sum())
    half = len(ug) // 2
    ra, rb = basin_logratio(ug[:half]), basin_logratio(ug[half:])
    # jackknife over blocks of walkers for the error of each half
    def jk(u, nblk=8):
        blk = np.array_split(np.arange(len(u)), nblk)
        vals = np.array([basin_logratio(np.delete(u, b, axis=0)) for b in blk])
        return np.sqrt((nblk - 1) / nblk * ((vals - vals.mean()) ** 2).sum())
    err = np.hypot(jk(ug[:half]), jk(ug[half:]))
    assert np.isfinite(ra) and np.isfinite(rb)
    assert abs(ra - rb) < 5 * err + 0.05, (ra, rb, err)


def test_tree_reduction_of_large_batches_matches_a_host_sum():
    """More than 256 walkers take the one-CTA-per-bin reduction (deterministic, not the serial order)."""
    n = 300
    g, up = make_gpu_walkers("ice1_gen_weights", nwalkers=n, overrides={"eq_mc_cycles": 1})
    g.set_rng_philox(SEED, 0, 1000000)
    g.mc_run(4)
    per = [g.bins(w) for w in range(n)]
    w0 = per[0][0].copy()                                       # weights start equal, so the base is anything minus its own delta
    g.comms_allreduce_bins()
    base_w = np.zeros_like(w0)                                  # eta_last_sync = the (zero) start weights, hist_last_sync = 0
    tot_w = base_w + np.sum([p[0] - base_w for p in per], axis=0)
    tot_h = np.sum([p[1] for p in per], axis=0)
    for w in (0, 1, n - 1):
        wg, hg, _ = g.bins(w)
        np.testing.assert_allclose(wg, tot_w, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(hg, tot_h, rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(g.bins(0)[0], g.bins(n - 1)[0])
    g.comms_allreduce_bins()                                    # nothing new since the sync: a fixed point
    np.testing.assert_allclose(g.bins(5)[0], tot_w, rtol=1e-12, atol=1e-12)


def test_bookkeeping_golden_vectors_on_gpu():
    """The committed inputs of tests/golden/bookkeeping_vectors.npz replayed through the C ABI."""
    import os
    from tests.helpers import GOLDEN
    v = np.load(os.path.join(GOLDEN, "bookkeeping_vectors.npz"))
    g, up = make_gpu_walkers("ice1_sample_dd", nwalkers=4, size=4)
    for w in range(4):
        g.set_bins(w, weight=v["dd/weight"][w], unbiased_hist=v["dd/uhist"][w])
    for ov in (0, 2, 5):
        np.testing.assert_allclose(g.comms_join_uhist(ov), v[f"dd/join_uhist_{ov}"], rtol=1e-13)
        np.testing.assert_allclose(g.comms_join_eta(ov), v[f"dd/join_eta_{ov}"], rtol=1e-13, atol=1e-13)
    dg, normP = g.mc_compute_deltaG_from_hist()
    assert abs(dg - v["dd/deltaG"][0]) < 1e-11 * max(1.0, abs(v["dd/deltaG"][0]))
    np.testing.assert_allclose(normP, v["dd/normP"], rtol=1e-12)
    g, up = make_gpu_walkers("ice1_sample", nwalkers=3)
    for w in range(3):
        g.set_bins(w, unbiased_hist=v["mw/uhist_increments"][w])    # increments over the zero base
    dg, normP = g.mc_compute_deltaG_from_hist()
    assert abs(dg - v["mw/deltaG"][0]) < 1e-11 * max(1.0, abs(v["mw/deltaG"][0]))
    np.testing.assert_allclose(normP, v["mw/normP"], rtol=1e-12)
    k = 0
    for sched in (0, 1, 2):
        for h in v["flat/hists"]:
            g, up = make_gpu_walkers("ice1_gen_weights", nwalkers=1)
            rec = g.checkpoint_record(0)
            rec.update(mc_cycle_num=100, wl_factor=0.004, histogram=h, weight=v["flat/weights"])
            g.mc_restore(0, rec)                                     # firstcycle = .false., bases re-set
            r = g.mc_check_flatness(sched, 20, float(np.float32(0.05)), False)
            row = v["flat/rows"][k]
            assert [sched, r.checked, r.hist_reset, r.flat] == row[:4].tolist()
            assert [r.mean, r.max_pct, r.min_pct, r.wl_factor] == row[4:].tolist()       # bit for bit
            np.testing.assert_array_equal(g.bins(0)[0], v["flat/weights_after"][k])
            k += 1
