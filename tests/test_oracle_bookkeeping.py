"""Oracle restatement of the periodic bookkeeping on the reduced arrays
checked against independent numpy restatements of the reference formulae and hand-made cases:
mc_check_flatness (mc_moves.F90:1936-2185), mc_compute_deltaG_from_hist (:2498-2621),
comms_join_uhist / comms_join_eta (comms_mpi.f90:299-459).  No GPU."""
import numpy as np

from oracle import orc
from tests.helpers import make_oracle_walkers

F32_TOL = float(np.float32(0.05))


def _walkers(n, ex="ice1_gen_weights", ov=None):
    return make_oracle_walkers(ex, n, overrides=ov or {})


def _set_hist(ws, h):
    for s in ws:
        s.histogram[:] = h
        s.arr_d("hist_last_sync", (s.nbins,))[:] = h          # as after a sync: no pending increments


def test_flatness_guard_and_first_reset():
    ws = _walkers(3)
    nb = ws[0].nbins
    rep = orc.mc_check_flatness(ws)                            # empty histogram: :1961 returns
    assert rep.checked == 0
    _set_hist(ws, np.full(nb, 30.0))                           # every bin > wl_minhist = 20 on the first f
    ws[0].weight[:] = np.arange(nb, dtype=float)
    rep = orc.mc_check_flatness(ws, wl_minhist=20)
    assert rep.checked == 1 and rep.hist_reset == 1 and rep.flat == 0
    for s in ws:
        assert not s.histogram.any() and not s.arr_d("hist_last_sync", (nb,)).any()
        assert s.geti("histogram_reset") == 1 and s.geti("firstcycle") == 1
    assert ws[0].getd("wl_factor") == ws[0].getd("orig_wl_factor")                       # unchanged


def test_flatness_schedules_and_weight_shift():
    for sched, hist_fn, want in [
        (0, lambda nb: np.full(nb, 100.0), 1),
        (0, lambda nb: np.r_[np.full(nb - 1, 100.0), 100.0 * (1 + 2 * F32_TOL)], 0),
        (1, lambda nb: np.r_[np.full(nb - 1, 500.0), 20.4], 1),      # nint(20.4) = 20 >= wl_minhist
        (1, lambda nb: np.r_[np.full(nb - 1, 500.0), 19.4], 0),
        (2, lambda nb: np.r_[np.full(nb - 1, 100.0), 400.0], 1),     # only bins BELOW (1-tol)*mean break flatness
        (2, lambda nb: np.r_[np.full(nb - 1, 100.0), 10.0], 0),
    ]:
        ws = _walkers(2)
        nb = ws[0].nbins
        for s in ws:
            s.seti("histogram_reset", 1)                          # past the one-off reset
            s.seti("mc_cycle_num", 100)
            s.weight[:] = np.linspace(3.0, 7.0, nb)
        _set_hist(ws, hist_fn(nb))
        f0 = ws[0].getd("wl_factor")
        w_before = ws[0].weight.copy()
        rep = orc.mc_check_flatness(ws, wl_schedule=sched, wl_minhist=20, wl_flattol=F32_TOL)
        assert rep.checked == 1 and rep.flat == want, (sched, want)
        if want:
            for s in ws:
                assert s.getd("wl_factor") == 0.5 * f0 and s.geti("firstcycle") == 0
                assert not s.histogram.any()
                np.testing.assert_array_equal(s.weight, w_before - w_before[nb // 2])      # weight(nbins/2+1) -> 0
        else:
            assert ws[0].getd("wl_factor") == f0
            np.testing.assert_array_equal(ws[0].weight, w_before)
        # the report is the log line of :1996-1997
        h = hist_fn(nb)
        np.testing.assert_allclose(rep.mean, h.sum() / nb, rtol=1e-13)
        np.testing.assert_allclose(rep.max_pct, 100.0 * h.max() / (h.sum() / nb), rtol=1e-13)


def test_flatness_reduces_histogram_increments_over_ranks():
    ws = _walkers(3)
    nb = ws[0].nbins
    for r, s in enumerate(ws):
        s.seti("histogram_reset", 1); s.seti("mc_cycle_num", 50)
        s.histogram[:] = 10.0 * (r + 1)                            # increments since the last sync (base 0)
    rep = orc.mc_check_flatness(ws, wl_schedule=1, wl_minhist=1000)     # not flat: nothing is reset
    assert rep.flat == 0
    for s in ws:
        np.testing.assert_array_equal(s.histogram, np.full(nb, 60.0))   # comms_allreduce_hist
    assert rep.mean == 60.0


def test_switch_to_inverse_time():
    ws = _walkers(1)
    s = ws[0]
    nb = s.nbins
    s.seti("histogram_reset", 1); s.seti("mc_cycle_num", 10)
    s.setd("wl_factor", 1e-4)
    _set_hist(ws, np.r_[np.full(nb - 1, 100.0), 1.0])              # not flat
    rep = orc.mc_check_flatness(ws, wl_useinvt=True)
    assert rep.flat == 0 and rep.invt_switched == 1 and s.geti("wl_invt_active") == 1
    assert s.getd("wl_factor") == nb / (10.0 * s.nwater)           # :2134
    rep = orc.mc_check_flatness(ws, wl_useinvt=True)               # 1/t active: no flatness test any more
    assert rep.checked == 1 and rep.flat == 0 and rep.invt_switched == 0


def _np_join(arrs, overlap, eta):
    """numpy restatement of comms_join_eta / comms_join_uhist (comms_mpi.f90:299-459)."""
    size, nb = arrs.shape
    bpw = nb // size
    joined = arrs[0].copy()
    for ir in range(1, size):
        e = ir * bpw
        sl = slice(e - overlap - 1, e + overlap)
        if eta:
            shift = joined[sl].sum() / (2 * overlap + 1) - arrs[ir][sl].sum() / (2 * overlap + 1)
            joined[e:] = arrs[ir][e:] + shift
        else:
            shift = np.log(joined[sl]).sum() / (2 * overlap + 1) - np.log(arrs[ir][sl]).sum() / (2 * overlap + 1)
            joined[e:] = arrs[ir][e:] * np.exp(shift)
    if eta:
        joined = joined - joined[nb // 2]
    return joined


def test_window_joins_match_numpy_restatement():
    rng = np.random.default_rng(7)
    ws = _walkers(4, ex="ice1_sample_dd")
    nb = ws[0].nbins
    x = np.linspace(-2, 2, nb)
    truth = np.exp(-x * x)                                         # a smooth unbiased histogram
    for r, s in enumerate(ws):
        s.unbiased_hist[:] = truth * (3.0 ** r) * (1 + 0.01 * rng.standard_normal(nb))     # windows differ by a scale
        s.weight[:] = x * x + 5.0 * r + 0.01 * rng.standard_normal(nb)                     # ... weights by an offset
    U = np.array([s.unbiased_hist.copy() for s in ws]); Wt = np.array([s.weight.copy() for s in ws])
    for ov in (0, 2, 5):
        np.testing.assert_allclose(orc.join_uhist(ws, ov), _np_join(U, ov, False), rtol=1e-13)
        np.testing.assert_allclose(orc.join_eta(ws, ov), _np_join(Wt, ov, True), rtol=1e-13, atol=1e-13)
    # the stitched histogram is continuous across the seams although the windows are not
    j = orc.join_uhist(ws, 2)
    bpw = nb // 4
    for ir in (1, 2, 3):
        assert abs(np.log(j[ir * bpw] / j[ir * bpw - 1]) - np.log(truth[ir * bpw] / truth[ir * bpw - 1])) < 0.1


def test_deltaG_from_hist_formula():
    ws = _walkers(3, ex="ice1_sample")
    nb = ws[0].nbins
    rng = np.random.default_rng(3)
    incs = [rng.random(nb) for _ in ws]
    for s, inc in zip(ws, incs):
        s.unbiased_hist[:] = inc                                    # increments over a zero base
    bw = ws[0].binwidth.copy()
    dG, normP = orc.mc_deltaG_from_hist(ws)
    tot = incs[0] + incs[1] + incs[2]
    np.testing.assert_allclose(normP, tot / (tot * bw).sum(), rtol=1e-14)
    pA = (normP[: nb // 2] * bw[: nb // 2]).sum(); pB = (normP[nb // 2:] * bw[nb // 2:]).sum()
    assert abs(dG - np.log(pA / pB)) < 1e-13
    assert abs((normP * bw).sum() - 1.0) < 1e-13
    for s in ws:                                                    # comms_allreduce_uhist happened (:2530)
        np.testing.assert_allclose(s.unbiased_hist, tot, rtol=1e-15)


def test_bookkeeping_golden_vectors():
    """tests/golden/bookkeeping_vectors.npz (tests/golden/make_fixtures.py): the oracle must keep reproducing it."""
    import os
    from tests.helpers import GOLDEN
    v = np.load(os.path.join(GOLDEN, "bookkeeping_vectors.npz"))
    ws = _walkers(4, ex="ice1_sample_dd")
    for w, s in enumerate(ws):
        s.unbiased_hist[:] = v["dd/uhist"][w]; s.weight[:] = v["dd/weight"][w]
    for ov in (0, 2, 5):
        np.testing.assert_array_equal(orc.join_uhist(ws, ov), v[f"dd/join_uhist_{ov}"])
        np.testing.assert_array_equal(orc.join_eta(ws, ov), v[f"dd/join_eta_{ov}"])
    dg, normP = orc.mc_deltaG_from_hist(ws)
    assert dg == v["dd/deltaG"][0]; np.testing.assert_array_equal(normP, v["dd/normP"])
    ws = _walkers(3, ex="ice1_sample")
    for s, u in zip(ws, v["mw/uhist_increments"]):
        s.unbiased_hist[:] = u
    dg, normP = orc.mc_deltaG_from_hist(ws)
    assert dg == v["mw/deltaG"][0]; np.testing.assert_array_equal(normP, v["mw/normP"])
    k = 0
    for sched in (0, 1, 2):
        for h in v["flat/hists"]:
            s = _walkers(1)[0]
            s.seti("firstcycle", 0); s.seti("mc_cycle_num", 100); s.setd("wl_factor", 0.004)
            s.weight[:] = v["flat/weights"]; s.histogram[:] = h; s.arr_d("hist_last_sync", (s.nbins,))[:] = h
            r = orc.mc_check_flatness([s], sched, 20, F32_TOL, False)
            assert [sched, r.checked, r.hist_reset, r.flat, r.mean, r.max_pct, r.min_pct, r.wl_factor] == v["flat/rows"][k].tolist()
            np.testing.assert_array_equal(s.weight, v["flat/weights_after"][k])
            k += 1
    assert v["flat/rows"][:, 3].sum() >= 3                       # several flat and several non-flat cases
