"""GPU parity for SURVEY.md 8(f) row 3: therm rows recorded by the walker kernel (main.f90:200-223) and the
checkpoint -> restart path (mc_moves.F90:324-501, :842-862) against the oracle."""
import numpy as np
import pytest

from mc_water_ls_mw_b200 import decks
from tests.helpers import make_gpu_walkers, make_oracle_walkers, rel_err

pytestmark = pytest.mark.gpu
SEED = 20141211


class _Row:
    def __init__(self, **kw):
        self.__dict__.update(kw)


@pytest.mark.parametrize("ex", ["ice1_sample", "ice1_gen_weights", "single_box"])
def test_therm_rows_match_oracle(ex):
    ov = {"eq_mc_cycles": 3, "file_output_int": 5}               # the decks say 50
    g, up = make_gpu_walkers(ex, nwalkers=3, overrides=ov)
    ws = make_oracle_walkers(ex, 3, overrides=ov)
    g.set_rng_philox(SEED, 0, 1000000)
    for i, s in enumerate(ws):
        s.set_rng_philox(SEED, i, 1000000)
    g.set_therm(up.file_output_int, capacity=16)
    want = [[] for _ in ws]
    for chunk in (7, 16):                                        # launches longer than the output interval
        g.mc_run(chunk)
        for i, s in enumerate(ws):
            for _ in range(chunk):
                assert s.mc_run(1) == 0
                c = s.geti("mc_cycle_num")
                if c % up.file_output_int == 0:
                    want[i].append(_Row(icyc=c, ls=s.geti("ls"), model_energy=list(s.model_energy) + [0.0],
                                        ls_mu=s.getd("ls_mu"), volume=list(s.volume) + [0.0], hmatrix1=list(s.hmatrix[0])))
    for i in range(3):
        rows, dropped = g.therm(i)
        assert dropped == 0 and [r.icyc for r in rows] == [r.icyc for r in want[i]] == [5, 10, 15, 20]
        for a, b in zip(rows, want[i]):
            assert a.ls == b.ls and list(a.hmatrix1) == b.hmatrix1
            nl = up.num_lattices
            assert rel_err(list(a.model_energy)[:nl], b.model_energy[:nl]) < 1e-11
            assert list(a.volume)[:nl] == b.volume[:nl]
            if nl == 2:
                assert abs(a.ls_mu - b.ls_mu) < 1e-9 * max(1.0, abs(b.ls_mu))
            assert decks.format_therm_row(a, up) == decks.format_therm_row(b, up)      # the line the reference writes
        assert g.therm(i) == ([], 0)                             # drained


def test_therm_ring_reports_dropped_rows():
    g, up = make_gpu_walkers("ice1_sample", nwalkers=1)
    g.set_rng_philox(SEED, 0, 1000000)
    g.set_therm(2, capacity=3)
    g.mc_run(12)                                                 # 6 rows wanted, 3 fit
    rows, dropped = g.therm(0)
    assert [r.icyc for r in rows] == [2, 4, 6] and dropped == 3
    g.set_therm(0, 0)                                            # recording off
    g.mc_run(4)
    assert g.therm(0) == ([], 0)


@pytest.mark.parametrize("ex", ["ice1_gen_weights", "ice1_sample", "single_box"])
def test_checkpoint_restart_continues_like_the_oracle(ex, tmp_path):
    ov = {"eq_mc_cycles": 4, "mc_vol_prob": 0.02}
    g, up = make_gpu_walkers(ex, nwalkers=2, overrides=ov)
    ws = make_oracle_walkers(ex, 2, overrides=ov)
    g.set_rng_philox(SEED, 0, 1000000)
    for i, s in enumerate(ws):
        s.set_rng_philox(SEED, i, 1000000)
    g.mc_run(17)
    for s in ws:
        assert s.mc_run(17) == 0
    # --- write the checkpoints from the device state, in the reference's file format
    recs = []
    for w in range(2):
        p = str(tmp_path / f"checkpoint{w:03d}.dat.1")
        decks.write_checkpoint(p, g.checkpoint_record(w), g.nbins, up.samplerun)
        recs.append(decks.read_checkpoint(p, g.nbins, up.num_lattices, up.samplerun))
        np.testing.assert_array_equal(recs[w]["ljr"], ws[w].ljr)
        assert recs[w]["mc_cycle_num"] == 17 and recs[w]["ls"] == ws[w].geti("ls")
    # --- a fresh start-up (input configuration, energy_init, mc_init) followed by the restart path, both sides
    g2, _ = make_gpu_walkers(ex, nwalkers=2, overrides=ov)
    ws2 = make_oracle_walkers(ex, 2, overrides=ov)
    for w in range(2):
        g2.mc_restore(w, recs[w])
        ws2[w].mc_restore(recs[w])
    g2.set_rng_philox(SEED, 50, 777)
    for i, s in enumerate(ws2):
        s.set_rng_philox(SEED, 50 + i, 777)
    for w, s in enumerate(ws2):                                  # state right after the restart
        st = g2.state(w)
        assert st.mc_cycle_num == 17 and st.ls == s.geti("ls")
        assert rel_err(list(st.model_energy)[: up.num_lattices], s.model_energy) < 1e-11
        assert abs(st.ls_mu - s.getd("ls_mu")) < 1e-9 * max(1.0, abs(s.getd("ls_mu")))
        ljr, ref, hm = g2.download(w)
        np.testing.assert_array_equal(ljr, s.ljr); np.testing.assert_array_equal(ref, s.ref_ljr)
        np.testing.assert_array_equal(hm, s.hmatrix)
    g2.mc_run(13)
    for w, s in enumerate(ws2):
        assert s.mc_run(13) == 0
        st = g2.state(w)
        ljr, ref, hm = g2.download(w)
        np.testing.assert_array_equal(ljr, s.ljr); np.testing.assert_array_equal(hm, s.hmatrix)
        assert list(st.accepted) == [s.geti("acc_r"), s.geti("acc_v"), s.geti("acc_s")]
        assert st.mc_cycle_num == 30 and st.rng_index == s.geti("rng_index")
        assert rel_err(list(st.model_energy)[: up.num_lattices], s.model_energy) < 1e-11
        if up.num_lattices == 2:
            wt, h, u = g2.bins(w)
            np.testing.assert_allclose(h, s.histogram, rtol=0, atol=1e-9)
            np.testing.assert_allclose(wt, s.weight, rtol=1e-11, atol=1e-11)


def test_restart_continues_the_random_stream_and_rejected_restore_leaves_the_walker_alone():
    """The reference's checkpoint carries no generator state (mc_moves.F90:353-379; it re-seeds from the clock): a
    restart has to advance the Philox stream itself -- mwgpu_mc_set_rng_index with mwgpu_mc_get_state().rng_index of
    the checkpoint -- or it replays the numbers the first segment consumed.  And a restore whose arguments are
    rejected must not touch the walker."""
    from mc_water_ls_mw_b200._lib import MwgpuError
    from tests.helpers import make_oracle_walker
    ov = {"eq_mc_cycles": 4}
    first, up = make_gpu_walkers("ice1_sample", nwalkers=1, overrides=ov)
    first.set_rng_philox(SEED, 3, 1000000)
    first.mc_run(9)
    rec = first.checkpoint_record(0)
    idx = first.state(0).rng_index
    assert idx > 1000000 + 9 * 48 * 6
    g2, _ = make_gpu_walkers("ice1_sample", nwalkers=1, overrides=ov)
    g2.set_rng_philox(SEED, 3, 1000000)
    before = g2.download(0)[0].copy()
    bad = dict(rec); bad["ls"] = 7
    with pytest.raises(MwgpuError):
        g2.mc_restore(0, bad)
    np.testing.assert_array_equal(g2.download(0)[0], before)      # untouched by the rejected call
    g2.mc_restore(0, rec)
    assert g2.state(0).rng_index == 1000000                      # restore does not know where the stream was ...
    g2.set_rng_index(idx, walker=0)                               # ... the caller says so
    o, _ = make_oracle_walker("ice1_sample", overrides=ov)
    o.mc_restore(rec)
    o.set_rng_philox(SEED, 3, idx)
    g2.mc_run(6); assert o.mc_run(6) == 0
    np.testing.assert_array_equal(g2.download(0)[0], o.ljr)
    assert g2.state(0).rng_index == o.geti("rng_index") > idx and list(g2.state(0).accepted)[0] > 0
