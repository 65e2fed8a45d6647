"""Regenerates tests/golden/ from the read-only reference checkout (build container only).

1. copies the reference's example DATA fixtures (input decks, xmol lattices,
   eta_weights.dat) -- these are the BASELINE configs and the only golden data
   the reference ships;
2. writes oracle_vectors.npz: seeded oracle outputs (energies, neighbour lists,
   a short Markov chain) so that `-m "not gpu"` tests can detect any drift of the
   oracle itself, and the GPU box (which has no /root/reference) has fixed vectors.

Run:  python tests/golden/make_fixtures.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/examples"


def copy_examples():
    for ex in sorted(os.listdir(REF)):
        dst = os.path.join(HERE, "examples", ex)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(REF, ex))):
            shutil.copyfile(os.path.join(REF, ex, f), os.path.join(dst, f))
            os.chmod(os.path.join(dst, f), 0o644)


def oracle_vectors():
    sys.path.insert(0, ROOT)
    import numpy as np
    from oracle import orc
    from tests.helpers import make_oracle_walker

    out = {}
    for ex in ("ice1_sample", "single_box", "ice1_gen_weights"):
        w, up = make_oracle_walker(ex)
        out[f"{ex}/energy0"] = np.array(w.model_energy)
        out[f"{ex}/nn"] = np.array(w.nn)
        out[f"{ex}/jn"] = np.array(w.jn)
        out[f"{ex}/vn"] = np.array(w.vn)
        out[f"{ex}/local"] = np.array([[w.compute_local_real_energy(i + 1, l + 1) for i in range(w.nwater)]
                                        for l in range(w.nlat)])
        out[f"{ex}/mu0"] = np.array([w.getd("ls_mu")])
        w.set_rng_philox(20141211, 0, 1000000)
        assert w.mc_run(30) == 0
        out[f"{ex}/ljr30"] = np.array(w.ljr)
        out[f"{ex}/h30"] = np.array(w.hmatrix)
        out[f"{ex}/energy30"] = np.array(w.model_energy)
        out[f"{ex}/mu30"] = np.array([w.getd("ls_mu")])
        out[f"{ex}/counters30"] = np.array([w.geti(k) for k in ("acc_r", "acc_v", "acc_s", "att_r", "att_v", "att_s", "ls")])
        out[f"{ex}/rng_index30"] = np.array([w.geti("rng_index")])
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)


def bookkeeping_vectors():
    """Seeded inputs + oracle outputs of the periodic bookkeeping on the reduced arrays:
    window joins, deltaG, flatness decisions.  The GPU box replays the inputs through the C ABI."""
    sys.path.insert(0, ROOT)
    import numpy as np
    from oracle import orc
    from tests.helpers import make_oracle_walkers

    out = {}
    rng = np.random.default_rng(20141211)
    ws = make_oracle_walkers("ice1_sample_dd", 4, size=4)
    nb = ws[0].nbins
    x = np.linspace(-2, 2, nb)
    U = np.array([np.exp(-x * x) * 2.0 ** w * (1 + 0.01 * rng.standard_normal(nb)) for w in range(4)])
    Wt = np.array([x * x + 3.0 * w + 0.01 * rng.standard_normal(nb) for w in range(4)])
    for w, s in enumerate(ws):
        s.unbiased_hist[:] = U[w]; s.weight[:] = Wt[w]
    out["dd/uhist"] = U; out["dd/weight"] = Wt
    for ov in (0, 2, 5):
        out[f"dd/join_uhist_{ov}"] = orc.join_uhist(ws, ov)
        out[f"dd/join_eta_{ov}"] = orc.join_eta(ws, ov)
    dg, normP = orc.mc_deltaG_from_hist(ws)
    out["dd/deltaG"] = np.array([dg]); out["dd/normP"] = normP
    ws = make_oracle_walkers("ice1_sample", 3)
    inc = rng.random((3, nb))
    for s, u in zip(ws, inc):
        s.unbiased_hist[:] = u
    dg, normP = orc.mc_deltaG_from_hist(ws)
    out["mw/uhist_increments"] = inc; out["mw/deltaG"] = np.array([dg]); out["mw/normP"] = normP
    # flatness: (schedule, histogram) -> (flat, mean, max_pct, min_pct, wl_factor, weights after)
    ws0 = make_oracle_walkers("ice1_gen_weights", 1)
    nbg = ws0[0].nbins
    hists = np.array([np.full(nbg, 100.0), np.r_[np.full(nbg - 1, 100.0), 111.0], np.r_[np.full(nbg - 1, 500.0), 19.4],
                      np.r_[np.full(nbg - 1, 100.0), 400.0], 50.0 + 100.0 * rng.random(nbg)])
    wts = np.linspace(3.0, 7.0, nbg)
    rows = []; wafter = []
    for sched in (0, 1, 2):
        for h in hists:
            s = make_oracle_walkers("ice1_gen_weights", 1)[0]
            # state as after a restart with an already reduced increment (firstcycle = .false.): the GPU side
            # reaches it through mwgpu_mc_restore
            s.seti("firstcycle", 0); s.seti("mc_cycle_num", 100); s.setd("wl_factor", 0.004)
            s.weight[:] = wts; s.histogram[:] = h; s.arr_d("hist_last_sync", (nbg,))[:] = h
            r = orc.mc_check_flatness([s], sched, 20, float(np.float32(0.05)), False)
            rows.append([sched, r.checked, r.hist_reset, r.flat, r.mean, r.max_pct, r.min_pct, r.wl_factor])
            wafter.append(s.weight.copy())
    out["flat/hists"] = hists; out["flat/weights"] = wts
    out["flat/rows"] = np.array(rows); out["flat/weights_after"] = np.array(wafter)
    np.savez_compressed(os.path.join(HERE, "bookkeeping_vectors.npz"), **out)


if __name__ == "__main__":
    if os.path.isdir(REF):
        copy_examples()
    oracle_vectors()
    bookkeeping_vectors()
    print("fixtures written to", HERE)
