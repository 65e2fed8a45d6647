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


if __name__ == "__main__":
    if os.path.isdir(REF):
        copy_examples()
    oracle_vectors()
    print("fixtures written to", HERE)
