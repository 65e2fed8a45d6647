"""Generates tests/golden/numpy_vectors.npz from the independent numpy restatement (ref_numpy.py): lists, full and
local energies of both lattices, and the state after 3 MC cycles (accept / reject sequence, positions, counters,
bins) of ice1_sample, single_box, ice1_gen_weights and of single windows of the two domain-decomposed decks
under a host FIFO of random numbers.  tests/test_oracle_numpy.py holds the C
oracle to these vectors.

    python tests/golden/make_fixtures_numpy.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mc_water_ls_mw_b200 import decks          # input readers only (namelists, xmol, eta_weights: host I/O)
from tests.golden import ref_numpy as R

CASES = {
    # key: (deck, overrides, cycles, rng seed, rank, size)
    "ice1_sample": ("ice1_sample", {"eq_mc_cycles": 1, "mc_vol_prob": 0.04, "list_update_int": 2}, 3, 101, 0, 1),
    "single_box": ("single_box", {"eq_mc_cycles": 1, "mc_vol_prob": 0.04, "list_update_int": 2}, 3, 202, 0, 1),
    "ice1_gen_weights": ("ice1_gen_weights", {"eq_mc_cycles": 1, "list_update_int": 2}, 2, 303, 0, 1),
    # domain decomposition over the order parameter (mc_moves.F90:660-703): window 1 of 4 never holds the walker
    # (equilibration phase: no weights, no switches, no bins); window 3 of 4 holds it from the start (production
    # phase from cycle 2: in-window weights, out-of-window rejections, switches, unbiased histogram); the weight
    # generation deck updates the window's weights only (:1682-1685)
    "ice1_sample_dd@0of4": ("ice1_sample_dd", {"eq_mc_cycles": 100, "mc_vol_prob": 0.04, "list_update_int": 2}, 3, 404, 0, 4),
    "ice1_sample_dd@2of4": ("ice1_sample_dd", {"eq_mc_cycles": 2, "mc_vol_prob": 0.04, "list_update_int": 2}, 3, 505, 2, 4),
    "ice1_gen_weights_dd@2of4": ("ice1_gen_weights_dd", {"eq_mc_cycles": 2, "list_update_int": 2}, 3, 606, 2, 4),
    # branches of the switch / the weights the decks do not take by default
    "ice1_sample/nvt": ("ice1_sample", {"eq_mc_cycles": 1, "mc_ensemble": "nvt", "list_update_int": 2}, 2, 1101, 0, 1),
    "ice1_sample/leshift": ("ice1_sample", {"eq_mc_cycles": 1, "leshift": True, "mc_vol_prob": 0.04}, 2, 1202, 0, 1),
    "ice1_sample/no_interp": ("ice1_sample", {"eq_mc_cycles": 1, "eta_interp": False, "mc_vol_prob": 0.04}, 2, 1303, 0, 1),
    "ice1_gen_weights/switch_prob": ("ice1_gen_weights", {"eq_mc_cycles": 1, "mc_always_switch": False, "mc_switch_prob": 0.3,
                                                          "mc_vol_prob": 0.04}, 2, 1404, 0, 1),
    "ice1_gen_weights/swetnam": ("ice1_gen_weights", {"eq_mc_cycles": 1, "wl_swetnam": True}, 2, 1505, 0, 1),
}
# periodic bookkeeping that changes the walker's state, between two stretches of cycles:
# key: (deck, overrides, cycles before, event, cycles after, rng seed)
EVENT_CASES = {
    # mc_monitor_stats (mc_moves.F90:1722-1732, :1786-1810) in the equilibration phase: step sizes re-tuned, stored
    # energies replaced by fresh ones, counters reset -- the second stretch runs on all of that
    "ice1_sample+monitor": ("ice1_sample", {"eq_mc_cycles": 100, "mc_vol_prob": 0.06, "list_update_int": 2}, 2, "monitor", 2, 707),
    "single_box+monitor": ("single_box", {"eq_mc_cycles": 100, "mc_vol_prob": 0.06, "list_update_int": 2}, 2, "monitor", 2, 808),
    # mc_check_chain_synchronisation (:2217-2416) after volume moves have let the two cells drift apart
    "ice1_sample+chain_sync": ("ice1_sample", {"eq_mc_cycles": 1, "mc_vol_prob": 0.06, "list_update_int": 2}, 2, "chain_sync", 2, 910),
}


def load(name, ov, size=1):
    d = os.path.join(HERE, "examples", name)
    up = decks.read_input(os.path.join(d, "ice.input"), size=size)        # io.f90:249: one rank = no overlap
    for k, v in ov.items():
        setattr(up, k, v)
    h, r = decks.read_config(d, up)
    wl, w = 0.0, None
    p = os.path.join(d, "eta_weights.dat")
    if up.num_lattices == 2 and os.path.exists(p):
        wl, _, w = decks.read_eta_weights(p)
    return up, h, r, w, wl


def main():
    out = {}
    for name, (deck, ov, ncyc, seed, rank, size) in CASES.items():
        up, h, r, w, wl = load(deck, ov, size)
        b = R.Box(up, h, r, weights=w, file_wl_factor=wl, rank=rank, size=size)
        nl, N = b.nlat, b.N
        out[f"{name}/nn"] = np.array(b.nn, dtype=np.int32)
        out[f"{name}/jn"] = np.array(b.jn, dtype=np.int32)
        out[f"{name}/vn"] = np.array(b.vn, dtype=np.int32)
        out[f"{name}/energy0"] = np.array(b.model_energy[:nl])
        out[f"{name}/local0"] = np.array([[b.compute_local_real_energy(i, l) for i in range(N)] for l in range(nl)])
        out[f"{name}/mu0"] = np.array([b.ls_mu])
        out[f"{name}/mu_bin"] = np.array(b.mu_bin)
        out[f"{name}/binwidth"] = np.array(b.binwidth)
        out[f"{name}/scalars"] = np.array([b.r_pos, b.r_neg, b.av_binwidth, b.log_unbiased_norm])
        out[f"{name}/window"] = np.array([b.my_start_bin, b.my_end_bin, b.my_mu_min, b.my_mu_max, b.ls])
        out[f"{name}/weight0"] = np.array(b.weight)
        u = np.random.default_rng(seed).random(8 * N * ncyc + 16)
        out[f"{name}/fifo"] = u
        b.set_fifo(u)
        for _ in range(ncyc):
            b.mc_cycle()
        out[f"{name}/ncycles"] = np.array([ncyc])
        out[f"{name}/ljr"] = np.array(b.r)
        out[f"{name}/ref_ljr"] = np.array(b.ref)
        out[f"{name}/hmatrix"] = b.hflat()
        out[f"{name}/counters"] = np.array(b.acc + b.att + [b.ls, b.fpos, b.mc_cycle_num], dtype=np.int64)
        out[f"{name}/trace"] = np.array(b.trace, dtype=np.int8)
        out[f"{name}/energy"] = np.array(b.model_energy[:nl])
        out[f"{name}/mu"] = np.array([b.ls_mu])
        out[f"{name}/volume"] = np.array(b.volume[:nl])
        out[f"{name}/average_energy"] = np.array(b.average_energy[:nl])
        out[f"{name}/histogram"] = np.array(b.histogram)
        out[f"{name}/weight"] = np.array(b.weight)
        out[f"{name}/unbiased_hist"] = np.array(b.unbiased_hist)
        out[f"{name}/mc_translations"] = np.array(b.mc_translations, dtype=np.int32)
        out[f"{name}/nn_end"] = np.array(b.nn, dtype=np.int32)
        tr = np.array(b.trace)
        print(f"{name}: E0 = {b.model_energy[:nl]}, accepted {b.acc}, attempted {b.att}, draws {b.fpos}, "
              f"volume moves {(tr[:, 0] == 1).sum()}")
    for name, (deck, ov, n1, event, n2, seed) in EVENT_CASES.items():
        up, h, r, w, wl = load(deck, ov)
        b = R.Box(up, h, r, weights=w, file_wl_factor=wl)
        nl, N = b.nlat, b.N
        u = np.random.default_rng(seed).random(8 * N * (n1 + n2) + 16)
        out[f"{name}/fifo"] = u
        b.set_fifo(u)
        for _ in range(n1):
            b.mc_cycle()
        out[f"{name}/pre_counters"] = np.array(b.acc + b.att, dtype=np.int64)
        if event == "monitor":
            b.mc_monitor_stats()
        else:
            b.mc_check_chain_synchronisation()
        out[f"{name}/mid_ljr"] = np.array(b.r)
        out[f"{name}/mid_hmatrix"] = b.hflat()
        out[f"{name}/mid_energy"] = np.array(b.model_energy[:nl])
        out[f"{name}/mid_mu"] = np.array([b.ls_mu])
        out[f"{name}/mid_steps"] = np.array([b.mc_max_trans, b.mc_dv_max])
        for _ in range(n2):
            b.mc_cycle()
        out[f"{name}/ncycles"] = np.array([n1, n2])
        out[f"{name}/ljr"] = np.array(b.r)
        out[f"{name}/ref_ljr"] = np.array(b.ref)
        out[f"{name}/hmatrix"] = b.hflat()
        out[f"{name}/counters"] = np.array(b.acc + b.att + [b.ls, b.fpos, b.mc_cycle_num], dtype=np.int64)
        out[f"{name}/energy"] = np.array(b.model_energy[:nl])
        out[f"{name}/mu"] = np.array([b.ls_mu])
        out[f"{name}/volume"] = np.array(b.volume[:nl])
        out[f"{name}/average_energy"] = np.array(b.average_energy[:nl])
        out[f"{name}/mc_translations"] = np.array(b.mc_translations, dtype=np.int32)
        tr = np.array(b.trace)
        print(f"{name}: pre {out[f'{name}/pre_counters'].tolist()}, steps {b.mc_max_trans:.6f} {b.mc_dv_max:.6f}, accepted {b.acc}, "
              f"attempted {b.att}, draws {b.fpos}, volume moves {(tr[:, 0] == 1).sum()}")
    np.savez_compressed(os.path.join(HERE, "numpy_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "numpy_vectors.npz"))


if __name__ == "__main__":
    main()
