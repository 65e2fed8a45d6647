"""Shared helpers for the tests (oracle side)."""
from __future__ import annotations

import os

import numpy as np

from mc_water_ls_mw_b200 import decks
from oracle import orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXAMPLES = os.path.join(GOLDEN, "examples")


def example_dir(name: str) -> str:
    return os.path.join(EXAMPLES, name)


def load_example(name: str, size: int = 1):
    """(UserParams, hmatrix[nlat,9], ljr[nlat,N,3], weights-or-None, file_wl_factor)."""
    d = example_dir(name)
    up = decks.read_input(os.path.join(d, "ice.input"), size=size)
    h, r = decks.read_config(d, up)
    wpath = os.path.join(d, "eta_weights.dat")
    if up.num_lattices == 2 and os.path.exists(wpath):
        wl, _, w = decks.read_eta_weights(wpath)
    else:
        wl, w = 0.0, None
    return up, h, r, w, wl


def make_oracle_walker(name: str, rank: int = 0, size: int = 1, overrides: dict | None = None):
    up, h, r, w, wl = load_example(name, size=size)
    for k, v in (overrides or {}).items():
        setattr(up, k, v)
    s = orc.System(up.nwater, up.num_lattices)
    s.set_config(r, h)
    s.energy_init()
    for ils in range(1, up.num_lattices + 1):      # main.f90:125-128
        s.compute_model_energy(ils)
    rc = s.mc_init(orc.params_from_user(up), rank=rank, size=size, weights=w, file_wl_factor=wl)
    assert rc == 0
    return s, up
