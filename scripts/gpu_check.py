"""Quick GPU-vs-oracle diagnostic (development aid; the real checks live in tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests.helpers import make_oracle_walker, load_example
from mc_water_ls_mw_b200 import walkers as W
from oracle import orc

def rel(a, b): return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(1e-300, np.abs(np.asarray(b))))

ex = sys.argv[1] if len(sys.argv) > 1 else "ice1_sample"
ncyc = int(sys.argv[2]) if len(sys.argv) > 2 else 50
o, up = make_oracle_walker(ex)
up2, h, r, wts, wl = load_example(ex)
g = W.WalkerBatch(up.nwater, up.num_lattices, 1)
g.upload(r, h)
g.energy_init()
for l in range(1, up.num_lattices + 1):
    nn, jn, vn = g.get_neighbours(l)
    print("lat", l, "nn eq", np.array_equal(nn, o.nn[l-1]), "jn eq", np.array_equal(jn, o.jn[l-1]), "vn eq", np.array_equal(vn, o.vn[l-1]))
    e = g.compute_model_energy(l)
    print("  E gpu", e, "oracle", o.model_energy[l-1], "rel", rel(e, o.model_energy[l-1]))
    loc = g.compute_local_real_energy_all(l)
    oloc = np.array([o.compute_local_real_energy(i+1, l) for i in range(up.nwater)])
    print("  local rel", rel(loc, oloc))
    n, iv = g.compute_ivects(l)
    print("  nivect", n, "ivect eq", np.array_equal(iv[:n], o.ivect[l-1][:n]))
g.mc_init(W.params_from_user(up), 0, 1, wts, wl)
mu, bw, sc = g.grid()
print("grid eq", np.array_equal(mu, o.mu_bin), np.array_equal(bw, o.binwidth), sc, o.getd("log_unbiased_norm"))
s = g.state()
print("mu0 gpu", s.ls_mu, "oracle", o.getd("ls_mu"))
g.set_rng_philox(20141211, 0, 1000000)
o.set_rng_philox(20141211, 0, 1000000)
t = time.time(); g.mc_run(ncyc); tg = time.time() - t
t = time.time(); o.mc_run(ncyc); to = time.time() - t
s = g.state()
print("gpu time", tg, "oracle time", to)
print("counters gpu", list(s.accepted), list(s.attempted), "ls", s.ls, "rng", s.rng_index)
print("counters orc", o.counters(), "ls", o.geti("ls"), "rng", o.geti("rng_index"))
ljr, ref, hm = g.download()
print("pos bit-exact", np.array_equal(ljr, o.ljr), "ref", np.array_equal(ref, o.ref_ljr), "h", np.array_equal(hm, o.hmatrix))
print("max pos diff", np.max(np.abs(ljr - o.ljr)))
print("E gpu", list(s.model_energy), "orc", o.model_energy, "rel", rel(list(s.model_energy)[:up.num_lattices], o.model_energy))
print("mu gpu", s.ls_mu, "orc", o.getd("ls_mu"))
wg, hg, ug = g.bins()
print("hist max diff", np.max(np.abs(hg - o.histogram)), "uhist rel", rel(ug, np.maximum(o.unbiased_hist, 1e-300)) if o.unbiased_hist.max() > 0 else 0, "weight diff", np.max(np.abs(wg - o.weight)))
for l in range(1, up.num_lattices + 1):
    nn, jn, vn = g.get_neighbours(l)
    print("lists eq after run", l, np.array_equal(nn, o.nn[l-1]), np.array_equal(jn, o.jn[l-1]), np.array_equal(vn, o.vn[l-1]))
print("translations eq", np.array_equal(g.translations(), o.mc_translations))
