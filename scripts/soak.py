import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from tests.helpers import make_gpu_walkers
from mc_water_ls_mw_b200.schedule import CycleSchedule
for ex in ("ice1_sample", "ice1_gen_weights"):
    g, up = make_gpu_walkers(ex, nwalkers=2048, overrides={"eq_mc_cycles": 500, "deltaG_int": 2000, "flat_chk_int": 1000, "latt_sync_int": 2500})
    g.set_rng_philox(20141211, 0, 1000000)
    g.set_therm(up.file_output_int, capacity=64)
    sch = CycleSchedule(g, up)
    t0 = time.time(); sch.run(6500); dt = time.time() - t0
    st = g.states()
    errs = [s.error for s in st if s.error]
    inc = np.array([list(s.model_energy) for s in st])
    fresh = g.compute_model_energy_all()
    drift = np.max(np.abs(inc - fresh) / np.abs(fresh))
    acc = np.mean([s.accepted[0] / max(s.attempted[0], 1) for s in st])
    print(ex, "cycles 6500 walkers 2048 in %.1f s; errors %d; max rel drift incremental vs fresh energy %.2e; acc ratio %.3f; ls=2 fraction %.3f" % (dt, len(errs), drift, acc, np.mean([s.ls == 2 for s in st])))
    print("  flatness events:", [(c, r.checked, r.hist_reset, r.flat, round(r.wl_factor, 6)) for c, r in sch.log.flatness][:8])
    print("  deltaG estimates (kT per molecule):", [(c, round(d / up.nwater, 5)) for c, d, _ in sch.log.deltaG])
    rows, dropped = g.therm(0)
    print("  therm rows walker 0:", len(rows), "dropped", dropped, "| last:", __import__("mc_water_ls_mw_b200.decks", fromlist=["x"]).format_therm_row(rows[-1], up) if rows else None)
