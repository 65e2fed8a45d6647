"""Wave-tail probe of the walker kernel: per-walker start / end times of a 250-cycle launch over 4096 walkers, with
one unit per walker (chunk 250: waves of whole walkers) and with the launch cut into units of 16 / 8 / 4 cycles.  Development aid.
usage: python scripts/diag/tail_probe.py [walkers]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, bench
from mc_water_ls_mw_b200 import walkers as W
up, h, r, w, wl = bench._example()
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = W.WalkerBatch(up.nwater, up.num_lattices, nw)
g.upload(r, h); g.energy_init()
g.mc_init(W.params_from_user(up), 0, nw, w, wl)
g.set_rng_philox(20141211, 0, 1000000)
for _ in range(2):
    g.mc_run(500); g.mc_monitor()
for mode in (250, 32, 16, 8, 4, 16, 8, 32):
    g.set_schedule(mode, 0)
    ts = []
    for i in range(6):
        g.timer_start(); g.mc_run_async(250); ms = g.timer_stop(); ts.append(ms)
    t = g.walker_times().astype(np.int64)
    t0 = t[:, 0].min(); s = (t[:, 0] - t0) * 1e-6; e = (t[:, 1] - t0) * 1e-6; d = e - s
    total = e.max()
    # resident walkers over time: fraction of the launch spent below 90 % / 50 % of the peak residency
    ev = np.concatenate([np.stack([s, np.ones_like(s)], 1), np.stack([e, -np.ones_like(e)], 1)])
    ev = ev[np.argsort(ev[:, 0], kind='stable')]
    occ = np.cumsum(ev[:, 1]); dt = np.diff(np.append(ev[:, 0], total))
    peak = occ.max()
    below90 = dt[occ < 0.9 * peak].sum(); below50 = dt[occ < 0.5 * peak].sum()
    area = (occ * dt).sum() / (peak * total)
    print(f"chunk {mode}: step ms " + " ".join(f"{x:.1f}" for x in ts) + f" | span {total:.1f} ms, walker ms min {d.min():.1f} mean {d.mean():.1f} "
          f"max {d.max():.1f} sd {d.std():.1f} | peak resident {int(peak)}, below 90% {below90:.1f} ms, below 50% {below50:.1f} ms, residency area {area:.3f}")
    first = d[s < 1.0]; second = d[s >= 1.0]
    print(f"   first wave {len(first)} walkers mean {first.mean():.1f} ms; later {len(second)} mean {second.mean() if len(second) else 0:.1f} ms; "
          f"last start {s.max():.1f} ms")
