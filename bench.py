#!/usr/bin/env python
"""Benchmark of the hot path: attempted MC moves/s (whole box) and full mW energy evals/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], "synthetic scale-out"): 4096 independent
lattice-switch walkers PER GPU of the ice1 size (48 mW molecules per lattice,
cubic <-> hexagonal ice, deck + weights of examples/ice1_sample: 200 K, 1 atm,
samplerun, nbins 101, list_update_int 10).  Weak scaling by default: every rank
owns 4096 walkers (a B200 holds 2072 walkers resident; 4096 = two full waves);
the only exchange is the delta all-reduce of weights / histograms every
mpi_sync_int = 250 cycles (NCCL).  `--scaling strong` splits 4096 walkers in total
over the GPUs instead (512 per GPU on 8: a quarter of the residency, reported for
completeness in profiles/README.md).

A *step* = one call of the hot path over the whole batch between two exchanges of
the reference: `mc_run(mpi_sync_int = 250 cycles)` = 4096 x 48 x 250 attempted moves
per GPU, including the in-kernel neighbour-list rebuilds (every 10th cycle), followed
by the delta all-reduce of weights / histograms (comms_mpi.f90:244-277).  `value` is
device-timed with inputs resident in HBM; `e2e` is the same step through the C ABI
with HOST buffers (pinned host -> device upload of every walker's positions /
reference positions / cells, list + energy re-initialisation as after a checkpoint
load, the 250 cycles, and the device -> host read of positions and observables)
inside the timed region.

Both arms run the PRODUCTION phase of the deck: eq_mc_cycles is set to the 1000 untimed
preparation cycles + 1 (the deck's own 10 000 equilibration cycles would leave the histogram
updates of mc_update_wl_bins out of every timed step).

At N > 1 the line also carries `"strong"`: BASELINE configs[4] read literally (4096 walkers
IN TOTAL, split evenly over the GPUs), measured in the same run on a second batch, and
`"allreduce_check"`: the library's NCCL delta all-reduce against a host sum of the per-rank
increments and across ranks.

`--impl reference`: the reference cannot be compiled here (no Fortran compiler, no
MPI in this image or on the GPU box); the arm times the CPU oracle (C restatement of
the same algorithm, oracle/mw_oracle.c) built with -O3 -march=native -ffp-contract=fast
on the machine it runs on (the parity tests keep the strict -O2 -ffp-contract=off build),
on all host cores, same config and metric.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WALKERS_PER_GPU = 4096
CYCLES_PER_STEP = 250            # = mpi_sync_int of the example decks: one step ends with the delta all-reduce
FLOP_PER_MOVE = 10730.0          # BASELINE.md section 3 / SURVEY.md 8(d): average attempted LS move
FLOP_PER_EVAL = 35424.0          # one lattice full energy (mean of 34560 / 36288)
BYTES_PER_EVAL = 8336.0          # one lattice, reference int32 list layout (mean of 8144 / 8528)
EXAMPLE = "ice1_sample"
SEED = 20141211
PREP_CYCLES = 1000               # untimed equilibration: 2 x (500 cycles + monitor with eq_adjust_mc); ~30 volume-move
                                 # attempts per monitor, enough for the reference's acceptance-ratio tuning of the step sizes
                                 # (mc_moves.F90:1722-1732) to be more than noise; eq_mc_cycles = PREP_CYCLES + 1
ENERGY_REPLICAS = 8              # full-energy batch: the rank's decorrelated walkers x 8 = 65 536 evaluations per launch


def _example():
    from mc_water_ls_mw_b200 import decks
    d = os.path.join(ROOT, "tests", "golden", "examples", os.environ.get("MW_BENCH_EXAMPLE", EXAMPLE))
    up = decks.read_input(os.path.join(d, "ice.input"))
    up.eq_mc_cycles = PREP_CYCLES + 1    # production phase: histogram / unbiased-histogram updates inside the timed steps
    h, r = decks.read_config(d, up)
    wl, _, w = decks.read_eta_weights(os.path.join(d, "eta_weights.dat"))
    return up, h, r, w, wl


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (pynvml)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _config(nw: int, total: int, C: int, nwater: int = 48, cpu: bool = False):
    """The workload both arms are quoted on (BASELINE.json configs[4])."""
    work = {} if cpu else {"walkers_per_gpu": nw, "walkers_total": total, "moves_per_step": total * nwater * C}
    return {
        "workload": f"synthetic scale-out (BASELINE configs[4]): {nw} independent lattice-switch walkers per GPU, "
                    f"{EXAMPLE} deck (48 mW molecules per lattice, cubic<->hexagonal ice, 200 K, 1 atm, fixed weights), "
                    f"production phase after {PREP_CYCLES} equilibration cycles",
        **work, "cycles_per_step": C,
        "l2_policy": "walker state (~75 MB for 4096 walkers) lives in global memory; a walker's 15 KB image is loaded into shared "
                     "memory for every turn on a persistent block (every 8 cycles unless the walker is behind the batch's average "
                     "progress, when the batch exceeds the resident blocks) and stored back; the hot loop runs out of shared "
                     "memory, so cache state between steps does not matter",
        "rng": "Philox-4x32-10, one stream per walker",
    }


def csrc_sha256() -> str:
    """Fingerprint of the CUDA sources the committed ncu capture (profiles/traffic.json) was taken on."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "mc_water_ls_mw_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()


def _traffic_stale() -> bool:
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("csrc_sha256") != csrc_sha256()
    except Exception:
        return True


def _traffic(kernel: str, cycles: int, walkers: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture
    (profiles/traffic.json, written by scripts/prof_final.sh together with the sha256 of csrc/ it was taken on:
    `traffic_stale` in the bench line says whether the sources changed since); only valid for the launch shape
    it was taken on."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        if cycles == CYCLES_PER_STEP and walkers == WALKERS_PER_GPU:
            return float(t[kernel]["dram_bytes"])
    except Exception:
        pass
    return None


def _issue(kernel: str, cycles: int, walkers: int, k_ms: float, sm_mhz, n_sm: int = 148):
    """Warp-instruction issue rate of the dominant kernel: executed warp-instructions of one launch (committed ncu
    capture of the same launch shape) / (live launch duration x SM clock under load x SMs), against the 4 issue
    slots per cycle of an SM.  This, not the FP64 pipe, is the resource the walker kernel saturates."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        if cycles != CYCLES_PER_STEP or walkers != WALKERS_PER_GPU or not sm_mhz:
            return None
        inst = float(t[kernel]["warp_instructions"])
        ipc = inst / (k_ms * 1e-3 * sm_mhz * 1e6 * n_sm)
        return {"warp_instructions_per_launch": inst, "ipc_per_sm": ipc, "peak_ipc_per_sm": 4.0, "frac": ipc / 4.0,
                "instructions_per_move": inst / (walkers * 48 * cycles)}
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on the host cores
# --------------------------------------------------------------------------------------------------
def _oracle_walkers(n, up, h, r, w, wl, first_stream=0, fast=True):
    from oracle import orc
    ws = []
    for i in range(n):
        s = orc.System(up.nwater, up.num_lattices, fast=fast)
        s.set_config(r, h)
        s.energy_init()
        s.mc_init(orc.params_from_user(up), rank=i, size=max(n, 1), weights=w, file_wl_factor=wl)
        s.set_rng_philox(SEED, first_stream + i, 1000000)
        ws.append(s)
    return ws


def _check_fast_build(up, h, r, w, wl):
    """The timing build (-O3 -march=native, FMA contraction) against the strict parity build: full and local
    energies of a thermalised configuration to 1e-11 before anything is timed."""
    from oracle import orc
    a = _oracle_walkers(1, up, h, r, w, wl, fast=False)[0]
    assert a.mc_run(30) == 0
    b = orc.System(up.nwater, up.num_lattices, fast=True)
    b.set_config(np.array(a.ljr), np.array(a.hmatrix)); b.energy_init()
    worst = 0.0
    for l in range(1, up.num_lattices + 1):
        ea, eb = a.compute_model_energy(l), b.compute_model_energy(l)
        worst = max(worst, abs(ea - eb) / abs(ea))
        for i in (1, 17, up.nwater):
            la, lb = a.compute_local_real_energy(i, l), b.compute_local_real_energy(i, l)
            worst = max(worst, abs(la - lb) / abs(la))
    if not worst < 1e-11:
        raise SystemExit(f"bench.py: the -O3 oracle build disagrees with the strict build ({worst:.3g})")
    return worst


def _decorrelate_oracle(ws, nthreads):
    """Same untimed preparation as the GPU arm: 2 x (500 cycles + monitor with eq_adjust_mc)."""
    from oracle import orc
    for _ in range(2):
        assert orc.mc_run_many(ws, PREP_CYCLES // 2, nthreads) == 0
        for s in ws:
            s.mc_monitor()


def _oracle_step(ws, nsync, nthreads):
    """nsync x (mpi_sync_int cycles on all host threads + the delta merge of comms_mpi.f90:244-277)."""
    from oracle import orc
    for _ in range(nsync):
        assert orc.mc_run_many(ws, CYCLES_PER_STEP, nthreads) == 0
        orc.allreduce_bins(ws)


def _cpu_flags():
    from oracle import orc
    return f"gcc {orc.FAST_FLAGS} (timed) / {orc.STRICT_FLAGS} (parity tests)"


def cpu_baseline(target_seconds: float = 12.0):
    """Oracle (timing build) on all host cores on a bounded sample of the same workload."""
    from oracle import orc
    up, h, r, w, wl = _example()
    agree = _check_fast_build(up, h, r, w, wl)
    nthreads = orc.max_threads()
    ws = _oracle_walkers(nthreads * 2, up, h, r, w, wl)
    _decorrelate_oracle(ws, nthreads)
    t0 = time.perf_counter(); _oracle_step(ws, 1, nthreads); dt = time.perf_counter() - t0
    nsync = max(1, int(target_seconds / max(dt, 1e-6)))
    t0 = time.perf_counter(); _oracle_step(ws, nsync, nthreads); dt = time.perf_counter() - t0
    ncyc = nsync * CYCLES_PER_STEP
    moves = len(ws) * up.nwater * ncyc
    # energy evaluations (single lattice evals / s)
    t0 = time.perf_counter(); reps = 0
    while time.perf_counter() - t0 < 2.0:
        orc.model_energy_many(ws, nthreads); reps += 1
    evals = reps * len(ws) * 2 / (time.perf_counter() - t0)
    # one walker on one thread: the stand-in for the reference's serial build (COMMS_ARCH=serial)
    one = ws[:1]
    t0 = time.perf_counter(); ncs = 0
    while time.perf_counter() - t0 < 2.0:
        assert orc.mc_run_many(one, 50, 1) == 0; ncs += 50
    serial = up.nwater * ncs / (time.perf_counter() - t0)
    return {
        "value": moves / dt, "unit": "attempted MC moves/s", "cores": nthreads, "kind": "port",
        "sample": f"{len(ws)} walkers x {ncyc} cycles of {EXAMPLE}, bins merged every {CYCLES_PER_STEP} cycles (oracle "
                  f"restatement, not the Fortran binary; {dt:.1f} s on {nthreads} threads; {_cpu_flags()}; "
                  f"timing build vs strict build on energies: {agree:.1e})",
        "energy_evals_per_s": evals,
        "serial": {"value": serial, "unit": "attempted MC moves/s", "cores": 1,
                   "sample": f"1 walker x {ncs} cycles on one thread (stand-in for the reference's serial build)"},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    up, h, r, w, wl = _example()
    agree = _check_fast_build(up, h, r, w, wl)
    nthreads = orc.max_threads()
    ws = _oracle_walkers(nthreads * 2, up, h, r, w, wl)
    _decorrelate_oracle(ws, nthreads)
    # size one step to ~3 s so that K+W steps end within minutes
    t0 = time.perf_counter(); _oracle_step(ws, 1, nthreads); dt = time.perf_counter() - t0
    nsync = max(1, int(3.0 / max(dt, 1e-6)))
    per_step = nsync * CYCLES_PER_STEP
    for _ in range(args.warmup):
        _oracle_step(ws, nsync, nthreads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _oracle_step(ws, nsync, nthreads)
    dt = time.perf_counter() - t0
    moves = len(ws) * up.nwater * per_step * args.steps
    val = moves / dt
    unit = "attempted MC moves/s"
    sample = (f"{len(ws)} walkers x {per_step} cycles per step of {EXAMPLE} on {nthreads} host threads, bins merged every "
              f"{CYCLES_PER_STEP} cycles (CPU oracle, {_cpu_flags()}; timing build vs strict build on energies: {agree:.1e})")
    print(json.dumps({
        "impl": "reference", "metric": "attempted MC moves/sec (whole box)", "value": val, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload as the GPU arm; each CPU step is the bounded sample of it described in cpu_sample
        "config": dict(_config(WALKERS_PER_GPU, WALKERS_PER_GPU * max(args.gpus, 1), CYCLES_PER_STEP, up.nwater, cpu=True),
                       cpu_sample={"walkers": len(ws), "cycles_per_step": per_step,
                                   "moves_per_step": len(ws) * up.nwater * per_step}),
        "cpu_baseline": {"value": val, "unit": unit, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mc_water_ls_mw_b200 import comms, walkers as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # weak scaling (default): args.walkers per GPU; strong: args.walkers in total, split evenly over the ranks
    # (BASELINE configs[4] read literally: "4096 walkers sharded over 1/2/4/8 B200")
    strong = args.scaling == "strong"
    if strong and args.walkers % world:
        raise SystemExit("bench.py: --scaling strong needs --walkers divisible by the number of GPUs")
    nw = args.walkers // world if strong else args.walkers
    total = nw * world
    up, h, r, w, wl = _example()

    def make_batch(n, first_global, size, tag):
        b_ = W.WalkerBatch(up.nwater, up.num_lattices, n, device=local)
        b_.upload(r, h)
        b_.energy_init()
        b_.mc_init(W.params_from_user(up), first_rank=first_global, size=size, weights=w, file_wl_factor=wl)
        b_.set_rng_philox(SEED, first_global, 1000000)
        if world > 1:
            comms.init_nccl(b_, rank, world)
        # untimed equilibration with the reference's step-size adjustment
        for _ in range(2):
            b_.mc_run(PREP_CYCLES // 2)
            b_.mc_monitor()
        if world > 1:
            b_.comms_allreduce_bins()   # untimed: the first NCCL collective of a communicator sets up its channels
        return b_

    C = args.cycles
    sync_every = max(1, up.mpi_sync_int // C)

    def barrier(b_):
        b_.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step(b_, i):
        b_.mc_run_async(C)
        if (i + 1) % sync_every == 0:
            b_.comms_allreduce_bins()

    def timed(b_, steps, warmup):
        for i in range(warmup):
            step(b_, i)
        barrier(b_)
        l0 = b_.kernel_launches()
        b_.timer_start()
        for i in range(steps):
            step(b_, i)
        ms_ = b_.timer_stop()
        barrier(b_)
        launches_ = b_.kernel_launches() - l0
        if world > 1:
            t = torch.tensor([ms_], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_, launches_

    g = make_batch(nw, rank * nw, total, "main")
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = timed(g, args.steps, args.warmup)
    clocks = sampler.finish()
    moves_per_step = total * up.nwater * C
    value = moves_per_step * args.steps / (ms * 1e-3)

    # ---- dominant kernel (k_mc_run2) timed live with CUDA events on its own stream
    kms = []
    for i in range(min(args.steps, 10)):
        g.mc_run(C)
        kms.append(g.last_kernel_ms())
    k_ms = float(np.mean(kms))
    fp64_peak = W.measure_fp64_peak(local)
    achieved_tf = nw * up.nwater * C * FLOP_PER_MOVE / (k_ms * 1e-3) / 1e12
    peaks, peak_kind = _peaks()

    # ---- the N > 1 extras: NCCL delta all-reduce checked against a host sum; BASELINE's strong shape
    allreduce_check = None
    strong_leg = None
    if world > 1:
        g.comms_allreduce_bins()                                   # sync point: every walker holds the merged arrays
        w0, h0, u0 = g.bins(0)
        g.mc_run(10)
        ptr, n = g.comms_reduce_local()                            # this rank's summed increments [3][nbins padded]
        loc = torch.as_tensor(comms._DevBuf(ptr, n), device=torch.device("cuda", local)).clone()
        parts = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc)
        host_total = np.sum(np.stack([p_.cpu().numpy() for p_ in parts]), axis=0).reshape(-1, n // (3 if up.samplerun else 2))
        g.comms_allreduce_bins()                                   # the library's own NCCL all-reduce + re-base
        w1, h1, u1 = g.bins(0)
        nb = len(w1)
        d_host = max(float(np.max(np.abs((w1 - w0) - host_total[0][:nb]))), float(np.max(np.abs((h1 - h0) - host_total[1][:nb]))),
                     float(np.max(np.abs((u1 - u0) - host_total[2][:nb]))) if up.samplerun else 0.0)
        mine = torch.tensor(np.concatenate([w1, h1, u1]), device="cuda", dtype=torch.float64)
        alls = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(alls, mine)
        d_ranks = max(float((a_ - alls[0]).abs().max().item()) for a_ in alls)
        wl_, hl_, ul_ = g.bins(nw - 1)
        allreduce_check = {"max_abs_diff_across_ranks": d_ranks,
                           "max_abs_diff_first_vs_last_walker": float(max(np.max(np.abs(wl_ - w1)), np.max(np.abs(hl_ - h1)), np.max(np.abs(ul_ - u1)))),
                           "vs_host_sum": d_host, "histogram_increment_total": float(np.sum(h1 - h0)),
                           "note": "10 cycles of increments: library NCCL all-reduce vs numpy sum of the per-rank partial sums"}
        if not strong and args.walkers % world == 0:
            nws = args.walkers // world
            g2 = make_batch(nws, rank * nws, args.walkers, "strong")
            ms2, _ = timed(g2, args.steps, args.warmup)
            v2 = args.walkers * up.nwater * C * args.steps / (ms2 * 1e-3)
            strong_leg = {"value": v2, "unit": "attempted MC moves/s", "ms_per_step": ms2 / args.steps,
                          "walkers_total": args.walkers, "walkers_per_gpu": nws,
                          "speedup_vs_one_gpu": v2 / (value / world), "efficiency_vs_one_gpu": v2 / value,
                          "note": "BASELINE configs[4] read literally: 4096 walkers in total, split evenly; efficiency = "
                                  "value / (N x the per-GPU rate of the weak leg of this run, i.e. of 4096 walkers on one GPU)"}
            del g2

    # ---- full mW energy evals / s (second half of the metric): batched kernel on 65 536 evaluations per launch
    # (the rank's decorrelated walkers x ENERGY_REPLICAS, device resident; > 126 MB of state: not L2 resident)
    ljr, ref, hm = g.download_all()
    rep = ENERGY_REPLICAS
    big = W.WalkerBatch(up.nwater, up.num_lattices, nw * rep, device=local)
    big.upload_all(np.ascontiguousarray(np.tile(ljr, (rep, 1, 1, 1))), np.ascontiguousarray(np.tile(hm, (rep, 1, 1))))
    big.energy_init()
    e_out = np.empty((nw * rep, up.num_lattices))
    for _ in range(3):
        big.compute_model_energy_all(e_out)
    ems = []
    for _ in range(20):
        big.compute_model_energy_all(e_out)
        ems.append(big.last_kernel_ms())
    e_ms = float(np.mean(ems))
    n_evals = nw * rep * up.num_lattices
    evals_per_s = n_evals / (e_ms * 1e-3) * world
    hbm_gbs = n_evals * BYTES_PER_EVAL / (e_ms * 1e-3) / 1e9
    del big

    # ---- end to end through the C ABI with host buffers (pinned), every step
    pin = [torch.from_numpy(a.copy()).pin_memory() for a in (ljr, ref, hm)]
    hl, hr, hh = [p.numpy() for p in pin]
    out_l = torch.empty_like(pin[0]).pin_memory(); out_r = torch.empty_like(pin[1]).pin_memory()
    out_h = torch.empty_like(pin[2]).pin_memory()
    import ctypes as Ct
    states = (W.WalkerState * nw)()

    bufs = [(hl, hr, hh), (out_l.numpy(), out_r.numpy(), out_h.numpy())]

    def e2e_step():
        (il, ir, ih), (ol, or_, oh) = bufs
        g.upload_all(il, ih, ir)                        # H2D: positions, reference positions, cells
        g.energy_init()                                 # lists + energies + mu, as after a checkpoint load
        g.mc_run(C)
        W.check(g.L.mwgpu_download_all(g.h, W._dp(ol), W._dp(or_), W._dp(oh)))
        W.check(g.L.mwgpu_mc_get_states(g.h, states))   # D2H: energies, mu, volumes, counters
        bufs.reverse()                                  # this step's result is the next step's input (no host copy)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    barrier(g)
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_step()
    barrier(g)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = moves_per_step * n_e2e / e2e_s
    h2d = hl.nbytes + hr.nbytes + hh.nbytes
    d2h = h2d + Ct.sizeof(states)
    bad = [s.error for s in states if s.error]
    if bad:
        raise SystemExit(f"bench.py: device error bits {bad[:4]}")

    if rank == 0:
        tb = _traffic("k_mc_run", C, nw)
        line = {
            "metric": "attempted MC moves/sec (whole box)",
            "value": value, "unit": "attempted MC moves/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": _config(nw, total, C, up.nwater),
            "energy_evals_per_s": evals_per_s,
            "energy_evals": {"value": evals_per_s, "unit": "single-lattice full mW energy evals/s (dual-lattice = /2)",
                             "evals_per_launch": n_evals, "ms_per_batch": e_ms, "hbm_gbs_algorithmic": hbm_gbs,
                             "hbm_frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0), "hbm_peak": peak_kind,
                             "fp64_tflops_algorithmic": n_evals * FLOP_PER_EVAL / (e_ms * 1e-3) / 1e12,
                             "kernel": "v2::k_model_energy4<48>",
                             "inputs": f"the {nw} decorrelated walkers x {rep} replicas"},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak else None, "traffic": tb,
                         "traffic_stale": _traffic_stale(),
                         "kernel": "v2::k_mc_run2<2,48>", "kernel_ms": k_ms,
                         "issue": _issue("k_mc_run", C, nw, k_ms, clocks.get("sm_mhz")),
                         "note": "algorithmic flop (10730 per attempted move, BASELINE.md) / measured DFMA peak of this GPU "
                                 "(mwgpu_measure_fp64_peak; MEASURED_PEAKS.json has no fp64 entry); traffic / issue come from "
                                 "the committed ncu capture of the same launch shape (profiles/traffic.json), traffic_stale "
                                 "says whether csrc/ changed since that capture"},
            # the same kernel against the HBM roofline, for the record: the walker state crosses HBM once per launch
            "roofline_hbm": {"bound": "hbm", "achieved": (tb / (k_ms * 1e-3) / 1e9) if tb else None,
                             "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                             "frac": (tb / (k_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0)) if tb else None,
                             "traffic": tb, "peak_kind": peak_kind,
                             "note": "measured DRAM bytes of one launch / live launch duration: the path is not HBM-bound"},
            "e2e": {"value": e2e_value, "unit": "attempted MC moves/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if strong_leg:
            line["strong"] = strong_leg
        if allreduce_check:
            line["allreduce_check"] = allreduce_check
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers", type=int, default=WALKERS_PER_GPU, help="walkers per GPU (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --walkers per GPU (default, what the roofline numbers are quoted on); "
                         "strong: --walkers in total, split evenly over the GPUs")
    ap.add_argument("--cycles", type=int, default=CYCLES_PER_STEP, help="MC cycles per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
