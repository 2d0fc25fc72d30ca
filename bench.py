#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 gravitational force engine.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload direct|tree]

Default workload = BASELINE.json configs[1]: DirectForceComputer, 2^20 particles,
softened gravity (eps = 0.01), one force evaluation per step.  Metric: pairwise
interactions/s (N_targets x N_sources per evaluation, self pair included -- it
contributes exactly 0).  At N > 1 the SAME 2^20-particle problem is target-
sharded over the ranks (strong scaling): each step all-gathers the source
positions with NCCL and every rank evaluates its N/G targets against all sources.

`value`  : device-resident throughput (CUDA events on the launching stream,
           max over ranks, L2 flushed between timed steps).
`e2e`    : the same metric through the host-pointer C-ABI call
           (b200_direct_forces_host; pinned host buffers, H2D + D2H inside the
           timed region).  This is the number to hold against --impl reference.
`roofline`: dominant kernel (direct_kernel) against the FP32 FMA peak measured
           in this run by an FFMA probe (MEASURED_PEAKS.json has no FP32 figure).
`tree_summary`: (default direct run on 1 GPU) a short Barnes-Hut measurement on the same particles --
           ms per build + walk step, interactions/s, particle-steps/s; `--workload tree` gives the full
           line (roofline with ncu traffic, e2e, CPU reference tree), `--ic zeldovich` / `--order morton`
           other inputs, `--kdk` full leapfrog steps.
`cpu_baseline` / --impl reference: the reference's own CPU direct sum
           (oracle/_ref = /root/reference sources compiled in place; falls back
           to the restated port) on all host cores, bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "lambda-cdm-raytracing_b200", "python"))

FLOP_PER_INTERACTION = 20          # GPU-Gems-3 convention
EPS = 0.01


# --------------------------------------------------------------------------- inputs
def make_particles(n, seed=42, order="random"):
    """Uniform in [-50,50)^3, unit masses (what every reference generator emits).  order="morton":
    the same particles stored in Morton (Z-curve) order, the way a production run keeps them, so that a
    contiguous index range -- one rank's shard -- is a compact region of space."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-50.0, 50.0, size=(n, 3)).astype(np.float32)
    if order == "morton":
        q = np.clip(((pos + 50.0) * (1024.0 / 100.0)).astype(np.int64), 0, 1023)

        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v
        key = (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])
        pos = np.ascontiguousarray(pos[np.argsort(key, kind="stable")])
    return pos, np.ones(n, np.float32)


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons while the timed region runs."""

    def __init__(self, index=0, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

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
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- CPU baseline
TREE_SAMPLE = 65536          # particles per single-threaded reference tree evaluation (~1 s per core)
DIRECT_SAMPLE = 16384


def _ref_worker(args):
    workload, kind, n, seed, reps = args
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle, Ref
    pos, mass = make_particles(n, seed)
    if kind != "reference":
        os.environ["OMP_NUM_THREADS"] = "1"
    t0 = time.perf_counter()
    if workload == "direct":
        if kind == "reference":
            r = Ref()
            for _ in range(reps):
                r.direct(pos)                  # TreeForceComputer, one root leaf: the CPU direct sum
        else:
            o = Oracle()
            for _ in range(reps):
                o.direct_f32(pos, None, eps=EPS)
    else:
        if kind == "reference":
            r = Ref()
            for _ in range(reps):
                r.tree_forces(pos, mass, 0.5, 8, 20, 100.0)      # TreeForceComputer::compute_forces: build + walk
        else:
            o = Oracle()
            for _ in range(reps):
                o.tree_forces(o.tree_build(pos, mass), pos, 0.5)
    return time.perf_counter() - t0


def _tree_sample_interactions(n_sample):
    """Cell + pair interactions of ONE reference tree evaluation of an n_sample uniform set (counted
    once, untimed, by the restated oracle whose counters equal the reference walk's by construction)."""
    from oracle.pyoracle import Oracle
    o = Oracle()
    pos, mass = make_particles(n_sample, 1000)
    _, cnt = o.tree_forces(o.tree_build(pos, mass), pos, 0.5, counters=True)
    return float(cnt[1] + cnt[2])


def cpu_reference_batch(cores, n_sample, reps, pool, kind, workload="direct", per_eval=None):
    """One bounded batch: every core runs the reference's single-threaded CPU path on its own
    n_sample-particle subsample.  Returns (interactions, seconds)."""
    t0 = time.perf_counter()
    list(pool.map(_ref_worker, [(workload, kind, n_sample, 1000 + c, reps) for c in range(cores)]))
    dt = time.perf_counter() - t0
    if per_eval is None:
        per_eval = float(n_sample) * float(n_sample)
    return cores * reps * per_eval, dt


def cpu_kind():
    from oracle.pyoracle import Ref
    return "reference" if Ref.available() else "port"


def _cpu_sample(workload):
    if workload == "direct":
        return DIRECT_SAMPLE, None, ("the reference CPU direct sum (TreeForceComputer, leaf_capacity>N => leaf "
                                     "pair loop)")
    return TREE_SAMPLE, _tree_sample_interactions(TREE_SAMPLE), ("the reference CPU TreeForceComputer "
                                                                 "(theta 0.5, leaf 8: build + walk)")


def run_cpu_baseline(workload="direct", target_seconds=12.0):
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    n_sample, per_eval, what = _cpu_sample(workload)
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as pool:
        cpu_reference_batch(cores, 2048, 1, pool, kind, workload)                   # warm: spawn workers, page in
        inter, dt = cpu_reference_batch(cores, n_sample, 1, pool, kind, workload, per_eval)   # calibrate
        reps = max(1, int(round(target_seconds / dt)))
        inter, dt = cpu_reference_batch(cores, n_sample, reps, pool, kind, workload, per_eval)
    return {"value": inter / dt, "unit": "interactions/s", "cores": cores, "kind": kind,
            "sample": f"{cores} independent single-threaded instances of {what}, {n_sample} uniform particles each x "
                      f"{reps} evaluations; {inter:.3e} interactions in {dt:.1f} s"}


def bench_reference(args):
    """--impl reference: the reference's own CPU path, all host cores, bounded steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    n_sample, per_eval, what = _cpu_sample(args.workload)
    reps = 1
    times, inter = [], 0.0
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as pool:
        cpu_reference_batch(cores, 2048, 1, pool, kind, args.workload)
        for _ in range(args.warmup):
            cpu_reference_batch(cores, n_sample, reps, pool, kind, args.workload, per_eval)
        for _ in range(args.steps):
            inter, dt = cpu_reference_batch(cores, n_sample, reps, pool, kind, args.workload, per_eval)
            times.append(dt)
    total = sum(times)
    value = inter * args.steps / total
    sample = (f"each step = {cores} independent single-threaded instances of {what}, "
              f"{n_sample} uniform particles each ({inter:.3e} interactions/step); the rate is per interaction, "
              f"so it transfers to the {args.particles}-particle workload" +
              ("" if args.workload == "direct" else " (a larger tree does more interactions per particle, "
                                                    "at the same cost per interaction)"))
    line = {"impl": "reference", "metric": "pairwise_interactions_per_s", "value": value, "unit": "interactions/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload == "tree":
        line["particle_steps_per_s"] = cores * reps * n_sample * args.steps / total
    print(json.dumps(line))


def zeldovich_grid(n):
    """Power-of-two grid with at least n points: 128 for 2^20 (stride-2 subsample of the
    grid, grid_to_particles), 256 for 2^24 (the first n grid points)."""
    g = 4
    while g ** 3 < n:
        g *= 2
    return g


def workload_config(args):
    cfg = _workload_config(args)
    if getattr(args, "order", "random") == "morton" and getattr(args, "ic", "uniform") == "uniform":
        cfg["workload"] = cfg["workload"].replace("uniform [-50,50)^3", "uniform [-50,50)^3 stored in Morton order")
        cfg["order"] = "morton"
    if getattr(args, "ic", "uniform") == "zeldovich":
        cfg["workload"] = cfg["workload"].replace(
            "uniform [-50,50)^3",
            f"Zel'dovich ICs (b200_zeldovich_ics_dev: {zeldovich_grid(args.particles)}^3 grid, 100 Mpc/h box, "
            f"z = {args.z_initial:g}, seed 12345, sigma_8 0.81, origin-centred)")
        cfg["ic"] = "zeldovich"
    if getattr(args, "kdk", False):
        cfg["workload"] += "; each step = full KDK leapfrog step (omega_m 0.31, omega_lambda 0.69, h 0.67, a0 1, dt 1e-4)"
    return cfg


def _workload_config(args):
    n = args.particles
    if args.workload == "direct":
        return {"workload": f"DirectForceComputer, {n} particles (BASELINE configs[1] when n = 2^20), "
                            f"uniform [-50,50)^3, unit masses, eps = {EPS}, one force evaluation per step; "
                            f"targets sharded over {args.gpus} GPU(s), sources all-gathered",
                "particles": n, "eps": EPS, "l2": "flushed between timed steps (512 MB write)",
                "sources": getattr(args, "sources", "allgather")}
    return {"workload": f"TreeForceComputer Barnes-Hut theta=0.5 leaf=8 max_depth=20, {n} particles "
                        f"(BASELINE configs[2] scale), uniform [-50,50)^3, build + walk per step",
            "particles": n, "theta": 0.5, "leaf_capacity": 8, "l2": "flushed between timed steps (512 MB write)"}


# --------------------------------------------------------------------------- GPU arm
def bench_gpu(args):
    import torch
    import torch.distributed as dist
    import b200grav

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) are
    # pointed at stderr for the whole run; the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = b200grav.Engine(local)

    n = args.particles
    lo, hi = rank * n // world, (rank + 1) * n // world
    nl = hi - lo
    if args.ic == "zeldovich":
        # generated on the device by the engine's own IC step (b200_zeldovich_ics_dev): every rank
        # produces the same replicated particle set (counter-based RNG), origin-centred convention
        posm = torch.empty((n, 4), dtype=torch.float32, device=dev)
        vel_ic = torch.empty((n, 3), dtype=torch.float32, device=dev)
        ic_stats = eng.zeldovich_ics_dev(posm, vel_ic, n_particles=n, grid=zeldovich_grid(n), box=100.0,
                                         z_initial=args.z_initial, seed=12345, origin_shift=50.0)
        del vel_ic
        torch.cuda.synchronize()
        posm_host = posm.cpu().numpy()
        pos, mass = np.ascontiguousarray(posm_host[:, :3]), np.ascontiguousarray(posm_host[:, 3])
    else:
        pos, mass = make_particles(n, order=args.order)
        posm_host = np.ascontiguousarray(np.concatenate([pos, mass[:, None]], 1), np.float32)
        posm = torch.from_numpy(posm_host).to(dev)             # full source set, resident in HBM
    shard = posm[lo:hi].clone() if world > 1 else posm     # this rank's particles (all-gather input)
    acc = torch.zeros((nl, 3), dtype=torch.float32, device=dev)
    vel = torch.zeros((nl, 3), dtype=torch.float32, device=dev)
    kdk = {"a": 1.0, "dt": 1e-4}                           # C4: a0 = 1, dt = 1e-4
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=dev)   # 512 MB > 126 MB L2
    equal_shards = (n % world == 0)

    def gather_sources():
        if world > 1:
            if equal_shards:
                dist.all_gather_into_tensor(posm.view(-1), shard.view(-1))
            else:
                outs = [posm[r * n // world:(r + 1) * n // world] for r in range(world)]
                dist.all_gather(outs, shard)

    masses_equal = bool((mass == mass[0]).all())    # every rank holds the same (replicated) initial masses
    peers = None
    if world > 1 and args.sources == "peer" and args.workload == "direct":
        def _exchange(mine):
            out = [None] * world
            dist.all_gather_object(out, mine)
            return out
        _tok = torch.zeros(1, device=dev)
        peers = b200grav.PeerSources(eng, n, rank, world, lambda: dist.all_reduce(_tok), _exchange)

    def step():
        if peers is not None:       # fused: no all-gather, source tiles are pulled from peer HBM over NVLink
            if args.kdk:
                eng.leapfrog_dev(shard, vel, acc, nl, 2, np.float32(kdk["dt"] * 0.5), kdk["a"], np.float32(kdk["dt"]), 0.0)
                kdk["a"] = eng.scale_factor_step(kdk["a"], kdk["dt"])
            parts = peers.publish(shard)
            eng.direct_forces_parts_dev(parts, peers.lens, shard, nl, acc, eps=EPS, all_masses_equal=masses_equal)
            return
        if args.kdk:    # closing half-kick of the previous step + opening half-kick + drift, one pass
            eng.leapfrog_dev(shard, vel, acc, nl, 2, np.float32(kdk["dt"] * 0.5), kdk["a"], np.float32(kdk["dt"]), 0.0)
            kdk["a"] = eng.scale_factor_step(kdk["a"], kdk["dt"])
        gather_sources()
        if args.workload == "direct":
            eng.direct_forces_dev(posm, acc, lo, nl, eps=EPS)
        else:
            eng.tree_build_dev(posm, n, 100.0, 8, 20)
            eng.tree_walk_dev(acc, lo, nl, theta=0.5)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # FP32 peak of this GPU, measured now (the roofline denominator)
    peak_ffma, _ = eng.fp32_peak_probe(0, 4000)
    peak_ffma2, _ = eng.fp32_peak_probe(1, 4000)
    peak = max(peak_ffma, peak_ffma2)

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    sampler = ClockSampler(local)
    sampler.start()
    eng.set_timing(True)
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    for k in range(args.steps):
        flush.zero_()                                      # evict L2 (untimed)
        sync_all()
        ev[k][0].record()
        step()
        ev[k][1].record()
        torch.cuda.synchronize()
        kernel_ms.append(eng.last_kernel_ms())
    sync_all()
    launches = eng.launches - launches0
    eng.set_timing(False)
    sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    kern_ms = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    kern_s = float(kern_ms.item()) * 1e-3

    # ---- end-to-end through the host-pointer C ABI (pinned host memory) ----
    pin_pos = torch.from_numpy(pos).pin_memory()
    pin_mass = torch.from_numpy(mass).pin_memory()
    pin_out = torch.empty((n if world == 1 else nl, 3), dtype=torch.float32).pin_memory()
    pin_shard = torch.from_numpy(posm_host[lo:hi].copy()).pin_memory()

    def e2e_step():
        if world == 1:
            if args.workload == "direct":
                eng.direct_forces_host(pin_pos.numpy(), pin_mass.numpy(), eps=EPS, out=pin_out.numpy())
            else:
                eng.tree_forces_host(pin_pos.numpy(), pin_mass.numpy(), 0.5, 8, 20, 100.0, out=pin_out.numpy())
        else:   # each rank owns its shard on the host: H2D shard, all-gather, kernels, D2H shard result
            shard.copy_(pin_shard, non_blocking=True)
            step()
            pin_out.copy_(acc, non_blocking=True)
            torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    sync_all()
    t_e2e = []
    for _ in range(e2e_steps):
        flush.zero_()
        sync_all()
        t0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        t_e2e.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([sum(t_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    clocks = sampler.summary()

    tree_bytes = None
    if args.workload == "direct":
        per_step = float(n) * float(n)                      # whole job: all targets x all sources
        per_launch = float(nl) * float(n)                   # this rank's main kernel
    else:
        eng.tree_set_counting(True)
        eng.tree_walk_dev(acc, lo, nl, theta=0.5)
        torch.cuda.synchronize()
        cnt = eng.tree_counters()
        eng.tree_set_counting(False)
        c = torch.tensor([float(cnt[1] + cnt[2])], dtype=torch.float64, device=dev)
        per_launch = float(c.item())
        if world > 1:
            dist.all_reduce(c)
        per_step = float(c.item())
        # algorithmic bytes of one walk launch: 32 B per visited node (centre of mass +
        # node record), 16 B per leaf-pair source, 16 B in + 12 B out per target
        tree_bytes = 32.0 * float(cnt[0]) + 16.0 * float(cnt[2]) + 28.0 * nl
        tree_counts = [int(x) for x in cnt]

    traffic = l2_bytes = l1_bytes = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if world == 1 and args.ic == "uniform":
            traffic = tr.get(args.workload, {}).get(str(n))
            l2_bytes = tr.get(args.workload + "_l2_bytes", {}).get(str(n))
            l1_bytes = tr.get(args.workload + "_l1_bytes", {}).get(str(n))
    except Exception:
        pass

    if rank == 0:
        value = per_step * args.steps / total_s
        e2e_value = per_step * e2e_steps / e2e_s
        achieved = FLOP_PER_INTERACTION * per_launch / kern_s / 1e12
        line = {
            "metric": "pairwise_interactions_per_s", "value": value, "unit": "interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "particle_steps_per_s": float(n) * args.steps / total_s,
            "e2e": {"value": e2e_value, "unit": "interactions/s",
                    "h2d_bytes_per_step": int(16 * n) if world == 1 else int(16 * nl) * world,
                    "d2h_bytes_per_step": int(12 * n), "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "api": (("b200_direct_forces_host" if args.workload == "direct" else "b200_tree_forces_host") +
                            " (pinned host buffers)") if world == 1 else
                           "per-rank pinned shard H2D + NCCL all-gather + b200_*_dev + D2H"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if args.workload == "direct":
            line["roofline"] = {
                "bound": "fp32_fma",
                "kernel": ("direct_kernel<R=8,THREADS=256,1 CTA/SM,open,equal-mass> (11 FP32 lane-ops + 1 MUFU per "
                           "interaction; unequal masses run the 12-op instance)"),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": "FFMA/FFMA2 register-chain probe run in this process (b200_fp32_peak_probe); "
                               "MEASURED_PEAKS.json has no FP32 figure",
                "peak_ffma": peak_ffma, "peak_ffma2": peak_ffma2,
                "nominal_peak": 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12,
                "flop_per_interaction": FLOP_PER_INTERACTION,
                "kernel_ms": 1e3 * kern_s,
            }
        else:
            hbm = 6532.2
            try:
                hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
                src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
            except Exception:
                src = "fallback 6532.2 GB/s (MEASURED_PEAKS.json not on this box)"
            gbs = tree_bytes / kern_s / 1e9
            line["roofline"] = {
                "bound": "hbm", "kernel": "walk_warp_kernel (stackless theta walk over internal nodes; a warp walks the "
                                          "union of the traversals of 32 Hilbert-adjacent targets, per-lane accept test, "
                                          "leaf particles grouped per parent)",
                "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": traffic,
                "peak_source": src,
                "note": "algorithmic bytes = 32 B x nodes visited + 16 B x leaf-pair sources + 28 B x targets; most of it "
                        "is served by L1/L2 (a warp's 32 Hilbert-adjacent targets visit nearly the same nodes), so this is "
                        "an L1/L2 figure quoted against the HBM peak and can exceed it; the ncu capture of the same launch (traffic = DRAM "
                        "bytes, ncu_l2_bytes, ncu_l1_bytes; profiles/r1_walk_warp_kernel_ncu_full.txt) shows a kernel bound by "
                        "instruction issue (81 % of issue slots, l1tex 57 %), not by any memory level",
                "ncu_l2_bytes": l2_bytes, "ncu_l1_bytes": l1_bytes,
                "ncu_l2_gbs": (l2_bytes / kern_s / 1e9) if l2_bytes else None,
                "ncu_l1_gbs": (l1_bytes / kern_s / 1e9) if l1_bytes else None,
                "walk_counters_nodes_cells_pairs": tree_counts, "kernel_ms": 1e3 * kern_s,
                "interactions_per_s_walk_only": per_launch / kern_s,
            }
        if world == 1 and args.workload == "direct" and not args.kdk and not args.no_tree_summary:
            # the second kernel family of the hot path, same particles, a few steps: so that the default
            # run's one line also says where the Barnes-Hut path (BASELINE configs[2]) stands
            line["tree_summary"] = tree_summary(eng, posm, n, flush)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = run_cpu_baseline(args.workload)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        if peers is not None:
            peers.close()
        dist.destroy_process_group()
    eng.close()


def tree_summary(eng, posm, n, flush, steps=5):
    """Barnes-Hut theta = 0.5, leaf 8, max_depth 20 on the device-resident particles: build + walk per step,
    CUDA events around each step, L2 flushed between steps.  The full line is `--workload tree`."""
    import torch
    acc = torch.empty((n, 3), dtype=torch.float32, device=posm.device)
    for _ in range(3):
        eng.tree_build_dev(posm, n, 100.0, 8, 20)
        eng.tree_walk_dev(acc, 0, n, theta=0.5)
    torch.cuda.synchronize()
    eng.set_timing(True)
    ms, walk_ms = [], []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.tree_build_dev(posm, n, 100.0, 8, 20)
        eng.tree_walk_dev(acc, 0, n, theta=0.5)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        walk_ms.append(eng.last_kernel_ms())
    eng.set_timing(False)
    eng.tree_set_counting(True)
    eng.tree_walk_dev(acc, 0, n, theta=0.5)
    torch.cuda.synchronize()
    cnt = eng.tree_counters()
    eng.tree_set_counting(False)
    step_s = float(np.mean(ms)) * 1e-3
    return {"workload": f"TreeForceComputer theta=0.5 leaf=8 max_depth=20, {n} particles, build + walk per step",
            "ms_per_step": 1e3 * step_s, "walk_kernel_ms": float(np.mean(walk_ms)),
            "interactions_per_s": float(cnt[1] + cnt[2]) / step_s, "particle_steps_per_s": n / step_s,
            "walk_counters_nodes_cells_pairs": [int(x) for x in cnt], "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="direct", choices=["direct", "tree"])
    ap.add_argument("--particles", type=int, default=1 << 20)
    ap.add_argument("--ic", default="uniform", choices=["uniform", "zeldovich"],
                    help="synthetic inputs: uniform random (default) or Zel'dovich initial conditions generated on "
                         "the device")
    ap.add_argument("--z-initial", type=float, default=49.0, help="redshift of the Zel'dovich ICs")
    ap.add_argument("--order", default="random", choices=["random", "morton"],
                    help="index order of the uniform particles: as drawn (default) or sorted along a Morton curve")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tree-summary", action="store_true",
                    help="default (direct, 1 GPU) run: skip the short Barnes-Hut measurement added as `tree_summary`")
    ap.add_argument("--sources", default="allgather", choices=["allgather", "peer"],
                    help="N>1 direct sum: NCCL all-gather of the shards (default) or peer-mapped source tiles "
                         "pulled over NVLink by the force kernel itself")
    ap.add_argument("--kdk", action="store_true",
                    help="each step is a full Lambda-CDM KDK leapfrog step (fused kick-kick-drift pass, scale-factor "
                         "update, source all-gather, force evaluation) instead of a bare force evaluation")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_gpu(args)


if __name__ == "__main__":
    main()
