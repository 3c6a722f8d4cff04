#!/usr/bin/env python
"""bench.py -- MCL update throughput on B200 (and the reference's CPU path beside it).

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A "step" is one full MCL update (resample -> motion -> ray cast -> weights -> normalise ->
expected pose) of the workload.

--workload track (default; BASELINE.json configs[2], the configuration the north_star target is
quoted on): Spielberg_map (2000x2000, MAX_RANGE_PX 207), 1,048,576 particles x 60 beams per GPU,
tracking cloud on the track, synthetic scan + odometry replay.  At N > 1 the ONE global filter is
particle-sharded (weak scaling: 1M particles per rank, the reference's exact global multinomial
resampling; every exchange is done by the kernels over NVLink, include/mcl_b200.h).
--workload batch (BASELINE.json configs[3]): 1024 independent 4000-particle filters on sibal1,
filter-sharded over the GPUs (1024 / N filters per rank, no collective).

  value  : ray-casts/s with inputs resident in HBM (mcl_update_dev), timed per step with CUDA
           events on the launching stream, L2 flushed between steps, max over ranks.
  e2e    : the same metric through the host-facing C-ABI call mcl_update with HOST buffers
           (action + scan copied in, pose copied out, inside the timed region).
  roofline: the dominant kernel (the ray march): algorithmic bytes / its CUDA-event
           duration vs the measured HBM peak in MEASURED_PEAKS.json, plus what really bounds it
           (warp issue, from the committed ncu capture when it matches this configuration);
           `kernels` lists EVERY kernel of the update with its own CUDA-event time, algorithmic
           bytes and fraction of the HBM peak.
  cpu_baseline: the reference's own CPU update (oracle/_ref, else the oracle port) on a
           bounded sample, all host threads -- a reported baseline, not the target.

`--impl reference` times only that CPU arm, same metric/unit/config (default: the full
1,048,576-particle filter; the reference then holds a 1.44 GB query matrix).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 20
METRIC = "MCL ray-casts/s (particles x beams x updates/s)"
UNIT = "rays/s"
WORKLOADS = {
    "track": {"map": "Spielberg_map", "speed": 8.0, "max_range_px": 207},
    "batch": {"map": "sibal1", "speed": 3.0, "max_range_px": 239, "filters": 1024, "particles": 4000},
}


def workload_config(workload: str, n_gpus: int, n_particles: int, R: int, exchange: str = "fused", route: str = "two-hop") -> dict:
    wl = WORKLOADS[workload]
    if workload == "batch":
        F = wl["filters"]
        return {"workload": "BASELINE configs[3]: %s, %d independent filters x %d particles x %d beams, multi-car replay" % (
            wl["map"], F, wl["particles"], R),
                "map": wl["map"], "filters": F, "filters_per_gpu": F // n_gpus, "particles_per_filter": wl["particles"],
                "beams": R, "max_range_px": wl["max_range_px"],
                "sharding": "single GPU" if n_gpus == 1 else "filter-sharded x%d, no collective" % n_gpus,
                "l2": "flushed between timed steps (256 MiB write)", "rng": "device Philox (no injected noise)"}
    return {"workload": "BASELINE configs[2]: %s, %d particles x %d beams per GPU, tracking replay" % (
        wl["map"], n_particles, R),
            "map": wl["map"], "particles_per_gpu": n_particles, "particles_global": n_particles * n_gpus,
            "beams": R, "max_range_px": wl["max_range_px"],
            "sharding": "single GPU" if n_gpus == 1 else
            "particle-sharded x%d, exact global multinomial resampling, slice-local state, %s routing, %s exchange" % (
                n_gpus, route, exchange),
            "l2": "flushed between timed steps (256 MiB write)", "rng": "device Philox (no injected noise)"}


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
def make_replay(ctx, grid, n_steps: int, seed: int, speed: float):
    """Synthetic trajectory + scans; scans are ray cast by the product's own calc_range_many."""
    from monte_carlo_localization_b200 import synth
    angles_full = synth.laser_angles()
    gt, actions = synth.trajectory(grid, n_steps, speed)
    rng = np.random.default_rng(seed)
    obs = np.stack([synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(n_steps)])
    return gt, actions, np.ascontiguousarray(obs, dtype=np.float32)


def cpu_reference_arm(workload: str, n_particles: int, steps: int, warmup: int, threads: int | None = None) -> dict:
    """Times the reference's own CPU MCL()+expected_pose() (oracle/_ref) or, where that was
    never built, the oracle port, on the bench workload's map / scans at `n_particles`."""
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    wl = WORKLOADS[workload]
    grid = maps.load_named_map(wl["map"])
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    R = len(angles)
    cores = threads or (os.cpu_count() or 1)
    n_tot = steps + warmup
    gt, actions = synth.trajectory(grid, n_tot, wl["speed"])
    helper = ob.Oracle(grid, angles, max_particles=1, num_threads=cores)
    rng = np.random.default_rng(777 + 3)
    scans = [synth.scan_from_pose(helper.calc_range_many, gt[t + 1], angles_full, rng) for t in range(n_tot)]
    kind = "reference" if ob.have_reference() else "port"
    times = []
    if kind == "reference":
        ref = ob.Reference(grid, 20250 + 3, max_particles=n_particles, num_threads=cores,
                           use_parallel_raycasting=True)
        ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scans[0])
        ref.init_pose(gt[0])
        for t in range(n_tot):
            ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scans[t])
            t0 = time.perf_counter()
            ref.mcl(actions[t], scans[t][::18])
            times.append(time.perf_counter() - t0)
        buckets = ref.timing()
    else:
        orc = ob.Oracle(grid, angles, max_particles=n_particles, num_threads=cores)
        ns = ob.NoiseStream(20250 + 3)
        orc.init_pose(gt[0], ns.normal(3 * n_particles))
        for t in range(n_tot):
            u, z = ns.update_noise(n_particles)
            t0 = time.perf_counter()
            orc.update(actions[t], scans[t][::18], u, z)
            orc.expected_pose()
            times.append(time.perf_counter() - t0)
        buckets = orc.timing()
    timed = times[warmup:]
    sec = float(np.sum(timed))
    rays = n_particles * R * len(timed)
    cnt = max(1, buckets.get("count", 1))
    return {"value": rays / sec, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%s, %d particles x %d beams, %d updates after %d warm-up, median %.1f ms/update" % (
                wl["map"], n_particles, R, len(timed), warmup, 1e3 * float(np.median(timed))),
            "ms_per_update": 1e3 * sec / len(timed), "updates_per_s": len(timed) / sec,
            "buckets_ms_per_update": {k: v / cnt for k, v in buckets.items() if k.endswith("_ms")},
            "particles": n_particles, "beams": R}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    if args.workload == "batch":
        # one step = a bounded sample of the batch: single 4000-particle filter updates, as the reference runs them
        n_ref = wl["particles"]
        note = "each step = one reference MCL()+expected_pose() of ONE %d-particle filter of the batch" % n_ref
    else:
        n_ref = args.ref_particles or args.particles
        note = "each step = one reference MCL()+expected_pose() on %d particles%s" % (
            n_ref, " (the reference materialises a %.2f GB query matrix)" % (n_ref * 60 * 24 / 1e9) if n_ref >= 500000 else "")
    cb = cpu_reference_arm(args.workload, n_ref, args.steps, max(min(args.warmup, 3), 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_update"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "updates_per_s": cb["updates_per_s"],
            "config": dict(workload_config(args.workload, args.gpus, args.particles, cb["beams"], args.shard_exchange,
                                           args.shard_route if args.shard_route != "auto" else ("two-hop" if args.gpus >= 3 else "one-hop")),
                           reference_sample=note),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "buckets_ms_per_update")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def kernel_table(kernel_samples, N: int, R: int, cbar: float | None, peak_gbs: float) -> list:
    """Per-kernel CUDA-event time (mean over the profiled steps), algorithmic bytes per launch
    (SURVEY 8d accounting, FP64 state) and the fraction of the measured HBM peak they amount to."""
    if not kernel_samples:
        return []
    log2n = math.ceil(math.log2(max(N, 2)))
    alg = {
        "k_prepare_obs": lambda: R * 208 * 16,
        "k_tile_sums": lambda: 8 * N,
        "k_exact_pass(S1)": lambda: 10 * N,
        "k_exact_pass(normalise+pose+S2)": lambda: 42 * N,
        "k_exact_pass(S2)": lambda: 10 * N,
        "k_exact_pass(cdf)": lambda: 10 * N,
        "k_exact_emit": lambda: 18 * N,
        "k_resample_motion": lambda: (104 + 8 * log2n) * N,
        "k_resample_motion(routed)": lambda: (32 + 24 + 32 + 32 + 4) * N,
        "k_route": lambda: (32 + 32 + 8 * log2n) * N,
        "k_route_request": lambda: 4 * N,
        "k_route_serve": lambda: (4 + 32 + 32 + 8 * log2n) * N,
        "k_sort_scatter": lambda: 20 * N,
        "k_dir_gather": lambda: 68 * N,
        "k_dir_plan": lambda: 4 * 4096,
        "k_raycast_dir": lambda: (N * R * cbar + 32 * N + N * R) if cbar else None,
        "k_raycast_weight": lambda: (N * R * cbar + 32 * N) if cbar else None,
        "k_weight_steps": lambda: N * (R + 12),
    }
    # the steady-state launch sequence (the first update after a state restore also rebuilds tile sums and S2)
    seqs = {}
    for s in kernel_samples:
        seqs.setdefault(tuple(n for n, _ in s), []).append(s)
    names, samples = max(seqs.items(), key=lambda kv: len(kv[1]))
    out = []
    for i, name in enumerate(names):
        ms = float(np.mean([s[i][1] for s in samples]))
        b = alg.get(name, lambda: None)()
        row = {"name": name, "ms": ms}
        if b and ms > 0:
            gbs = b / (ms * 1e-3) / 1e9
            row.update({"algorithmic_bytes": int(b), "gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak_gbs})
            if gbs / peak_gbs > 1.0:
                row["note"] = "L2 / shared-memory resident: the bytes counted are the reference's reads, not DRAM traffic"
        out.append(row)
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from monte_carlo_localization_b200 import MclContext, maps, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG=VERSION is set in the environment;
        # stdout must carry exactly ONE JSON line, so file descriptor 1 points at stderr while the
        # communicators are created
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
    try:
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        wl = WORKLOADS[args.workload]
        batch = args.workload == "batch"
        grid = maps.load_named_map(wl["map"])
        angles = synth.beam_angles()
        R = len(angles)
        N = wl["particles"] if batch else args.particles
        F = wl["filters"] // world if batch else 1
        K, W = args.steps, args.warmup
        n_tot = K + W + 4

        flt = None
        if batch:
            ctx = MclContext(device=local_rank, max_particles=N, num_filters=F, seed=20254 + rank)
            ctx.set_map(grid)
            ctx.set_beam_angles(angles)
        elif world > 1:
            from monte_carlo_localization_b200.sharded import ShardedFilter
            flt = ShardedFilter(grid, angles, n_local=N, rank=rank, world=world, device=local_rank, seed=20250 + 3,
                                exchange=args.shard_exchange, route=args.shard_route)
            ctx = flt.ctx
        else:
            ctx = MclContext(device=local_rank, max_particles=N, seed=20250 + 3)
            ctx.set_map(grid)
            ctx.set_beam_angles(angles)
    finally:
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if args.ray_mode:
        ctx.set_ray_mode(args.ray_mode)
    if args.no_pdl:
        ctx.set_pdl(False)
    # the library launches on this (non-default) torch stream so torch CUDA events time its kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    first = rank * F   # batch: every car starts at its own phase of the lap
    gt, actions, obs = make_replay(ctx, grid, n_tot + (wl["filters"] if batch else 0), seed=777 + 3, speed=wl["speed"])
    if batch:
        for f in range(F):
            ctx.init_pose(gt[first + f], filter=f)
    elif flt is not None:
        flt.init_pose(gt[0])
    else:
        ctx.init_pose(gt[0])
    d_actions = torch.from_numpy(np.ascontiguousarray(actions)).cuda()
    d_obs = torch.from_numpy(obs).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fidx = torch.arange(first, first + F, device="cuda") if batch else None
    torch.cuda.synchronize()

    def inputs_dev(t):
        if batch:
            return d_actions[fidx + t].contiguous(), d_obs[fidx + t].contiguous()
        return d_actions[t], d_obs[t]

    def inputs_host(t):
        if batch:
            return np.ascontiguousarray(actions[first + t:first + t + F]), np.ascontiguousarray(obs[first + t:first + t + F])
        return np.ascontiguousarray(actions[t]), np.ascontiguousarray(obs[t])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks are sampled from the start of the warm-up to the end of the timed region (the timed
    # region alone can be shorter than nvidia-smi's first report)
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up ---------------------------------------------------------------------------
    t = 0
    for _ in range(W):
        a_, o_ = inputs_dev(t)
        ctx.update_dev(a_.data_ptr(), o_.data_ptr())
        t += 1
    barrier()
    # snapshot of the filter at the start of the timed region, so that the legs below replay
    # exactly the same K steps from exactly the same state
    t_start = t
    snap = [(ctx.get_particles(f), ctx.get_weights(f)) for f in range(F)]

    def restore():
        for f in range(F):
            ctx.set_particles(snap[f][0], snap[f][1], filter=f)

    # ---- timed: inputs resident in HBM -----------------------------------------------------
    launches0 = ctx.kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    keep = []
    barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        a_, o_ = inputs_dev(t)
        keep.append((a_, o_))
        if not os.environ.get('BENCH_NOFLUSH'):
            flush.zero_()                  # L2 flush, outside the timed interval
        ev[k][0].record(stream)
        ctx.update_dev(a_.data_ptr(), o_.data_ptr())
        ev[k][1].record(stream)
        t += 1
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches() - launches0
    ray_stage = ctx.ray_stage_info()      # which ray stage the last timed update ran
    dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    clocks = sampler.stop()

    # ---- e2e: host buffers through the C-ABI call, the same K steps from the same state ------
    host_in = [inputs_host(t_start + i) for i in range(K)]
    restore()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        # H2D action+scan, pose back to the host, host sync -- every step
        pose = ctx.update(host_in[i][0], host_in[i][1])
    barrier()
    e_sec = time.perf_counter() - t0
    pose = np.asarray(pose).reshape(-1, 3)
    want = gt[first + np.arange(F) + t_start + K] if batch else gt[t_start + K][None]
    pose_err = float(np.median(np.hypot(pose[:, 0] - want[:, 0], pose[:, 1] - want[:, 1])))
    if world > 1:
        te = torch.tensor([e_sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_sec = float(te.item())
    rays_per_step = N * F * world * R
    e2e = {"value": rays_per_step * K / e_sec, "unit": UNIT, "h2d_bytes_per_step": (24 + 4 * R) * F,
           "d2h_bytes_per_step": 24 * F, "steps": K, "ms_per_step": 1e3 * e_sec / K}

    # ---- profiled replay of the same K steps from the same state: CUDA events around every kernel
    # (kept out of the legs above so that the event records do not sit in their timed regions) ---
    restore()
    ctx.set_profiling(True)
    stages, kernel_samples = [], []
    barrier()
    for i in range(K):
        ctx.update(host_in[i][0], host_in[i][1])
        stages.append(ctx.stage_ms())
        kernel_samples.append(ctx.kernel_ms())
    ctx.set_profiling(False)
    stage = {k: float(np.mean([s_[k] for s_ in stages])) for k in stages[0]} if stages else None
    # ---- C-bar, the cells the reference march samples per ray, on a few of the same steps
    # (storing per-ray steps slows the kernels: diagnostics only, never timed; the sharded filter
    # samples its own slice) --------------------------------------------------------------------
    cbar = None
    if not batch:
        restore()
        every = max(1, K // 4)
        cb_samples = []
        barrier()
        for i in range(K):
            keep_it = (i % every) == every - 1
            ctx.set_keep_ranges(keep_it)
            ctx.update(host_in[i][0], host_in[i][1])
            if keep_it:
                st_ = ctx.range_steps()
                cb_samples.append(float(np.where(st_ >= ctx.M, ctx.M, st_.astype(np.int64) + 1).mean()))
        ctx.set_keep_ranges(False)
        cbar = float(np.mean(cb_samples)) if cb_samples else None
    barrier()

    # ---- max over ranks --------------------------------------------------------------------
    if world > 1:
        tm = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dev_ms = float(tm.item())
    ms_per_step = dev_ms / K
    value = rays_per_step / (ms_per_step * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        kernels = kernel_table(kernel_samples, N * F, R, cbar, peak)
        roof = None
        ktime = {k_["name"]: k_["ms"] for k_ in kernels}
        directional = ray_stage.get("last_mode") == 1
        kname = "k_raycast_dir" if directional else "k_raycast_weight"
        if kname in ktime and ktime[kname] > 0:
            # dominant kernel: the ray march; its own CUDA-event time is the denominator
            k_ms = ktime[kname]
            c_eff = cbar if cbar is not None else 42.0   # batch workload: SURVEY's sibal1 figure (not measured here)
            # cells the reference samples (1 B each) + ray-start record read + step/weight write
            alg_bytes = N * F * R * c_eff + (N * F * R * 1 + N * F * 32 if directional else N * F * (24 + 8))
            ach = alg_bytes / (k_ms * 1e-3) / 1e9
            traffic, issue = None, {"frac": None, "reason": "no ncu capture committed for this configuration"}
            tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tp):
                try:
                    prof = json.load(open(tp))
                    # only a capture of THIS kernel at THIS configuration says anything about this run
                    match = (prof.get("map") == wl["map"] and prof.get("particles") == N * F and prof.get("beams") == R
                             and kname + "_warp_instructions_per_launch" in prof)
                    if match:
                        traffic = prof.get(kname + "_dram_bytes_per_launch")
                        winst = prof[kname + "_warp_instructions_per_launch"]
                        if clocks.get("sm_mhz"):
                            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                            peak_issue = 4.0 * sms * clocks["sm_mhz"] * 1e6
                            issue = {"warp_instructions_per_launch": winst, "source": "profiles/ncu_traffic.json (%s)" % prof.get("capture", "ncu capture"),
                                     "achieved_per_s": winst / (k_ms * 1e-3), "peak_per_s": peak_issue,
                                     "frac": winst / (k_ms * 1e-3) / peak_issue,
                                     "lane_efficiency": prof.get(kname + "_threads_per_instruction"),
                                     "shared_bank_conflict_share": prof.get(kname + "_shared_conflict_share")}
                    else:
                        issue = {"frac": None, "reason": "profiles/ncu_traffic.json was captured at another configuration"}
                except Exception:
                    pass
            # gather-rate context for the same kernel: random byte reads/s the chip sustains from a
            # shared-memory window and from a 4 MB L2-resident array
            from monte_carlo_localization_b200 import capi as _capi
            gather = {"shared_memory_peak_per_s": _capi.microbench_gather(True, device=local_rank),
                      "l2_4mb_peak_per_s": _capi.microbench_gather(False, device=local_rank),
                      "reference_samples_per_s": N * F * R * c_eff / (k_ms * 1e-3)}
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "mean_cells_per_ray": cbar, "kernel_ms": k_ms,
                    "share_of_step": k_ms / max(sum(ktime.values()), 1e-9),
                    "resident": "L2 + shared memory: the skip maps never leave the chip and most of the reference's "
                                "samples are proven unnecessary, so the algorithmic-byte rate can exceed the HBM peak; "
                                "the kernel is bound by warp issue (see `issue`), not by HBM",
                    "issue": issue, "gather": gather}
        cb = None
        if world == 1 and not args.no_cpu:
            n_cpu = wl["particles"] if batch else (args.ref_particles or 100000)
            cb = cpu_reference_arm(args.workload, n_cpu, 3, 1)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "buckets_ms_per_update")}
            # the reference's shipped thread count (config/mcl_config.yaml:40 num_threads: 3), for context
            shipped = cpu_reference_arm(args.workload, max(1000, n_cpu // 4), 2, 1, threads=3)
            cb["shipped_num_threads_3"] = {"value": shipped["value"], "unit": UNIT, "cores": 3, "sample": shipped["sample"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "updates_per_s": 1e3 / ms_per_step * (F * world if batch else 1),
                "config": workload_config(args.workload, world, N, R, args.shard_exchange, flt.route if flt is not None else "none"), "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cb, "stage_ms": stage, "kernels": kernels,
                "ray_stage": ray_stage, "wall_ms_per_step_incl_flush": 1e3 * wall / K, "pose_error_m": pose_err}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="track", choices=sorted(WORKLOADS),
                    help="track: BASELINE configs[2] (default, the headline); batch: configs[3], 1024 x 4000-particle filters")
    ap.add_argument("--particles", type=int, default=N_PER_GPU, help="particles per GPU (track workload)")
    ap.add_argument("--ref-particles", type=int, default=None,
                    help="particles of the CPU runs: default the full filter for --impl reference, 100000 for the "
                         "cpu_baseline leg of the GPU arm (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ray-mode", type=int, default=0, choices=[0, 1, 2],
                    help="0 auto (default), 1 isotropic skip-map kernel only, 2 directional stage always")
    ap.add_argument("--shard-exchange", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU: how ranks meet at an exchange (in-kernel NVLink flags, or a one-word ncclAllGather)")
    ap.add_argument("--no-pdl", action="store_true", help="plain stream order between the kernels of an update (comparison)")
    ap.add_argument("--shard-route", default="auto", choices=["auto", "two-hop", "one-hop"],
                    help="multi-GPU: request routing of the resampling draws, every rank testing all draws, or the library's "
                         "choice by world size (default: two hops from 3 ranks on)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)   # each step ~4 s of CPU work at 1M particles on 16 cores
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
