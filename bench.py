#!/usr/bin/env python
"""bench.py -- MCL update throughput on B200 (and the reference's CPU path beside it).

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A "step" is one full MCL update (resample -> motion -> ray cast -> weights -> normalise ->
expected pose) of the workload below.

Workload (BASELINE.json configs[2], the configuration the north_star target is quoted on):
Spielberg_map (2000x2000, MAX_RANGE_PX 207), 1,048,576 particles x 60 beams per GPU, tracking
cloud on the track, synthetic scan + odometry replay.  At N > 1 the ONE global filter is
particle-sharded (weak scaling: 1M particles per rank, exact global multinomial resampling,
one all-gather of the rank blocks per update).

  value  : ray-casts/s with inputs resident in HBM (mcl_update_dev), timed per step with CUDA
           events on the launching stream, L2 flushed between steps, max over ranks.
  e2e    : the same metric through the host-facing C-ABI call mcl_update with HOST buffers
           (action + scan copied in, pose copied out, inside the timed region).
  roofline: k_raycast_weight, algorithmic bytes / its CUDA-event duration vs the
           measured HBM peak in MEASURED_PEAKS.json.
  cpu_baseline: the reference's own CPU update (oracle/_ref, else the oracle port) on a
           bounded sample, all host threads -- a reported baseline, not the target.

`--impl reference` times only that CPU arm, same metric/unit/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAP_NAME = "Spielberg_map"
N_PER_GPU = 1 << 20
SPEED = 8.0
METRIC = "MCL ray-casts/s (particles x beams x updates/s)"
UNIT = "rays/s"


def workload_config(n_gpus: int, n_particles: int, R: int, shard_mode: str = "p2p") -> dict:
    return {"workload": "BASELINE configs[2]: %s, %d particles x %d beams per GPU, tracking replay" % (
        MAP_NAME, n_particles, R),
            "map": MAP_NAME, "particles_per_gpu": n_particles, "particles_global": n_particles * n_gpus,
            "beams": R, "max_range_px": 207,
            "sharding": "single GPU" if n_gpus == 1 else "particle-sharded x%d, exact global resampling, %s exchange" % (
                n_gpus, shard_mode),
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
def make_replay(ctx, grid, n_steps: int, seed: int):
    """Synthetic trajectory + scans; scans are ray cast by the product's own calc_range_many."""
    from monte_carlo_localization_b200 import synth
    angles_full = synth.laser_angles()
    gt, actions = synth.trajectory(grid, n_steps, SPEED)
    rng = np.random.default_rng(seed)
    obs = np.stack([synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(n_steps)])
    return gt, actions, np.ascontiguousarray(obs, dtype=np.float32)


def cpu_reference_arm(n_particles: int, steps: int, warmup: int, threads: int | None = None) -> dict:
    """Times the reference's own CPU MCL()+expected_pose() (oracle/_ref) or, where that was
    never built, the oracle port, on the bench workload at a bounded particle count."""
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    grid = maps.load_named_map(MAP_NAME)
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    R = len(angles)
    cores = threads or (os.cpu_count() or 1)
    n_tot = steps + warmup
    gt, actions = synth.trajectory(grid, n_tot, SPEED)
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
                MAP_NAME, n_particles, R, len(timed), warmup, 1e3 * float(np.median(timed))),
            "ms_per_update": 1e3 * sec / len(timed), "updates_per_s": len(timed) / sec,
            "buckets_ms_per_update": {k: v / cnt for k, v in buckets.items() if k.endswith("_ms")},
            "particles": n_particles, "beams": R}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_ref = args.ref_particles
    cb = cpu_reference_arm(n_ref, args.steps, max(min(args.warmup, 3), 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_update"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "updates_per_s": cb["updates_per_s"],
            "config": dict(workload_config(args.gpus, N_PER_GPU, cb["beams"]),
                           reference_sample="each step = one reference MCL()+expected_pose() on %d particles" % n_ref),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "buckets_ms_per_update")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
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
        # communicator is created
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    grid = maps.load_named_map(MAP_NAME)
    angles = synth.beam_angles()
    R = len(angles)
    N = args.particles
    K, W = args.steps, args.warmup
    n_tot = K + W + 4

    if world > 1:
        from monte_carlo_localization_b200.sharded import ShardedFilter
        flt = ShardedFilter(grid, angles, n_local=N, rank=rank, world=world, device=local_rank, seed=20250 + 3,
                            mode=args.shard_mode)
        ctx = flt.ctx
        args.shard_mode = flt.mode      # "allgather" if the peer mapping was refused on this host
    else:
        flt = None
        ctx = MclContext(device=local_rank, max_particles=N, seed=20250 + 3)
        ctx.set_map(grid)
        ctx.set_beam_angles(angles)
    if args.ray_mode:
        ctx.set_ray_mode(args.ray_mode)
    # the library launches on this (non-default) torch stream so torch CUDA events time its kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    gt, actions, obs = make_replay(ctx, grid, n_tot, seed=777 + 3)
    if flt is not None:
        flt.init_pose(gt[0])
    else:
        ctx.init_pose(gt[0])
    d_actions = torch.from_numpy(np.ascontiguousarray(actions)).cuda()
    d_obs = torch.from_numpy(obs).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def step_dev(t):
        if flt is not None:
            flt.update_dev(d_actions[t].data_ptr(), d_obs[t].data_ptr())
        else:
            ctx.update_dev(d_actions[t].data_ptr(), d_obs[t].data_ptr())

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
        step_dev(t)
        t += 1
    barrier()
    # snapshot of the filter at the start of the timed region, so that the e2e leg below
    # replays exactly the same K steps from exactly the same state
    t_start = t
    snap_p, snap_w = flt.gather_state() if flt is not None else (ctx.get_particles(), ctx.get_weights())

    # ---- timed: inputs resident in HBM -----------------------------------------------------
    launches0 = ctx.kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    wall0 = time.perf_counter()
    for k in range(K):
        if not os.environ.get('BENCH_NOFLUSH'):
            flush.zero_()                  # L2 flush, outside the timed interval
        ev[k][0].record(stream)
        step_dev(t)
        ev[k][1].record(stream)
        t += 1
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches() - launches0
    ray_stage = ctx.ray_stage_info()      # which ray stage the last timed update ran
    dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    clocks = sampler.stop()

    # ---- e2e: host buffers through the C-ABI call, the same K steps from the same state ------
    # (stage events are recorded too: the host-facing call synchronises every step, so the
    # per-kernel CUDA-event times of exactly these K steps can be read back)
    Ke = K
    acts_h = [np.ascontiguousarray(actions[t_start + i]) for i in range(Ke)]
    obs_h = [np.ascontiguousarray(obs[t_start + i]) for i in range(Ke)]
    ctx.set_particles(snap_p, snap_w)
    stages = []
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        # H2D action+scan, pose back to the host, host sync -- every step
        pose = ctx.update(acts_h[i], obs_h[i]) if flt is None else flt.update(acts_h[i], obs_h[i])
    barrier()
    e_sec = time.perf_counter() - t0
    pose_err = float(np.hypot(*(np.asarray(pose)[:2] - gt[t_start + Ke][:2])))
    if world > 1:
        te = torch.tensor([e_sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_sec = float(te.item())
    e2e = {"value": N * world * R * Ke / e_sec, "unit": UNIT, "h2d_bytes_per_step": 24 + 4 * R,
           "d2h_bytes_per_step": 24, "steps": Ke, "ms_per_step": 1e3 * e_sec / Ke}

    # ---- third replay of the same K steps from the same state: per-stage CUDA-event times
    # (profiling on; kept out of the e2e leg so that the event records do not sit in its timed
    # region) and C-bar, the cells the reference march samples per ray, on a sample of the steps
    cbar = None
    if flt is None:
        ctx.set_particles(snap_p, snap_w)
        ctx.set_profiling(True)
        every = max(1, K // 16)
        cb_samples = []
        for i in range(K):
            keep = (i % every) == every - 1
            ctx.set_keep_ranges(keep)      # storing per-ray steps slows the kernels: diagnostics only
            ctx.update(acts_h[i], obs_h[i])
            if keep:
                st_ = ctx.range_steps()
                cb_samples.append(float(np.where(st_ >= ctx.M, ctx.M, st_.astype(np.int64) + 1).mean()))
            else:
                stages.append(ctx.stage_ms())
        ctx.set_keep_ranges(False)
        ctx.set_profiling(False)
        cbar = float(np.mean(cb_samples))
    elif args.shard_stages:
        # sharded filter: per-stage times of this rank from a profiled replay of the same steps
        # ("exchange" = the NCCL all-gather between the local stages and the finish stage)
        flt.set_state(snap_p, snap_w)
        ctx.set_profiling(True)
        for i in range(K):
            flt.update(acts_h[i], obs_h[i])
            stages.append(ctx.stage_ms())
        ctx.set_profiling(False)
    stage = {k: float(np.mean([s_[k] for s_ in stages])) for k in stages[0]} if stages else None

    # ---- max over ranks --------------------------------------------------------------------
    if world > 1:
        tm = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dev_ms = float(tm.item())
    ms_per_step = dev_ms / K
    n_global = N * world
    value = n_global * R / (ms_per_step * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        roof = None
        if stage is not None and cbar is not None and stage["raycast_weight"] > 0:
            # dominant kernel: the ray march.  k_raycast_dir (directional stage) when the plan chose it,
            # else k_raycast_weight; its own CUDA-event time (stage "ray_march") is the denominator
            directional = ray_stage.get("last_mode") == 1
            kname = "k_raycast_dir" if directional else "k_raycast_weight"
            k_ms = stage.get("ray_march") or stage["raycast_weight"]
            # cells the reference samples (1 B each) + ray-start record read + step/weight write
            alg_bytes = N * R * cbar + (N * R * 1 + N * 32 * 2 if directional else N * (24 + 8))
            ach = alg_bytes / (k_ms * 1e-3) / 1e9
            traffic, issue = None, None
            tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tp):
                try:
                    prof = json.load(open(tp))
                    traffic = prof.get(kname + "_dram_bytes_per_launch")
                    # what actually bounds the kernel: warp-instruction issue.  Instructions per launch from the
                    # committed ncu capture of this workload, time measured live, peak = 4 schedulers x SMs x clock
                    winst = prof.get(kname + "_warp_instructions_per_launch")
                    if winst and clocks.get("sm_mhz"):
                        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                        peak_issue = 4.0 * sms * clocks["sm_mhz"] * 1e6
                        issue = {"warp_instructions_per_launch": winst, "source": "profiles/ncu_traffic.json (ncu capture)",
                                 "achieved_per_s": winst / (k_ms * 1e-3), "peak_per_s": peak_issue,
                                 "frac": winst / (k_ms * 1e-3) / peak_issue}
                except Exception:
                    traffic, issue = None, None
            # gather-rate context for the same kernel: random byte reads/s the chip sustains from a
            # shared-memory window and from a 4 MB L2-resident array
            from monte_carlo_localization_b200 import capi as _capi
            gather = {"shared_memory_peak_per_s": _capi.microbench_gather(True, device=local_rank),
                      "l2_4mb_peak_per_s": _capi.microbench_gather(False, device=local_rank),
                      "reference_samples_per_s": N * R * cbar / (k_ms * 1e-3)}
            roof = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "gather": gather, "issue": issue,
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "mean_cells_per_ray": cbar,
                    "kernel_ms": k_ms,
                    "note": "skip maps are L2+shared-memory resident; algorithmic bytes are the reference's per-sample "
                            "grid reads, most of which the kernel proves unnecessary, so frac can exceed 1"}
        cb = None
        if world == 1 and not args.no_cpu:
            cb = cpu_reference_arm(args.ref_particles, 3, 1)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "buckets_ms_per_update")}
            # the reference's shipped thread count (config/mcl_config.yaml:40 num_threads: 3), for context
            shipped = cpu_reference_arm(max(1000, args.ref_particles // 4), 2, 1, threads=3)
            cb["shipped_num_threads_3"] = {"value": shipped["value"], "unit": UNIT, "cores": 3, "sample": shipped["sample"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "updates_per_s": 1e3 / ms_per_step,
                "config": workload_config(world, N, R, args.shard_mode), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roof, "cpu_baseline": cb, "stage_ms": stage, "ray_stage": ray_stage,
                "wall_ms_per_step_incl_flush": 1e3 * wall / K,
                "pose_error_m": pose_err}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=N_PER_GPU, help="particles per GPU")
    ap.add_argument("--ref-particles", type=int, default=100000,
                    help="particles of the bounded CPU sample (reference arm / cpu_baseline)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ray-mode", type=int, default=0, choices=[0, 1, 2],
                    help="0 auto (default), 1 isotropic skip-map kernel only, 2 directional stage always")
    ap.add_argument("--shard-stages", action="store_true", help="multi-GPU: also report rank 0's per-stage times (extra replay)")
    ap.add_argument("--shard-mode", default="p2p", choices=["p2p", "allgather"],
                    help="multi-GPU exchange: NVLink peer reads of source poses + weight all-gather, or full all-gather")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)   # each step ~0.4 s of CPU work at the default --ref-particles
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
