#!/usr/bin/env python
"""BASELINE config 5 as specified: global (kidnapped-robot) initialisation, 16 M particles uniform
over the free space of the levine stand-in map, ONE filter particle-sharded over the GPUs of a box;
reports the updates and the wall time until the pose estimate stays within 0.25 m / 0.1 rad of the
ground truth for 10 updates.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
      --master-port 29520 scripts/run_config5_sharded.py [--particles-per-gpu 2097152]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles-per-gpu", type=int, default=2097152)
    ap.add_argument("--updates", type=int, default=60)
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"])
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from monte_carlo_localization_b200 import MclContext, maps, synth
    from monte_carlo_localization_b200.sharded import ShardedFilter
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    g = maps.load_named_map("basement_fixed")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        flt = ShardedFilter(g, angles, n_local=a.particles_per_gpu, rank=rank, world=world, device=local_rank,
                            seed=20255, exchange=a.exchange)
        ctx = flt.ctx
    else:
        flt = None
        ctx = MclContext(device=local_rank, max_particles=a.particles_per_gpu, seed=20255)
        ctx.set_map(g)
        ctx.set_beam_angles(angles)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    gt, actions = synth.trajectory(g, a.updates, 3.0)
    rng = np.random.default_rng(782)
    obs = np.stack([synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(a.updates)]).astype(np.float32)
    (flt or ctx).init_global()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    streak, conv, times, modes = 0, None, [], []
    for t in range(a.updates):
        t0 = time.perf_counter()
        pose = (flt or ctx).update(actions[t], obs[t])
        times.append(time.perf_counter() - t0)
        modes.append(ctx.ray_stage_info()["last_mode"])
        d = float(np.hypot(*(np.asarray(pose)[:2] - gt[t + 1][:2])))
        dth = abs((pose[2] - gt[t + 1][2] + np.pi) % (2 * np.pi) - np.pi)
        streak = streak + 1 if (d < 0.25 and dth < 0.1) else 0
        if streak == 10 and conv is None:
            conv = t + 1 - 9
    if rank == 0:
        n_global = a.particles_per_gpu * world
        print(json.dumps({
            "config": 5, "map": "basement_fixed (levine stand-in: levine.pgm is missing from the reference checkout)",
            "n_gpus": world, "particles": n_global, "beams": len(angles), "free_cells": ctx.num_free_cells(),
            "sharding": "single GPU" if world == 1 else "particle-sharded x%d, slice-local state, %s exchange" % (world, flt.exchange),
            "updates_run": a.updates, "converged_at_update": conv,
            "ms_per_update_first5": 1e3 * float(np.mean(times[:5])), "ms_per_update_last5": 1e3 * float(np.mean(times[-5:])),
            "time_to_converge_s": None if conv is None else float(np.sum(times[:conv + 9])),
            "rays_per_s_last5": n_global * len(angles) / float(np.mean(times[-5:])),
            "final_pose_err_m": d, "directional_stage_from_update": (modes.index(1) + 1) if 1 in modes else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
