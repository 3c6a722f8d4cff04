#!/usr/bin/env python
"""BASELINE config 4 as specified: 1024 independent 4000-particle filters (multi-car replay) on
sibal1, FILTER-sharded over the GPUs of a box: rank r owns filters [r F/P, (r+1) F/P).  Filters
never interact, so there is no collective on the data path; only the timing is reduced (max over
ranks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
      --master-port 29521 scripts/run_config4_sharded.py [--filters 1024] [--steps 50]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from monte_carlo_localization_b200 import MclContext, maps, synth
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    F, N = a.filters // world, 4000
    g = maps.load_named_map("sibal1")
    angles_full = synth.laser_angles()
    ctx = MclContext(device=local_rank, max_particles=N, num_filters=F, seed=20254 + rank)
    ctx.set_map(g)
    ctx.set_beam_angles(synth.downsample(angles_full))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    total = a.steps + a.warmup
    gt, actions = synth.trajectory(g, total + a.filters, 3.0)
    rng = np.random.default_rng(781)
    obs = np.stack([synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(len(actions))]).astype(np.float32)
    first = rank * F                       # every car starts at its own phase of the lap
    for f in range(F):
        ctx.init_pose(gt[first + f], filter=f)
    d_act = torch.from_numpy(np.ascontiguousarray(actions)).cuda()
    d_obs = torch.from_numpy(obs).cuda()
    R = obs.shape[1]

    def step(t):
        idx = torch.arange(first, first + F, device="cuda") + t
        act = d_act[idx].contiguous()
        ob = d_obs[idx].contiguous()
        ctx.update_dev(act.data_ptr(), ob.data_ptr())
        return act, ob                     # keep the inputs alive until the kernels have run

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    keep = [step(t) for t in range(a.warmup)]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    keep = [step(t) for t in range(a.warmup, total)]
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    poses = ctx.read_pose().reshape(F, 3)
    want = gt[first + np.arange(F) + total]
    err = np.hypot(poses[:, 0] - want[:, 0], poses[:, 1] - want[:, 1])
    e = torch.tensor([float(np.median(err)), float((err < 0.3).mean())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.SUM)
        e /= world
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        print(json.dumps({"config": 4, "map": "sibal1", "n_gpus": world, "filters": F * world, "filters_per_gpu": F,
                          "particles_per_filter": N, "beams": R, "batch_steps": a.steps,
                          "ms_per_batch_step": 1e3 * sec / a.steps, "filter_updates_per_s": F * world * a.steps / sec,
                          "rays_per_s": F * world * N * R * a.steps / sec, "sharding": "filter-sharded, no collective",
                          "mean_of_rank_median_pose_err_m": float(e[0].item()), "frac_filters_within_0.3m": float(e[1].item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
