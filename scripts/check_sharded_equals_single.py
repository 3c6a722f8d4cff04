#!/usr/bin/env python
"""Real multi-GPU parity: one particle-sharded filter over the ranks of a torchrun launch against
the same filter on ONE GPU (rank 0 runs it as well), same seed, device RNG, T updates.  Weights
must be bit-identical, particles bit-identical, poses equal up to the rounding of the pose
reduction (per-rank partial sums).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      --master-port 29530 scripts/check_sharded_equals_single.py [--particles-per-gpu 262144]
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
    ap.add_argument("--particles-per-gpu", type=int, default=262144)
    ap.add_argument("--updates", type=int, default=12)
    ap.add_argument("--shard-mode", default="p2p")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from monte_carlo_localization_b200 import MclContext, maps, synth
    from monte_carlo_localization_b200.sharded import ShardedFilter
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = maps.load_named_map("Spielberg_map")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    flt = ShardedFilter(g, angles, n_local=a.particles_per_gpu, rank=rank, world=world, device=local_rank, seed=99,
                        mode=a.shard_mode)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flt.ctx.set_stream(stream.cuda_stream)
    gt, actions = synth.trajectory(g, a.updates, 8.0)
    rng = np.random.default_rng(5)
    obs = np.stack([synth.scan_from_pose(flt.ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(a.updates)]).astype(np.float32)
    flt.init_pose(gt[0])
    if os.environ.get("CHECK_BARRIER_AFTER_INIT"):
        # every rank's initial state must be complete before any peer's first update reads it
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    poses = [np.asarray(flt.update(actions[t], obs[t])).copy() for t in range(a.updates)]
    p, w = flt.gather_state()
    ok = True
    if rank == 0:
        single = MclContext(device=local_rank, max_particles=a.particles_per_gpu * world, seed=99)
        single.set_map(g)
        single.set_beam_angles(angles)
        single.set_graphs(False)
        single.init_pose(gt[0])
        sposes = [single.update(actions[t], obs[t]).copy() for t in range(a.updates)]
        sp, sw = single.get_particles(), single.get_weights()
        res = {"world": world, "mode": flt.mode, "particles": a.particles_per_gpu * world, "updates": a.updates,
               "weights_bit_identical": bool(np.array_equal(w, sw)), "particles_bit_identical": bool(np.array_equal(p, sp)),
               "max_pose_diff": float(np.abs(np.stack(poses) - np.stack(sposes)).max()),
               "weights_differing": int((w != sw).sum())}
        ok = res["weights_bit_identical"] and res["particles_bit_identical"] and res["max_pose_diff"] < 1e-9
        res["ok"] = ok
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
