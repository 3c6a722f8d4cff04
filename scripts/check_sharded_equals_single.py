#!/usr/bin/env python
"""Real multi-GPU parity: one particle-sharded filter over the ranks of a torchrun launch against
the same filter on ONE GPU (rank 0 runs it as well), same seed, device RNG, T updates.  Resample
indices, particles, raw and normalised weights must be bit-identical; poses equal up to the
rounding of the pose reduction (per-rank partial sums).  Prints one JSON line per case on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
      --master-port 29530 scripts/check_sharded_equals_single.py [--particles-per-gpu 262144]
      [--exchange fused|nccl] [--degenerate]
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
    ap.add_argument("--exchange", default="fused")
    ap.add_argument("--route", default="auto")
    ap.add_argument("--degenerate", action="store_true",
                    help="also one update from weights that put all the mass on one particle (overflow path of the exchange)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from monte_carlo_localization_b200 import MclContext, maps, synth
    from monte_carlo_localization_b200.sharded import ShardedFilter
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)          # NCCL banners must not reach the JSON stream
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        g = maps.load_named_map("Spielberg_map")
        angles_full = synth.laser_angles()
        angles = synth.downsample(angles_full)
        flt = ShardedFilter(g, angles, n_local=a.particles_per_gpu, rank=rank, world=world, device=local_rank, seed=99,
                            exchange=a.exchange, route=a.route)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    NG = a.particles_per_gpu * world
    gt, actions = synth.trajectory(g, a.updates + 1, 8.0)
    rng = np.random.default_rng(5)
    obs = np.stack([synth.scan_from_pose(flt.ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(a.updates + 1)]).astype(np.float32)
    # deliberately NO barrier between initialisation and the first update: initialisation is local and the
    # first cross-rank access is ordered by the exchange handshake itself (round 1's race, profiles/r2_*)
    flt.init_pose(gt[0])
    poses = [np.asarray(flt.update(actions[t], obs[t])).copy() for t in range(a.updates)]
    p, w = flt.gather_state()
    idx = flt.ctx.resample_indices()
    idx_all = [torch.empty(a.particles_per_gpu, dtype=torch.int32, device="cuda") for _ in range(world)]
    dist.all_gather(idx_all, torch.from_numpy(idx).cuda())
    idx_all = torch.cat(idx_all).cpu().numpy()
    degen = None
    if a.degenerate:
        wd = np.zeros(NG)
        wd[NG // 2 + 3] = 1.0
        flt.set_state(p, wd)
        pose_d = np.asarray(flt.update(actions[a.updates], obs[a.updates])).copy()
        pd, wdn = flt.gather_state()
        degen = (pose_d, pd, wdn)
    ok = True
    if rank == 0:
        single = MclContext(device=local_rank, max_particles=NG, seed=99)
        single.set_map(g)
        single.set_beam_angles(angles)
        single.set_graphs(False)
        single.init_pose(gt[0])
        sposes = [single.update(actions[t], obs[t]).copy() for t in range(a.updates)]
        sp, sw = single.get_particles(), single.get_weights()
        res = {"world": world, "exchange": a.exchange, "route": a.route, "particles": NG, "updates": a.updates,
               "indices_bit_identical": bool(np.array_equal(idx_all, single.resample_indices())),
               "weights_bit_identical": bool(np.array_equal(w, sw)), "particles_bit_identical": bool(np.array_equal(p, sp)),
               "max_pose_diff": float(np.abs(np.stack(poses) - np.stack(sposes)).max()),
               "weights_differing": int((w != sw).sum())}
        ok = res["weights_bit_identical"] and res["particles_bit_identical"] and res["indices_bit_identical"] and res["max_pose_diff"] < 1e-9
        if degen is not None:
            wd = np.zeros(NG)
            wd[NG // 2 + 3] = 1.0
            single.set_particles(sp, wd)
            sd_pose = single.update(actions[a.updates], obs[a.updates]).copy()
            res["degenerate_particles_bit_identical"] = bool(np.array_equal(degen[1], single.get_particles()))
            res["degenerate_weights_bit_identical"] = bool(np.array_equal(degen[2], single.get_weights()))
            res["degenerate_pose_diff"] = float(np.abs(degen[0] - sd_pose).max())
            ok = ok and res["degenerate_particles_bit_identical"] and res["degenerate_weights_bit_identical"]
        res["ok"] = bool(ok)
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
