#!/usr/bin/env python
"""The routing kernels (k_route_request / k_route_serve, or k_route with MCL_ROUTE=one-hop) on real GPUs (torchrun, one rank per GPU): SM-cycle stamps of its phases on every rank
(mcl_debug_pass_cycles kind 8) and rank 0's per-kernel CUDA-event times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from monte_carlo_localization_b200 import maps, synth  # noqa: E402
from monte_carlo_localization_b200.sharded import ShardedFilter  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
sys.stdout.flush()
saved = os.dup(1)
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
g = maps.load_named_map("Spielberg_map")
angles = synth.beam_angles()
flt = ShardedFilter(g, angles, n_local=1 << 20, rank=rank, world=world, device=lr, seed=5,
                    route=os.environ.get("MCL_ROUTE", "two-hop"))
dist.barrier()
sys.stdout.flush()
os.dup2(saved, 1)
ctx = flt.ctx
gt, actions = synth.trajectory(g, 16, 8.0)
rng = np.random.default_rng(1)
obs = [synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], synth.laser_angles(), rng)[::18] for t in range(16)]
flt.init_pose(gt[0])
ctx.set_graphs(False)
for t in range(4):
    flt.update(actions[t], obs[t])
phases = {}
for kind, name in ((8, "serve"), (9, "request")):
    ctx.debug_pass_cycles(kind)
    rows = []
    for t in range(4, 10):
        flt.update(actions[t], obs[t])
        rows.append(ctx.debug_pass_cycles(kind, read=True)[:3])
    ctx.debug_pass_cycles(-1)
    phases[name] = (np.median(np.asarray(rows, dtype=np.float64), axis=0) / 1965.0).tolist()
ctx.set_profiling(True)
flt.update(actions[10], obs[10])
flt.update(actions[11], obs[11])
km = ctx.kernel_ms()
out = [None] * world
dist.all_gather_object(out, {"rank": rank, "us [slowest CTA's work, with the system fence, last CTA's wait for the peers]": phases,
                             "kernel_ms": {k: v for k, v in km if "route" in k}})
if rank == 0:
    print(json.dumps({"world": world, "k_route_phases": out, "kernel_ms_rank0": km}))
dist.destroy_process_group()
