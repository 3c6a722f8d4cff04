#!/usr/bin/env python
"""Compute cost of the sharded filter's kernels WITHOUT the network and without rank skew: `world` ranks
emulated on ONE GPU (host-ordered exchange, tests/helpers.EmuRanks), 1 M particles per rank, per-kernel
CUDA-event times of rank 0.  Compared with the real multi-GPU kernel table this separates what k_route
costs to compute from what it waits for (NVLink stores, the slowest rank)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import EmuRanks  # noqa: E402
from monte_carlo_localization_b200 import maps, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_local = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
g = maps.load_named_map("Spielberg_map")
angles = synth.beam_angles()
ranks = EmuRanks(g, angles, n_local * world, world, keep_ranges=False, seed=5)
gt, actions = synth.trajectory(g, 10, 8.0)
rng = np.random.default_rng(1)
obs = [synth.scan_from_pose(ranks.ctxs[0].calc_range_many, gt[t + 1], synth.laser_angles(), rng)[::18] for t in range(10)]
ranks.run(lambda r, c: c.init_pose(gt[0]))
for t in range(3):
    ranks.update(actions[t], obs[t])
for c in ranks.ctxs:
    c.set_profiling(True)
rows = []
for t in range(3, 8):
    ranks.update(actions[t], obs[t])
    rows.append(dict(ranks.ctxs[0].kernel_ms()))
keys = list(rows[0].keys())
print(json.dumps({"world": world, "particles_per_rank": n_local, "emulated_on_one_gpu": True,
                  "kernel_ms_rank0_median": {k: float(np.median([r[k] for r in rows])) for k in keys}}))
ranks.close()
