#!/usr/bin/env python
"""A few updates of a 2-rank sharded filter emulated on one GPU (host-ordered exchange), for ncu captures of
k_route / k_resample_motion(routed) themselves (a multi-rank launch cannot run under ncu)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import EmuRanks  # noqa: E402
from monte_carlo_localization_b200 import maps, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = maps.load_named_map("Spielberg_map")
angles = synth.beam_angles()
ranks = EmuRanks(g, angles, (1 << 20) * world, world, keep_ranges=False, seed=5)
gt, actions = synth.trajectory(g, 8, 8.0)
rng = np.random.default_rng(1)
obs = [synth.scan_from_pose(ranks.ctxs[0].calc_range_many, gt[t + 1], synth.laser_angles(), rng)[::18] for t in range(8)]
ranks.run(lambda r, c: c.init_pose(gt[0]))
for t in range(4):
    ranks.update(actions[t], obs[t])
print("done")
ranks.close()
