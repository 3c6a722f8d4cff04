#!/usr/bin/env python
"""Scattered cloud (initialize_global, every update re-initialised so it never converges): isotropic kernel against the
directional stage marching the sector maps in global memory (ray mode 2).  One GPU."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monte_carlo_localization_b200 import MclContext, maps, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
for name in ("basement_fixed", "Spielberg_map"):
    g = maps.load_named_map(name)
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    ctx = MclContext(device=0, max_particles=N, seed=7)
    ctx.set_map(g)
    ctx.set_beam_angles(angles)
    ctx.set_graphs(False)
    gt, actions = synth.trajectory(g, 8, 3.0)
    rng = np.random.default_rng(1)
    obs = synth.scan_from_pose(ctx.calc_range_many, gt[1], angles_full, rng)[::18].astype(np.float32)
    out = {"map": name, "particles": N}
    for mode in (1, 2):
        ctx.set_ray_mode(mode)
        ts = []
        for it in range(5):
            ctx.init_global()
            ctx.set_profiling(True)
            ctx.update(actions[0], obs)
            km = dict(ctx.kernel_ms())
            ctx.set_profiling(False)
            ts.append((sum(km.values()), km.get("k_raycast_dir", 0.0) + km.get("k_raycast_weight", 0.0) + km.get("k_weight_steps", 0.0)))
        out["mode%d" % mode] = {"update_ms": float(np.median([t[0] for t in ts])), "ray_ms": float(np.median([t[1] for t in ts])),
                                "stage": ctx.ray_stage_info()}
    print(json.dumps(out))
    ctx.close()
