#!/usr/bin/env python
"""Where does an exact-sum pass spend its time?  SM-cycle stamps of the phases of k_exact_pass
(mcl_debug_pass_cycles) on the bench workload, plus the per-kernel CUDA-event times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monte_carlo_localization_b200 import MclContext, maps, synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
g = maps.load_named_map("Spielberg_map")
angles = synth.beam_angles()
ctx = MclContext(max_particles=N, seed=5)
ctx.set_map(g)
ctx.set_beam_angles(angles)
gt, actions = synth.trajectory(g, 16, 8.0)
rng = np.random.default_rng(1)
obs = [synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], synth.laser_angles(), rng)[::18] for t in range(16)]
ctx.init_pose(gt[0])
ctx.set_graphs(False)
for t in range(4):
    ctx.update(actions[t], obs[t])
names = ["slowest_cta_tile_phase", "last_cta_until_last", "pose_fold", "tile_scan", "opaque_list", "exchange", "eval_and_tile_starts",
         "opaque_chunks"]
for kind, label in ((0, "S1"), (1, "normalise+pose+S2"), (3, "cdf")):
    ctx.debug_pass_cycles(kind)
    rows = []
    for t in range(4, 10):
        ctx.update(actions[t], obs[t])
        rows.append(ctx.debug_pass_cycles(kind, read=True))
    med = np.median(np.asarray(rows, dtype=np.float64), axis=0)
    print(json.dumps({"pass": label, "particles": N, "cycles": dict(zip(names, [float(v) for v in med])),
                      "us_at_1965MHz": {n: round(float(v) / 1965.0, 2) for n, v in zip(names[:-1], med[:-1])}}))
ctx.debug_pass_cycles(-1)
ctx.set_profiling(True)
ctx.update(actions[10], obs[10])
ctx.update(actions[11], obs[11])
print(json.dumps({"kernel_ms": ctx.kernel_ms()}))
