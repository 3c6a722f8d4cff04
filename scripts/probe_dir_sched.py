"""Development aid: per-CTA counters of k_raycast_dir (a -DMCL_DIR_DIAG=1 build, MCL_B200_LIB=build/variants/...).
Runs a few updates of the bench workload and prints how the units, window stagings and scheduler time spread over CTAs."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from monte_carlo_localization_b200 import capi  # noqa: E402


def main():
    L = capi.load_library()
    fn = L.mcl_debug_dir_diag
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    from monte_carlo_localization_b200 import MclContext, maps, synth
    grid = maps.load_named_map("Spielberg_map")
    ctx = MclContext(max_particles=1 << 20, seed=20253)
    ctx.set_map(grid)
    ctx.set_beam_angles(synth.beam_angles())
    gt, actions, obs = bench.make_replay(ctx, grid, 12, seed=780, speed=8.0)
    ctx.init_pose(gt[0])
    for i in range(10):
        ctx.update(actions[i], obs[i])
    print(ctx.ray_stage_info())
    out = np.zeros((256, 8), dtype=np.uint64)
    assert fn(ctx._h, out.ctypes.data_as(C.POINTER(C.c_uint64))) == 0
    d = out[out[:, 0] > 0].astype(np.float64)
    g0 = d[:, 6].min()
    print("CTAs", len(d))
    for name, col, scale in (("cycles", 0, 1e-3), ("units", 1, 1), ("pieces", 2, 1), ("windows", 3, 1), ("sched kcyc", 4, 1e-3),
                             ("stage kcyc", 5, 1e-3)):
        v = d[:, col] * scale
        print("%-12s min %9.1f  mean %9.1f  max %9.1f  sum %11.1f" % (name, v.min(), v.mean(), v.max(), v.sum()))
    print("start spread us %.1f  end: min %.1f mean %.1f max %.1f us after the first start" % (
        (d[:, 6].max() - g0) / 1e3, (d[:, 7].min() - g0) / 1e3, (d[:, 7].mean() - g0) / 1e3, (d[:, 7].max() - g0) / 1e3))


if __name__ == "__main__":
    main()
