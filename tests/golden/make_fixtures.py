"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box only reads the outputs).

1. maps/<name>.npz  -- occupancy grids decoded from the reference's maps/*.yaml with the
   nav2 trinary rule (monte_carlo_localization_b200/maps.py); stored decoded so neither the
   GPU tests nor bench.py need /root/reference.
2. update_<map>_<N>.npz -- known-answer vectors of the hot path: inputs (state, action, scan,
   injected noise) and outputs (resample indices, range step indices, weights, pose) of
   consecutive MCL updates, produced by the UNMODIFIED reference (oracle/_ref) where it is
   built, and cross-checked bit-for-bit against the oracle restatement before being written.

Usage: python tests/golden/make_fixtures.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from monte_carlo_localization_b200 import maps, synth  # noqa: E402
from oracle import bindings as ob  # noqa: E402

REF_MAPS = "/root/reference/maps"
OUT = os.path.join(ROOT, "tests", "golden")

MAPS = {"sibal1": "sibal1.yaml", "Spielberg_map": "Spielberg_map.yaml",
        "basement_fixed": "basement_fixed.map.yaml", "first_map": "first_map.yaml"}


def make_maps():
    for name, yml in MAPS.items():
        g = maps.load_map_yaml(os.path.join(REF_MAPS, yml))
        g.name = name
        maps.save_grid_npz(os.path.join(OUT, "maps", name + ".npz"), g)
        print("map", name, g.width, g.height, g.resolution_f64, g.max_range_px(), g.counts())


def steps_from_ranges(r, res, M, max_range=12.0):
    r = np.asarray(r, dtype=np.float32)
    px = (r.astype(np.float64) / res).astype(np.float32)
    px = np.minimum(px, np.float32(M))
    return np.clip(np.round(px).astype(np.int64), 0, M).astype(np.uint8)


def make_update_vectors(map_name, N, n_updates, seed_filter, seed_scan, speed):
    g = maps.load_named_map(map_name)
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    orc = ob.Oracle(g, angles, max_particles=N)
    gt, actions = synth.trajectory(g, n_updates, speed)
    rng = np.random.default_rng(seed_scan)
    scans = np.stack([synth.scan_from_pose(orc.calc_range_many, gt[t + 1], angles_full, rng)
                      for t in range(n_updates)])
    obs = scans[:, ::18]
    ns = ob.NoiseStream(seed_filter)
    z_init = ns.normal(3 * N)
    orc.init_pose(gt[0], z_init)
    ref = None
    if ob.have_reference():
        ref = ob.Reference(g, seed_filter, max_particles=N)
        ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scans[0])
        assert np.array_equal(ref.beam_angles(), angles)
        ref.seed(seed_filter)
        ref.init_pose(gt[0])
    p0, w0 = orc.get_state()
    if ref is not None:
        pr, wr = ref.get_state()
        assert np.array_equal(pr, p0) and np.array_equal(wr, w0), "init mismatch oracle vs reference"
    out = dict(map=np.asarray(map_name), N=N, angles=angles, init_particles=p0, init_weights=w0,
               gt=gt, actions=actions, obs=obs, z_init=z_init,
               sensor_table_sha256=np.asarray(hashlib.sha256(orc.sensor_table().tobytes()).hexdigest()))
    U, Z, IDX, STEPS, W, WRAW, POSE, PART = [], [], [], [], [], [], [], []
    for t in range(n_updates):
        u, z = ns.update_noise(N)
        idx = orc.update(actions[t], obs[t], u, z)
        pose = orc.expected_pose()
        p, w = orc.get_state()
        if ref is not None:
            ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scans[t])
            pose_r = ref.mcl(actions[t], obs[t])
            pr, wr = ref.get_state()
            assert np.array_equal(pr, p) and np.array_equal(wr, w) and np.array_equal(pose_r, pose), \
                "update %d: oracle != reference" % t
            assert np.array_equal(ref.ranges(), orc.ranges().reshape(-1))
        U.append(u); Z.append(z); IDX.append(idx)
        STEPS.append(steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M))
        W.append(w); WRAW.append(orc.raw_weights()); POSE.append(pose); PART.append(p)
    out.update(u=np.stack(U), z=np.stack(Z), idx=np.stack(IDX), steps=np.stack(STEPS),
               weights=np.stack(W), raw_weights=np.stack(WRAW), pose=np.stack(POSE),
               particles=np.stack(PART), pinned_by_reference=np.asarray(ref is not None),
               mean_cells_per_ray=np.asarray(orc.mean_cells_per_ray()))
    path = os.path.join(OUT, "update_%s_%d.npz" % (map_name, N))
    np.savez_compressed(path, **out)
    print("wrote", path, "pinned_by_reference=%s" % (ref is not None), "C-bar=%.1f" % orc.mean_cells_per_ray(),
          "pose err %.3f m" % np.hypot(*(POSE[-1][:2] - gt[-1][:2])))


if __name__ == "__main__":
    make_maps()
    make_update_vectors("sibal1", 4000, 3, 20251, 778, 3.0)         # BASELINE config 1
    make_update_vectors("Spielberg_map", 2000, 2, 20253, 780, 8.0)   # config-3 map at KAT size
    make_update_vectors("basement_fixed", 1000, 2, 20252, 779, 3.0)  # config-2 stand-in
