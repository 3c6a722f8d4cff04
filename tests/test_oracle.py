"""CPU tests of the oracle: against the committed golden vectors (made by the unmodified
reference), against the reference itself where oracle/_ref is built, and of the facts the CUDA
path relies on."""
import os

import numpy as np
import pytest

from helpers import load_golden, steps_from_ranges
from monte_carlo_localization_b200 import maps, synth
from oracle import bindings as ob

FIXTURES = ["update_sibal1_4000.npz", "update_Spielberg_map_2000.npz", "update_basement_fixed_1000.npz"]


@pytest.mark.parametrize("fixture", FIXTURES)
def test_oracle_reproduces_golden_vectors(fixture):
    z = load_golden(fixture)
    assert bool(z["pinned_by_reference"]), "golden vectors must come from the unmodified reference"
    g = maps.load_named_map(str(z["map"]))
    N = int(z["N"])
    orc = ob.Oracle(g, z["angles"], max_particles=N)
    orc.init_pose(z["gt"][0], z["z_init"])
    p, w = orc.get_state()
    assert np.array_equal(p, z["init_particles"]) and np.array_equal(w, z["init_weights"])
    for t in range(len(z["u"])):
        idx = orc.update(z["actions"][t], z["obs"][t], z["u"][t], z["z"][t])
        pose = orc.expected_pose()
        p, w = orc.get_state()
        assert np.array_equal(idx, z["idx"][t])
        assert np.array_equal(p, z["particles"][t])
        assert np.array_equal(w, z["weights"][t])
        assert np.array_equal(orc.raw_weights(), z["raw_weights"][t])
        assert np.array_equal(pose, z["pose"][t])
        assert np.array_equal(steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M), z["steps"][t])


@pytest.mark.parametrize("name,N,seed", [("sibal1", 2000, 11), ("first_map", 500, 12)])
def test_oracle_equals_unmodified_reference(name, N, seed):
    """Tier B == Tier A bit for bit: init, 4 updates, global init, single rays, table."""
    if not ob.have_reference():
        ob.build()   # compiles oracle/_ref where /root/reference exists
    if not ob.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    g = maps.load_named_map(name)
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    orc = ob.Oracle(g, angles, max_particles=N)
    ref = ob.Reference(g, seed, max_particles=N)
    assert ref.M == orc.M and ref.map_resolution() == g.resolution_f64
    assert np.array_equal(ref.sensor_table(), orc.sensor_table())
    gt, actions = synth.trajectory(g, 4, 3.0)
    rng = np.random.default_rng(seed)
    scans = [synth.scan_from_pose(orc.calc_range_many, gt[t + 1], angles_full, rng) for t in range(4)]
    assert ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scans[0]) == len(angles)
    assert np.array_equal(ref.beam_angles(), angles)
    ns = ob.NoiseStream(seed)
    ref.seed(seed)
    ref.init_pose(gt[0])
    orc.init_pose(gt[0], ns.normal(3 * N))
    for t in range(4):
        ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scans[t])
        assert np.array_equal(ref.downsampled_ranges(), scans[t][::18])
        u, zz = ns.update_noise(N)
        pose_r = ref.mcl(actions[t], scans[t][::18])
        orc.update(actions[t], scans[t][::18], u, zz)
        pr, wr = ref.get_state()
        po, wo = orc.get_state()
        assert np.array_equal(pr, po) and np.array_equal(wr, wo)
        assert np.array_equal(pose_r, orc.expected_pose())
        assert np.array_equal(ref.ranges(), orc.ranges().reshape(-1))
    # global initialisation draws (uniform_int + uniform_real interleaved, :433-441)
    ref.seed(seed + 1)
    ref.init_global()
    ns2 = ob.NoiseStream(seed + 1)
    cell, th = ns2.global_init(N, orc.num_free_cells())
    orc.init_global(cell, th)
    pr, wr = ref.get_state()
    po, wo = orc.get_state()
    assert np.array_equal(pr, po) and np.array_equal(wr, wo)
    # single rays, including from outside the map
    rng = np.random.default_rng(3)
    for _ in range(300):
        x = rng.uniform(g.origin[0] - 0.5, g.origin[0] + g.width * g.resolution_f64 + 0.5)
        y = rng.uniform(g.origin[1] - 0.5, g.origin[1] + g.height * g.resolution_f64 + 0.5)
        a = rng.uniform(-4, 4)
        assert ref.cast_ray(x, y, a) == orc.cast_ray(x, y, a)


def test_max_range_px_and_table_properties():
    """float32 resolution -> MAX_RANGE_PX 239 / 207; table columns sum to 1."""
    for name, M in (("sibal1", 239), ("Spielberg_map", 207), ("basement_fixed", 238)):
        g = maps.load_named_map(name)
        orc = ob.Oracle(g, synth.beam_angles(), max_particles=4)
        assert orc.M == M == g.max_range_px()
        t = orc.sensor_table().reshape(M + 1, M + 1)   # [d][r] (column-major table(r, d))
        assert np.allclose(t.sum(axis=1), 1.0, atol=1e-12)
        assert t.min() > 1e-5


def test_step_index_round_trip():
    """the table index the reference derives from a returned range equals the step
    count (hit) or MAX_RANGE_PX (no hit) -- the CUDA path carries integer steps."""
    for res32 in (np.float32(0.05), np.float32(0.05796), np.float32(0.0504), np.float32(0.1)):
        res = float(np.float64(res32))
        M = int(12.0 / res)
        r = np.arange(M)
        ranges = (r * res).astype(np.float32)
        assert np.array_equal(steps_from_ranges(ranges, res, M), r)
        assert steps_from_ranges(np.float32(12.0), res, M) == M


def test_resample_edge_cases():
    """discrete_distribution: fewer than two weights -> every draw is 0; zero weights are
    never selected; the last CDF entry is forced to 1."""
    u = np.array([0.0, 0.3, 0.999999], dtype=np.float64)
    assert np.array_equal(ob.resample_indices(np.array([5.0]), u), [0, 0, 0])
    idx, cdf = ob.resample_indices(np.array([0.0, 1.0, 0.0, 3.0]), u, want_cdf=True)
    # lower_bound quirk: u == 0.0 lands on index 0 even if its weight is zero; any u > 0 never
    # selects a zero-weight particle
    assert cdf[-1] == 1.0 and idx.tolist() == [0, 3, 3]
    w = np.array([0.25, 0.25, 0.5])
    idx = ob.resample_indices(w, np.array([0.25, 0.2500000001, 0.5, 0.75]))
    assert list(idx) == [0, 1, 1, 2]   # lower_bound: first cp >= u


def test_normalize_angle_matches_reference_loops():
    for a, want in ((0.0, 0.0), (4.0, 4.0 - 2 * np.pi), (-4.0, -4.0 + 2 * np.pi), (10.0, 10.0 - 4 * np.pi)):
        assert ob.normalize_angle(a) == pytest.approx(want, abs=1e-15)
    assert ob.normalize_angle(np.pi) == np.pi   # not wrapped: strict inequality


def test_noise_stream_is_deterministic_and_stateful():
    a, b = ob.NoiseStream(5), ob.NoiseStream(5)
    assert np.array_equal(a.canonical(10), b.canonical(10))
    # the normal distribution keeps its cached second value across calls (random.tcc:1812-1844)
    x = a.normal(3)
    y = np.concatenate([b.normal(1), b.normal(2)])
    assert np.array_equal(x, y)
    u = a.canonical(1000)
    assert (u >= 0).all() and (u < 1).all()


def test_oracle_resampling_equals_reference_viz_sampling():
    """visualize()'s weighted sub-sample (:946-958) draws discrete_distribution(weights_) from rng_: the oracle's
    lower_bound over its CDF with the twin generator's canonical uniforms picks the same particles."""
    if not ob.have_reference():
        pytest.skip("oracle/_ref absent")
    from monte_carlo_localization_b200 import maps
    g = maps.load_named_map("sibal1")
    N, k = 3000, 60
    rng = np.random.default_rng(2)
    w = rng.random(N) ** 4 + 1e-12
    w /= w.sum()
    p = rng.normal(size=(3, N))
    ref = ob.Reference(g, 77, max_particles=N, num_threads=2)
    ref.set_state(p, w)
    ref.seed(77)
    idx_ref = ref.viz_sample(k)
    u = ob.NoiseStream(77).canonical(k)
    assert np.array_equal(ob.resample_indices(w, u), idx_ref)
