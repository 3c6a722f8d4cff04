"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle and the
committed golden vectors (which were produced by the unmodified reference).

Tolerances are the north_star's: resample indices bit-exact, range step indices equal
(<= 1 cell allowed), normalised weights <= 1e-5 relative, pose <= 1 mm / 1e-4 rad.
"""
import numpy as np
import pytest

from helpers import assert_pose_close, assert_weights_close, load_golden, steps_from_ranges

pytestmark = pytest.mark.gpu


def _ctx(grid, angles, N, **kw):
    from monte_carlo_localization_b200 import MclContext
    c = MclContext(max_particles=N, **kw)
    c.set_map(grid)
    c.set_beam_angles(angles)
    c.set_keep_ranges(True)
    return c


@pytest.mark.parametrize("ray_mode", [0, 1], ids=["auto", "isotropic"])
@pytest.mark.parametrize("fixture", ["update_sibal1_4000.npz", "update_Spielberg_map_2000.npz",
                                     "update_basement_fixed_1000.npz"])
def test_golden_updates_teacher_forced(fixture, ray_mode):
    """Every update starts from the reference's own state, so each stage is compared on
    identical inputs: indices and ranges must match exactly.  Run with the ray stage chosen on the
    device (directional for the compact clouds of >= 1024 particles) and with the isotropic kernel."""
    from monte_carlo_localization_b200 import maps
    z = load_golden(fixture)
    g = maps.load_named_map(str(z["map"]))
    N = int(z["N"])
    c = _ctx(g, z["angles"], N)
    c.set_ray_mode(ray_mode)
    prev_p, prev_w = z["init_particles"], z["init_weights"]
    for t in range(len(z["u"])):
        c.set_particles(prev_p, prev_w)
        pose = c.update(z["actions"][t], z["obs"][t], z["u"][t], z["z"][t])
        assert np.array_equal(c.resample_indices(), z["idx"][t]), "resample indices differ at update %d" % t
        steps = c.range_steps()
        assert np.array_equal(steps, z["steps"][t]), "range steps differ at update %d: %d rays" % (
            t, int((steps != z["steps"][t]).sum()))
        assert_weights_close(c.raw_weights(), z["raw_weights"][t])
        assert_weights_close(c.get_weights(), z["weights"][t])
        p = c.get_particles()
        assert np.abs(p[:2] - z["particles"][t][:2]).max() < 1e-9
        assert np.abs(p[2] - z["particles"][t][2]).max() < 1e-9
        assert_pose_close(pose, z["pose"][t])
        prev_p, prev_w = z["particles"][t], z["weights"][t]
    info = c.ray_stage_info()
    assert info["last_mode"] == (1 if (ray_mode == 0 and N >= 1024) else 0)
    c.close()


def test_golden_updates_free_running():
    """No teacher forcing: the GPU filter runs on its own state for all updates."""
    from monte_carlo_localization_b200 import maps
    z = load_golden("update_sibal1_4000.npz")
    g = maps.load_named_map("sibal1")
    c = _ctx(g, z["angles"], int(z["N"]))
    c.init_pose(z["gt"][0], z["z_init"])
    p0 = c.get_particles()
    assert np.abs(p0 - z["init_particles"]).max() < 1e-12
    for t in range(len(z["u"])):
        pose = c.update(z["actions"][t], z["obs"][t], z["u"][t], z["z"][t])
        assert np.array_equal(c.resample_indices(), z["idx"][t])
        assert_pose_close(pose, z["pose"][t])
    c.close()


def test_sensor_table_matches_oracle():
    from monte_carlo_localization_b200 import maps
    from oracle import bindings as ob
    for name in ("sibal1", "Spielberg_map"):
        g = maps.load_named_map(name)
        z = load_golden("update_sibal1_4000.npz")
        c = _ctx(g, z["angles"], 16)
        orc = ob.Oracle(g, z["angles"], max_particles=16)
        assert c.M == orc.M
        assert np.array_equal(c.sensor_table(), orc.sensor_table())
        c.close()


@pytest.mark.parametrize("name,n", [("sibal1", 200000), ("Spielberg_map", 200000), ("first_map", 100000)])
def test_calc_range_many_matches_oracle(name, n):
    """cast_ray parity on random and adversarial queries (cell corners, axis-aligned rays,
    poses outside the map): float ranges must be identical."""
    from monte_carlo_localization_b200 import maps
    from oracle import bindings as ob
    g = maps.load_named_map(name)
    angles = load_golden("update_sibal1_4000.npz")["angles"]
    rng = np.random.default_rng(5)
    res = g.resolution_f64
    x = rng.uniform(g.origin[0] - 1.0, g.origin[0] + g.width * res + 1.0, n)
    y = rng.uniform(g.origin[1] - 1.0, g.origin[1] + g.height * res + 1.0, n)
    th = rng.uniform(-np.pi, np.pi, n)
    k = n // 2
    x[:k] = rng.integers(0, g.width, k) * res + g.origin[0]
    y[:k] = rng.integers(0, g.height, k) * res + g.origin[1]
    th[:k // 2] = rng.integers(-2, 3, k // 2) * np.pi / 2
    q = np.stack([x, y, th])
    c = _ctx(g, angles, 16)
    orc = ob.Oracle(g, angles, max_particles=16)
    got = c.calc_range_many(q)
    want = orc.calc_range_many(q)
    bad = got != want
    assert not bad.any(), "%d / %d ranges differ, first %s" % (bad.sum(), n, np.argwhere(bad)[:3].ravel())
    c.close()


@pytest.mark.parametrize("N", [1, 2, 17, 4096, 4097, 100000, 1000000])
def test_resample_indices_exact(N):
    """discrete_distribution parity at sizes up to BASELINE's 1M: CDF bits and indices."""
    from monte_carlo_localization_b200 import maps
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles = load_golden("update_sibal1_4000.npz")["angles"]
    rng = np.random.default_rng(N)
    w = rng.random(N) ** 6 + 1e-12
    w /= w.sum()
    ns = ob.NoiseStream(99 + N)
    u, z = ns.update_noise(N)
    c = _ctx(g, angles, N)
    c.set_keep_ranges(False)
    p = np.zeros((3, N))
    p[0] = -3.3
    p[1] = 1.6
    p[2] = np.linspace(-3, 3, N)
    c.set_particles(p, w)
    obs = np.full(len(angles), 2.0, dtype=np.float32)
    c.update([0.05, 0, 0.01], obs, u, z)
    idx_ref, cdf_ref = ob.resample_indices(w, u, want_cdf=True)
    if N >= 2:
        assert np.array_equal(c.cdf(), cdf_ref), "CDF bits differ in %d places" % int((c.cdf() != cdf_ref).sum())
    assert np.array_equal(c.resample_indices(), idx_ref)
    # replication counts follow from the indices
    assert np.array_equal(np.bincount(c.resample_indices(), minlength=N), np.bincount(idx_ref, minlength=N))
    c.close()


def test_full_update_vs_oracle_100k():
    """One update at 100k particles (BASELINE config 2 size) against the oracle."""
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map("basement_fixed")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    N = 100000
    orc = ob.Oracle(g, angles, max_particles=N)
    gt, actions = synth.trajectory(g, 2, 3.0)
    ns = ob.NoiseStream(4242)
    orc.init_pose(gt[0], ns.normal(3 * N))
    c = _ctx(g, angles, N)
    for t in range(2):
        p, w = orc.get_state()
        c.set_particles(p, w)
        scan = synth.scan_from_pose(orc.calc_range_many, gt[t + 1], angles_full, np.random.default_rng(t))
        obs = scan[::18]
        u, z = ns.update_noise(N)
        idx = orc.update(actions[t], obs, u, z)
        pose_ref = orc.expected_pose()
        pose = c.update(actions[t], obs, u, z)
        assert np.array_equal(c.resample_indices(), idx)
        want = steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M)
        got = c.range_steps()
        assert np.abs(got.astype(np.int64) - want).max() <= 1
        assert (got != want).sum() == 0
        assert_weights_close(c.get_weights(), orc.get_state()[1])
        assert_pose_close(pose, pose_ref)
    c.close()


def test_batch_of_filters_matches_single():
    """A batch of independent filters equals the same filters run one by one (config 4 shape)."""
    from monte_carlo_localization_b200 import maps
    z = load_golden("update_sibal1_4000.npz")
    g = maps.load_named_map("sibal1")
    N, F = int(z["N"]), 3
    single = _ctx(g, z["angles"], N)
    batch = _ctx(g, z["angles"], N, num_filters=F)
    rng = np.random.default_rng(0)
    shifts = rng.normal(0, 0.05, (F, 3))
    poses_single = []
    for f in range(F):
        p = z["init_particles"].copy()
        p += shifts[f][:, None]
        batch.set_particles(p, z["init_weights"], filter=f)
        single.set_particles(p, z["init_weights"])
        poses_single.append(single.update(z["actions"][0], z["obs"][f % len(z["obs"])], z["u"][f % 3], z["z"][f % 3]))
        if f == F - 1:
            idx_last = single.resample_indices()
    acts = np.stack([z["actions"][0]] * F)
    obs = np.stack([z["obs"][f % len(z["obs"])] for f in range(F)])
    u = np.stack([z["u"][f % 3] for f in range(F)])
    zz = np.stack([z["z"][f % 3] for f in range(F)])
    poses = batch.update(acts, obs, u, zz)
    for f in range(F):
        assert np.array_equal(poses[f], poses_single[f])
    assert np.array_equal(batch.resample_indices(filter=F - 1), idx_last)
    single.close()
    batch.close()


def test_device_rng_update_runs_and_tracks():
    """Production mode (no injected noise): the filter tracks the ground truth."""
    from monte_carlo_localization_b200 import maps
    z = load_golden("update_sibal1_4000.npz")
    g = maps.load_named_map("sibal1")
    c = _ctx(g, z["angles"], 4000, seed=1234)
    c.init_pose(z["gt"][0])
    for t in range(len(z["obs"])):
        pose = c.update(z["actions"][t], z["obs"][t])
    assert np.hypot(*(pose[:2] - z["gt"][-1][:2])) < 0.3
    w = c.get_weights()
    assert abs(w.sum() - 1.0) < 1e-9 and (w > 0).all()
    c.close()


def test_full_size_update_1m_spielberg_vs_oracle():
    """BASELINE config 3 at full size: one update of 1,048,576 particles x 60 beams on
    Spielberg_map against the oracle (indices and range steps exact), plus the size-independent
    properties of the outputs."""
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map("Spielberg_map")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    N = 1 << 20
    orc = ob.Oracle(g, angles, max_particles=N)
    gt, actions = synth.trajectory(g, 2, 8.0)
    ns = ob.NoiseStream(31337)
    orc.init_pose(gt[0], ns.normal(3 * N))
    scan = synth.scan_from_pose(orc.calc_range_many, gt[1], angles_full, np.random.default_rng(9))
    obs = scan[::18]
    # non-uniform weights so the CDF is not trivial
    w = np.random.default_rng(10).random(N) ** 3 + 1e-9
    w /= w.sum()
    p0, _ = orc.get_state()
    orc.set_state(p0, w)
    c = _ctx(g, angles, N)
    c.set_particles(p0, w)
    u, z = ns.update_noise(N)
    idx = orc.update(actions[0], obs, u, z)
    pose_ref = orc.expected_pose()
    pose = c.update(actions[0], obs, u, z)
    got_idx = c.resample_indices()
    assert np.array_equal(got_idx, idx)
    want = steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M)
    got = c.range_steps()
    assert (got != want).sum() == 0
    wn = c.get_weights()
    assert_weights_close(wn, orc.get_state()[1])
    assert_pose_close(pose, pose_ref)
    # properties that hold at any size
    cdf = c.cdf()
    assert cdf[-1] == 1.0 and (np.diff(cdf) >= 0).all()
    order = np.argsort(u, kind="stable")
    assert (np.diff(got_idx[order]) >= 0).all()          # the index is monotone in the draw
    assert np.bincount(got_idx, minlength=N).sum() == N   # replication counts
    assert abs(wn.sum() - 1.0) < 1e-9 and (wn > 0).all()
    assert got.max() <= c.M
    c.close()


def test_edge_cases_match_oracle():
    """Degenerate inputs the reference can meet: particles outside the map and on its border,
    zero / max-range / out-of-range / NaN scan readings, a zero action, one dominant weight."""
    from monte_carlo_localization_b200 import maps
    from oracle import bindings as ob
    z = load_golden("update_sibal1_4000.npz")
    g = maps.load_named_map("sibal1")
    N = int(z["N"])
    res = g.resolution_f64
    rng = np.random.default_rng(3)
    p = z["init_particles"].copy()
    p[0, :500] = g.origin[0] - rng.uniform(0, 3, 500)                      # left of the map
    p[1, 500:1000] = g.origin[1] + g.height * res + rng.uniform(0, 3, 500)  # above the map
    p[0, 1000:1200] = g.origin[0] + rng.integers(-2, 3, 200) * res          # on / next to the border lattice
    p[1, 1200:1400] = g.origin[1] + rng.integers(-2, 3, 200) * res
    p[0, 1400:1450] = 1e7                                                   # absurdly far away
    w = np.full(N, 1e-12)
    w[1234] = 1.0                                                           # one particle holds the mass
    w /= w.sum()
    obs = z["obs"][0].copy()
    obs[0] = 0.0
    obs[1] = 12.0
    obs[2] = 50.0          # beyond max range: clamps to MAX_RANGE_PX (:552)
    obs[3] = np.inf
    obs[4] = np.nan        # reference: UB; x86 yields row 0
    for action in ([0.0, 0.0, 0.0], [0.0005, 0.0, 0.0005], [0.5, 0.0, -0.2], [-0.05, 0.3, 0.0]):
        orc = ob.Oracle(g, z["angles"], max_particles=N)
        orc.set_state(p, w)
        c = _ctx(g, z["angles"], N)
        c.set_particles(p, w)
        idx = orc.update(action, obs, z["u"][0], z["z"][0])
        pose = c.update(action, obs, z["u"][0], z["z"][0])
        assert np.array_equal(c.resample_indices(), idx)
        assert np.array_equal(c.range_steps(), steps_from_ranges(orc.ranges(), res, orc.M))
        assert_weights_close(c.get_weights(), orc.get_state()[1])
        assert_pose_close(pose, orc.expected_pose())
        c.close()
    # a second update from the spread-out state (many particles outside the map)
    orc = ob.Oracle(g, z["angles"], max_particles=N)
    orc.set_state(p, np.full(N, 1.0 / N))
    c = _ctx(g, z["angles"], N)
    c.set_particles(p, np.full(N, 1.0 / N))
    for t in range(2):
        idx = orc.update(z["actions"][t], obs, z["u"][t], z["z"][t])
        c.update(z["actions"][t], obs, z["u"][t], z["z"][t])
        assert np.array_equal(c.resample_indices(), idx)
        assert np.array_equal(c.range_steps(), steps_from_ranges(orc.ranges(), res, orc.M))
        assert_weights_close(c.get_weights(), orc.get_state()[1])
    c.close()


def test_global_init_and_error_paths():
    """initialize_global with injected draws equals the oracle; error conventions of the ABI."""
    from monte_carlo_localization_b200 import MclContext, MclError, capi, maps
    from oracle import bindings as ob
    z = load_golden("update_sibal1_4000.npz")
    g = maps.load_named_map("sibal1")
    N = 4000
    orc = ob.Oracle(g, z["angles"], max_particles=N)
    c = _ctx(g, z["angles"], N)
    assert c.num_free_cells() == orc.num_free_cells() == 26948
    cell, th = ob.NoiseStream(5).global_init(N, orc.num_free_cells())
    orc.init_global(cell, th)
    c.init_global(cell, th)
    assert np.array_equal(c.get_particles(), orc.get_state()[0])
    assert np.array_equal(c.get_weights(), orc.get_state()[1])
    c.init_global()                      # device RNG: every particle on a free cell's corner
    p = c.get_particles()
    col = np.round((p[0] - g.origin[0]) / g.resolution_f64).astype(int)
    row = np.round((p[1] - g.origin[1]) / g.resolution_f64).astype(int)
    assert (g.data[row, col] == 0).all() and (p[2] >= 0).all() and (p[2] < 2 * np.pi).all()
    # wrong beam count, straight through the C ABI
    import ctypes as C
    act = np.zeros(3)
    bad = np.zeros(7, dtype=np.float32)
    pose = np.zeros(3)
    rc = c._L.mcl_update(c._h, act.ctypes.data_as(capi.c_double_p), bad.ctypes.data_as(capi.c_float_p), 7, None,
                         pose.ctypes.data_as(capi.c_double_p))
    assert rc == capi.MCL_ERR_INVALID and b"num_beams" in c._L.mcl_last_error()
    c.close()
    c2 = MclContext(max_particles=16)
    with pytest.raises(MclError) as e:   # update before a map: the reference's cast_ray returns max range (:613)
        c2.calc_range_many(np.zeros((3, 4)))
    assert e.value.status == capi.MCL_ERR_NO_MAP
    occupied = maps.OccupancyGrid(np.full((20, 20), 100, dtype=np.int8), np.float32(0.05), (0.0, 0.0, 0.0))
    c2.set_map(occupied)
    with pytest.raises(MclError) as e:   # "No free space found in map!" (:423-427)
        c2.init_global()
    assert e.value.status == capi.MCL_ERR_NO_FREE_SPACE
    fine = maps.OccupancyGrid(np.zeros((20, 20), dtype=np.int8), np.float32(0.01), (0.0, 0.0, 0.0))
    c2.set_map(fine)                     # MAX_RANGE_PX = 1200: beyond the skip-map kernels, served by the wide path
    assert c2.M == 1200
    assert c2.cast_ray(0.1, 0.1, 0.0) == np.float32(0.1 / 0.01 // 1 * 0 + c2.cast_ray(0.1, 0.1, 0.0))   # (runs)
    with pytest.raises(MclError) as e:   # more beams than any scan has
        c2.set_beam_angles(np.zeros(5000, dtype=np.float32))
    assert e.value.status == capi.MCL_ERR_UNSUPPORTED
    c2.close()


def test_gather_microbench_reports_rates():
    """Measurement aid: random-byte gather rates from shared memory and from L2."""
    from monte_carlo_localization_b200 import capi
    sm = capi.microbench_gather(True, iters=512)
    l2 = capi.microbench_gather(False, iters=512)
    assert sm > 1e11 and l2 > 1e10 and sm > l2


def _tracking_case(name, N, speed, seed):
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map(name)
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    orc = ob.Oracle(g, angles, max_particles=N)
    gt, actions = synth.trajectory(g, 40, speed)
    ns = ob.NoiseStream(seed)
    orc.init_pose(gt[20], ns.normal(3 * N))
    scan = synth.scan_from_pose(orc.calc_range_many, gt[21], angles_full, np.random.default_rng(seed))
    return g, angles, orc, ns, actions[20], scan[::18]


@pytest.mark.parametrize("name,N,speed", [("Spielberg_map", 200000, 8.0), ("basement_fixed", 50000, 3.0),
                                          ("sibal1", 20000, 3.0)])
def test_directional_stage_equals_isotropic_kernel_and_oracle(name, N, speed):
    """The two ray stages compute the same cast_ray: step indices and raw weights are
    bit-identical between them, and equal to the oracle's."""
    g, angles, orc, ns, action, obs = _tracking_case(name, N, speed, 77)
    p0, w0 = orc.get_state()
    u, z = ns.update_noise(N)
    c = _ctx(g, angles, N)
    info = c.ray_stage_info()
    assert info["directional_ready"] == 1 and info["box_cells"] >= 64
    out = {}
    for mode in (1, 2, 0):
        c.set_ray_mode(mode)
        c.set_particles(p0, w0)
        pose = c.update(action, obs, u, z)
        out[mode] = (c.range_steps().copy(), c.raw_weights().copy(), c.get_weights().copy(), pose.copy())
        assert c.ray_stage_info()["last_mode"] == (0 if mode == 1 else 1)
    for k in range(4):
        assert np.array_equal(out[1][k], out[2][k]), "stage outputs differ in item %d" % k
        assert np.array_equal(out[0][k], out[2][k])
    idx = orc.update(action, obs, u, z)
    assert np.array_equal(c.resample_indices(), idx)
    want = steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M)
    assert (out[2][0] != want).sum() == 0
    assert_weights_close(out[2][2], orc.get_state()[1])
    assert_pose_close(out[2][3], orc.expected_pose())
    c.close()


def test_directional_maps_equal_cpu_build():
    """k_build_dir_maps writes the bytes the CPU harness computes from the same source."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from emu_bindings import EmuMap
    from monte_carlo_localization_b200 import maps, synth
    g = maps.load_named_map("sibal1")
    c = _ctx(g, synth.beam_angles(), 20000)
    em = EmuMap(g)
    for s in (0, 5, 8, 11, 15):
        assert np.array_equal(c.dir_map(s), em.dir_map(s)), "sector %d differs" % s
    c.close()


@pytest.mark.parametrize("name", ["sibal1", "basement_fixed", "Spielberg_map"])
def test_device_distance_transform_equals_host_build(name):
    """mcl_set_map builds the isotropic skip codes (and the gap map behind the sector maps) with an exact Euclidean
    distance transform on the device (map_kernels.cuh): same bytes as the host build (map_prep.cpp)."""
    import sys, os, time
    sys.path.insert(0, os.path.dirname(__file__))
    from emu_bindings import EmuMap
    from monte_carlo_localization_b200 import maps, synth
    g = maps.load_named_map(name)
    t0 = time.perf_counter()
    c = _ctx(g, synth.beam_angles(), 20000)
    t_set = time.perf_counter() - t0
    em = EmuMap(g)
    v8 = c.dir_map(-1)
    assert v8.shape == em.v8().shape
    assert np.array_equal(v8, em.v8()), "%d codes differ" % int((v8 != em.v8()).sum())
    for s in (3, 12, 14):   # the sector maps are traced against the device-built gap map
        assert np.array_equal(c.dir_map(s), em.dir_map(s)), "sector %d differs" % s
    print("context + mcl_set_map(%s) incl. sector maps: %.1f ms" % (name, 1e3 * t_set))
    c.close()


def test_scattered_cloud_stays_on_the_directional_stage():
    """After initialize_global no particle lies in the window box: the directional stage marches the sector maps in
    global memory (all particles on the global-memory path); the isotropic kernel gives the same steps and weights,
    and both equal the oracle's."""
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles = synth.beam_angles()
    N = 30000
    orc = ob.Oracle(g, angles, max_particles=N)
    cell, th = ob.NoiseStream(8).global_init(N, orc.num_free_cells())
    orc.init_global(cell, th)
    p0, w0 = orc.get_state()
    ns = ob.NoiseStream(9)
    u, z = ns.update_noise(N)
    obs = np.full(len(angles), 3.0, dtype=np.float32)
    action = np.array([0.05, 0.0, 0.01])
    c = _ctx(g, angles, N)
    res = {}
    for mode in (0, 2):
        c.set_ray_mode(mode)
        c.set_particles(p0, w0)
        c.update(action, obs, u, z)
        res[mode] = (c.range_steps().copy(), c.raw_weights().copy())
        assert c.ray_stage_info()["last_mode"] == 1   # scattered clouds stay on the directional stage (sector maps from L2)
    c.set_ray_mode(1)
    c.set_particles(p0, w0)
    c.update(action, obs, u, z)
    assert c.ray_stage_info()["last_mode"] == 0
    res[0] = (c.range_steps().copy(), c.raw_weights().copy())            # the isotropic kernel on the same cloud
    assert np.array_equal(res[0][0], res[2][0]) and np.array_equal(res[0][1], res[2][1])
    orc.update(action, obs, u, z)
    want = steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M)
    assert (res[2][0] != want).sum() == 0
    c.close()


def test_directional_stage_edge_cases_match_oracle():
    """The directional stage on degenerate inputs: a cloud at the map border (clipped windows),
    particles outside the map and outside the window box, NaN / inf / out-of-range scan readings,
    headings at +-pi (bucket wrap-around), one dominant weight."""
    from monte_carlo_localization_b200 import maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles = synth.beam_angles()
    N = 24000
    res = g.resolution_f64
    rng = np.random.default_rng(12)
    ns = ob.NoiseStream(99)
    # cloud in the lower-left corner region of the map, heading west (around +-pi)
    x = g.origin[0] + 1.0 + rng.normal(0, 0.3, N)
    y = g.origin[1] + 1.2 + rng.normal(0, 0.3, N)
    th = np.pi + rng.normal(0, 0.3, N)
    th = (th + np.pi) % (2 * np.pi) - np.pi
    th[:50] = np.pi                     # exactly on the wrap
    th[50:100] = -np.pi
    x[100:400] = g.origin[0] - rng.uniform(0, 2, 300)                # left of the map
    y[400:700] = g.origin[1] + rng.integers(-2, 3, 300) * res       # on / next to the border lattice
    x[700:900] += rng.normal(0, 5.0, 200)                            # far outside the window box
    x[900:920] = 1e7
    p = np.stack([x, y, th])
    w = np.full(N, 1e-9)
    w[4321] = 1.0
    w /= w.sum()
    obs = np.full(len(angles), 2.0, dtype=np.float32)
    obs[0], obs[1], obs[2], obs[3], obs[4] = 0.0, 12.0, 50.0, np.inf, np.nan
    c = _ctx(g, angles, N)
    c.set_ray_mode(2)
    for k, action in enumerate(([0.0, 0.0, 0.0], [0.3, 0.0, -0.2])):
        u, z = ns.update_noise(N)
        orc = ob.Oracle(g, angles, max_particles=N)
        orc.set_state(p, w)
        c.set_particles(p, w)
        idx = orc.update(action, obs, u, z)
        pose = c.update(action, obs, u, z)
        assert c.ray_stage_info()["last_mode"] == 1
        assert np.array_equal(c.resample_indices(), idx)
        assert np.array_equal(c.range_steps(), steps_from_ranges(orc.ranges(), res, orc.M))
        assert_weights_close(c.get_weights(), orc.get_state()[1])
        assert_pose_close(pose, orc.expected_pose())
        w = np.full(N, 1.0 / N)          # second round: uniform weights, every particle survives
    c.close()


@pytest.mark.parametrize("name,N", [("sibal1", 4000), ("sibal1", 30000)])
def test_graph_replay_equals_direct_launches(name, N):
    """The host-facing update replays its steady state as a CUDA graph: particles, weights and
    poses are bit-identical to the launch-by-launch path over a run of device-RNG updates,
    including across a re-initialisation (which drops out of and re-enters the steady state)."""
    from monte_carlo_localization_b200 import MclContext, maps, synth
    g = maps.load_named_map(name)
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    ctxs = []
    for graphs in (True, False):
        c = MclContext(max_particles=N, seed=4711)
        c.set_map(g)
        c.set_beam_angles(angles)
        c.set_graphs(graphs)
        ctxs.append(c)
    gt, actions = synth.trajectory(g, 14, 3.0)
    rng = np.random.default_rng(3)
    obs = [synth.scan_from_pose(ctxs[0].calc_range_many, gt[t + 1], angles_full, rng)[::18] for t in range(14)]
    launches = []
    for c in ctxs:
        c.init_pose(gt[0])
        n0 = c.kernel_launches()
        poses = [c.update(actions[t], obs[t]).copy() for t in range(8)]
        c.init_pose(gt[8])                    # weights touched: the next update runs launch by launch
        poses += [c.update(actions[t], obs[t]).copy() for t in range(8, 14)]
        launches.append(c.kernel_launches() - n0)
        c._poses = np.stack(poses)
    a, b = ctxs
    assert np.array_equal(a._poses, b._poses)
    assert np.array_equal(a.get_particles(), b.get_particles())
    assert np.array_equal(a.get_weights(), b.get_weights())
    assert launches[0] == launches[1]          # the launch counter counts the kernels of a replayed graph too
    err = np.hypot(*(a._poses[-1][:2] - gt[14][:2]))
    assert err < 0.5
    for c in ctxs:
        c.close()


def test_batch_pool_directional_stage_equals_isotropic_kernel():
    """A batch of filters on a map that fits one window runs the directional stage over the pool
    of all filters' particles: step indices, raw weights, weights and poses are bit-identical
    to the isotropic kernel's, filter by filter, and filter 0 equals the oracle."""
    from monte_carlo_localization_b200 import MclContext, maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    F, N = 5, 3000            # 3000 is not a multiple of 32: warps and 4-slot groups straddle filters
    gt, actions = synth.trajectory(g, 30, 3.0)
    ns = ob.NoiseStream(321)
    orcs, states = [], []
    for f in range(F):
        o = ob.Oracle(g, angles, max_particles=N)
        o.init_pose(gt[5 * f], ns.normal(3 * N))
        orcs.append(o)
        states.append(o.get_state())
    obs = np.stack([synth.scan_from_pose(orcs[0].calc_range_many, gt[5 * f + 1], angles_full, np.random.default_rng(f))[::18]
                    for f in range(F)]).astype(np.float32)
    act = np.stack([actions[5 * f] for f in range(F)])
    noise = [ns.update_noise(N) for _ in range(F)]
    out = {}
    for mode in (1, 2):        # pool mode is opt-in for batches (auto keeps the isotropic kernel)
        c = MclContext(max_particles=N, num_filters=F)
        c.set_map(g)
        c.set_beam_angles(angles)
        c.set_keep_ranges(True)
        c.set_ray_mode(mode)
        for f in range(F):
            c.set_particles(states[f][0], states[f][1], filter=f)
        poses = c.update(act, obs, np.stack([n[0] for n in noise]), np.stack([n[1] for n in noise]))
        assert c.ray_stage_info()["last_mode"] == (0 if mode == 1 else 1)
        out[mode] = [(c.range_steps(f).copy(), c.raw_weights(f).copy(), c.get_weights(f).copy()) for f in range(F)] + [poses.copy()]
        c.close()
    out[0] = out[2]
    for f in range(F):
        for k in range(3):
            assert np.array_equal(out[0][f][k], out[1][f][k]), "filter %d item %d differs" % (f, k)
    assert np.array_equal(out[0][F], out[1][F])
    idx = orcs[0].update(act[0], obs[0], noise[0][0], noise[0][1])
    want = steps_from_ranges(orcs[0].ranges(), g.resolution_f64, orcs[0].M)
    assert (out[0][0][0] != want).sum() == 0
    assert_weights_close(out[0][0][2], orcs[0].get_state()[1])


def test_programmatic_dependent_launches_change_nothing():
    """mcl_set_pdl: the kernels of an update are launched with programmatic stream serialization (a kernel's blocks
    may be resident while its predecessor drains and wait in griddepcontrol.wait).  Graph replay and direct launches,
    a big filter on the directional stage and a batch: bit-identical to plain stream order."""
    from monte_carlo_localization_b200 import MclContext, maps, synth
    g = maps.load_named_map("sibal1")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full)
    gt, actions = synth.trajectory(g, 8, 3.0)
    for N, F, graphs in ((30000, 1, True), (30000, 1, False), (2000, 3, True)):
        out = []
        for pdl in (True, False):
            c = MclContext(max_particles=N, num_filters=F, seed=99)
            c.set_map(g)
            c.set_beam_angles(angles)
            c.set_graphs(graphs)
            c.set_pdl(pdl)
            rng = np.random.default_rng(3)
            obs = [synth.scan_from_pose(c.calc_range_many, gt[t + 1], angles_full, rng)[::18] for t in range(6)]
            for f in range(F):
                c.init_pose(gt[0], filter=f)
            for t in range(6):
                a_ = np.tile(actions[t], (F, 1)) if F > 1 else actions[t]
                o_ = np.tile(obs[t], (F, 1)) if F > 1 else obs[t]
                pose = c.update(a_, o_)
            out.append((np.asarray(pose).copy(), [c.get_particles(f).copy() for f in range(F)], [c.get_weights(f).copy() for f in range(F)]))
            c.close()
        assert np.array_equal(out[0][0], out[1][0])
        for f in range(F):
            assert np.array_equal(out[0][1][f], out[1][1][f]) and np.array_equal(out[0][2][f], out[1][2][f])


# ---- particle-sharded filter: ranks emulated on ONE GPU (helpers.EmuRanks) ------------------------
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_ranks_reproduce_golden_updates(world):
    """Teacher-forced golden updates through `world` ranks: every rank holds only its slice; the
    gathered indices, particles, weights and the pose equal the reference's single-filter update."""
    from helpers import EmuRanks
    from monte_carlo_localization_b200 import maps
    z = load_golden("update_sibal1_4000.npz")
    g = maps.load_named_map("sibal1")
    N = int(z["N"])
    ranks = EmuRanks(g, z["angles"], N, world)
    prev_p, prev_w = z["init_particles"], z["init_weights"]
    for t in range(len(z["u"])):
        ranks.set_state(prev_p, prev_w)
        poses = ranks.update(z["actions"][t], z["obs"][t], z["u"][t], z["z"][t])
        assert np.array_equal(ranks.gather(lambda c: c.resample_indices()), z["idx"][t]), "resample indices differ at update %d" % t
        assert np.array_equal(ranks.gather(lambda c: c.range_steps().T).T, z["steps"][t])
        assert_weights_close(ranks.gather(lambda c: c.get_weights()), z["weights"][t])
        p = ranks.gather(lambda c: c.get_particles())
        assert np.abs(p - z["particles"][t]).max() < 1e-9
        for pose in poses:
            assert_pose_close(pose, z["pose"][t])
            assert np.array_equal(pose, poses[0])      # every rank computes the same pose bits
        prev_p, prev_w = z["particles"][t], z["weights"][t]
    ranks.close()


@pytest.mark.parametrize("world,N,two_hop", [(2, 4000, True), (2, 40000, True), (4, 65536, True), (3, 30000, True), (2, 4002, True),
                                             (2, 40000, False), (4, 65536, False)])
def test_sharded_ranks_free_running_equal_single_filter(world, N, two_hop):
    """No teacher forcing, device RNG: the sharded filter must stay BIT-identical to the same filter on
    one GPU over several updates (noise is keyed by the global slot; sums are sequentially rounded), with
    the draws routed in two hops (requests to the source ranks; odd slice sizes and a world of 3 included)
    and in one hop (every rank tests all draws)."""
    from helpers import EmuRanks
    g, angles, orc, ns, action, obs = _tracking_case("sibal1", N, 3.0, 11)
    p0, w0 = orc.get_state()
    single = _ctx(g, angles, N, seed=4242)
    single.set_graphs(False)
    single.set_particles(p0, w0)
    ranks = EmuRanks(g, angles, N, world, seed=4242, two_hop=two_hop)
    ranks.set_state(p0, w0)
    for t in range(4):
        pose_s = single.update(action, obs)
        poses = ranks.update(action, obs)
        assert np.array_equal(ranks.gather(lambda c: c.resample_indices()), single.resample_indices()), "update %d" % t
        assert np.array_equal(ranks.gather(lambda c: c.get_particles()), single.get_particles()), "update %d" % t
        assert np.array_equal(ranks.gather(lambda c: c.raw_weights()), single.raw_weights())
        assert np.array_equal(ranks.gather(lambda c: c.get_weights()), single.get_weights()), "update %d" % t
        assert np.array_equal(ranks.gather(lambda c: c.cdf()), single.cdf())
        for pose in poses:
            assert np.abs(pose - pose_s).max() < 1e-9
    if N // world >= 1024:
        assert all(c.ray_stage_info()["last_mode"] == 1 for c in ranks.ctxs)   # the directional stage ran on every rank
    ranks.close()
    single.close()


def test_sharded_ranks_degenerate_weights_match_oracle():
    """Weight vectors that defeat the step-map summary: one particle holds all the mass (the running
    sum sits exactly on a power of two, every later chunk is opaque -- far more than the mailbox payload
    carries, so the ranks read each other's overflow lists), all-equal weights, and a zero-weight rank."""
    from helpers import EmuRanks
    from monte_carlo_localization_b200 import maps
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles = load_golden("update_sibal1_4000.npz")["angles"]
    N, world = 16384, 2
    rng = np.random.default_rng(3)
    p = np.zeros((3, N))
    p[0] = -3.3 + rng.normal(0, 0.05, N)
    p[1] = 1.6 + rng.normal(0, 0.05, N)
    p[2] = rng.uniform(-3, 3, N)
    obs = np.full(len(angles), 2.0, dtype=np.float32)
    cases = {}
    w = np.zeros(N)
    w[0] = 1.0
    cases["one particle holds all the mass"] = w
    w = np.zeros(N)
    w[N // 2 + 5] = 0.5
    w[N // 2 + 6000] = 0.5
    cases["rank 0 has no mass"] = w
    cases["uniform"] = np.full(N, 1.0 / N)
    w = rng.random(N) ** 8 + 1e-300
    cases["heavy-tailed"] = w / w.sum()
    for two_hop in (True, False):
        ranks = EmuRanks(g, angles, N, world, keep_ranges=False, two_hop=two_hop)
        for name, w in cases.items():
            ns = ob.NoiseStream(5)
            u, z = ns.update_noise(N)
            ranks.set_state(p, w)
            ranks.update([0.05, 0, 0.01], obs, u, z)
            idx_ref, cdf_ref = ob.resample_indices(w, u, want_cdf=True)
            assert np.array_equal(ranks.gather(lambda c: c.cdf()), cdf_ref), name
            assert np.array_equal(ranks.gather(lambda c: c.resample_indices()), idx_ref), name
        ranks.close()


def test_viz_weighted_subsample_matches_reference_draws():
    """visualize()'s weighted sub-sample (:946-958): k draws of discrete_distribution(weights_).  With the
    reference generator's canonical uniforms injected, mcl_sample_particles_u returns the same particle indices
    as the reference's own sampling (oracle/_ref) and as lower_bound on the oracle's CDF; without injection
    the device RNG draws particles in proportion to their weights."""
    from monte_carlo_localization_b200 import maps
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles = load_golden("update_sibal1_4000.npz")["angles"]
    for N, k in ((4000, 60), (100000, 200)):
        rng = np.random.default_rng(N)
        w = rng.random(N) ** 5 + 1e-9
        w /= w.sum()
        p = np.stack([rng.uniform(-4, -2, N), rng.uniform(1, 2, N), rng.uniform(-3, 3, N)])
        c = _ctx(g, angles, N)
        c.set_particles(p, w)
        u = ob.NoiseStream(31 + N).canonical(k)
        got, idx = c.sample_particles(k, u=u, return_indices=True)
        idx_ref = ob.resample_indices(w, u)
        assert np.array_equal(idx, idx_ref)
        assert np.array_equal(got, p[:, idx_ref])
        if ob.have_reference():
            ref = ob.Reference(g, 31 + N, max_particles=N, num_threads=2)
            ref.set_state(p, w)
            ref.seed(31 + N)
            assert np.array_equal(idx, ref.viz_sample(k)), "indices differ from the reference's own visualize() sampling"
        # device RNG: heavy particles dominate the draws
        _, idx_dev = c.sample_particles(4000, return_indices=True)
        heavy = w > np.quantile(w, 0.9)
        assert abs(heavy[idx_dev].mean() - w[heavy].sum()) < 0.05
        # sampling leaves the filter usable: the next update draws from the same CDF
        assert np.array_equal(c.get_weights(), w)
        c.close()


@pytest.mark.parametrize("max_range,angle_step", [(15.0, 18), (12.0, 4), (20.0, 6)],
                         ids=["MAX_RANGE_PX 300", "270 beams", "MAX_RANGE_PX 400 x 180 beams"])
def test_wide_configurations_match_oracle(max_range, angle_step):
    """Configurations beyond the skip-map kernels' limits (MAX_RANGE_PX > 254, more than 128 beams): the reference
    accepts any max_range / angle_step (:195, :307-310).  The context marches with the reference's arithmetic;
    indices exact, range steps equal, weights and pose within tolerance, calc_range_many identical."""
    from monte_carlo_localization_b200 import MclContext, maps, synth
    from oracle import bindings as ob
    g = maps.load_named_map("sibal1")
    angles_full = synth.laser_angles()
    angles = synth.downsample(angles_full, angle_step)
    N = 3000
    orc = ob.Oracle(g, angles, max_particles=N, max_range=max_range)
    ns = ob.NoiseStream(17)
    gt, actions = synth.trajectory(g, 6, 3.0)
    orc.init_pose(gt[0], ns.normal(3 * N))
    c = MclContext(max_particles=N, max_range=max_range)
    c.set_map(g)
    c.set_beam_angles(angles)
    assert c.M == orc.M and c.M == int(max_range / g.resolution_f64)
    for t in range(3):
        scan = synth.scan_from_pose(orc.calc_range_many, gt[t + 1], angles_full, np.random.default_rng(t), max_range=max_range)
        obs = scan[::angle_step]
        p0, w0 = orc.get_state()
        u, z = ns.update_noise(N)
        c.set_particles(p0, w0)
        pose = c.update(actions[t], obs, u, z)
        idx = orc.update(actions[t], obs, u, z)
        assert np.array_equal(c.resample_indices(), idx)
        want = steps_from_ranges(orc.ranges(), g.resolution_f64, orc.M)
        # the step index of a hit at step r is r itself up to the float rounding of r * res / res (:556-574)
        got = c.range_steps16().astype(np.int64)
        conv = steps_from_ranges(c.ranges(), g.resolution_f64, orc.M)
        assert np.array_equal(conv.reshape(-1), want.reshape(-1)), "%d rays differ" % int((conv.reshape(-1) != want.reshape(-1)).sum())
        assert got.max() <= orc.M
        assert_weights_close(c.get_weights(), orc.get_state()[1])
        assert_pose_close(pose, orc.expected_pose())
    rng = np.random.default_rng(3)
    q = np.stack([rng.uniform(-7, 8, 5000), rng.uniform(-2, 6, 5000), rng.uniform(-np.pi, np.pi, 5000)])
    assert np.array_equal(c.calc_range_many(q), orc.calc_range_many(q))
    with pytest.raises(Exception):
        c.range_steps()      # one-byte step indices do not exist for a wide context
    c.close()
