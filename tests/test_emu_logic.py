"""The algorithms the CUDA kernels run (csrc/march.cuh, csrc/exact_sum.cuh, csrc/map_prep.cpp),
compiled for the CPU by tests/emu and checked against the oracle -- so the kernel LOGIC is
verified in the GPU-less container.  The kernels themselves are checked by tests/test_gpu_parity.py."""
import numpy as np
import pytest

from emu_bindings import EmuMap, exact_scan
from helpers import steps_from_ranges
from monte_carlo_localization_b200 import maps, synth
from oracle import bindings as ob


def _seq_cumsum(a):
    return np.add.accumulate(np.asarray(a, dtype=np.float64))   # strictly sequential


CASES = {
    "pow8": lambda rng, n: rng.random(n) ** 8 + 1e-300,
    "uniform": lambda rng, n: np.full(n, 1.0 / n),
    "lognormal": lambda rng, n: np.exp(rng.normal(size=n) * 20),
    "spike": lambda rng, n: np.where(np.arange(n) == n // 2, 1e6, rng.random(n)),
    "onehot": lambda rng, n: np.where(np.arange(n) == min(77, n - 1), 1.0, 0.0),
    "stuck_at_binade": lambda rng, n: np.concatenate([[0.5 - 2.0 ** -40], np.full(n - 1, 2.0 ** -60)]),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 4000, 4096, 4097, 65536, 300001])
def test_exact_sum_and_cdf_equal_sequential(case, n):
    rng = np.random.default_rng(n)
    w = np.ascontiguousarray(CASES[case](rng, n), dtype=np.float64)
    seq = _seq_cumsum(w)
    total, prefix, _ = exact_scan(w)
    assert total == seq[-1]
    assert np.array_equal(prefix, seq)
    if seq[-1] > 0:
        p = w / seq[-1]
        cp = _seq_cumsum(p)
        cp[-1] = 1.0
        _, cdf, _ = exact_scan(w, div=float(seq[-1]), force_last_one=True)
        assert np.array_equal(cdf, cp)


def test_exact_cdf_matches_libstdcpp_discrete_distribution():
    rng = np.random.default_rng(1)
    for n in (2, 100, 4000, 100000):
        w = rng.random(n) ** 4 + 1e-9
        u = rng.random(2000)
        idx_ref, cdf_ref = ob.resample_indices(w, u, want_cdf=True)
        s, _, _ = exact_scan(w, want_prefix=False)
        _, cdf, _ = exact_scan(w, div=s, force_last_one=True)
        assert np.array_equal(cdf, cdf_ref)
        assert np.array_equal(np.searchsorted(cdf, u, side="left"), idx_ref)


def test_plain_parallel_scan_would_not_be_exact():
    """Why the exact machinery exists: a chunked parallel-style scan differs in the last bits."""
    rng = np.random.default_rng(2)
    w = rng.random(100000)
    seq = _seq_cumsum(w)
    chunked = np.cumsum(np.cumsum(w.reshape(-1, 16), axis=1)[:, -1])
    assert (chunked != seq[15::16]).any()


def _stress_poses(g, n, seed, angles):
    rng = np.random.default_rng(seed)
    res = g.resolution_f64
    x = rng.uniform(g.origin[0] - 1.0, g.origin[0] + g.width * res + 1.0, n)
    y = rng.uniform(g.origin[1] - 1.0, g.origin[1] + g.height * res + 1.0, n)
    th = rng.uniform(-np.pi, np.pi, n)
    k = n // 2   # cell corners (initialize_global poses) and axis-aligned rays stress the edge rule
    x[:k] = rng.integers(0, g.width, k) * res + g.origin[0]
    y[:k] = rng.integers(0, g.height, k) * res + g.origin[1]
    th[:k // 2] = rng.integers(-2, 3, k // 2) * np.pi / 2 - angles[rng.integers(0, len(angles), k // 2)].astype(np.float64)
    return x, y, th


@pytest.mark.parametrize("name", ["sibal1", "Spielberg_map", "basement_fixed", "first_map"])
@pytest.mark.parametrize("use_window", [False, True])
def test_skip_map_march_equals_reference_march(name, use_window):
    g = maps.load_named_map(name)
    angles = synth.beam_angles()
    em = EmuMap(g)
    n = 6000
    x, y, th = _stress_poses(g, n, 3, angles)
    orc = ob.Oracle(g, angles, max_particles=n)
    q = np.stack([np.repeat(x, len(angles)), np.repeat(y, len(angles)),
                  (th[:, None] + angles[None, :].astype(np.float64)).reshape(-1)])
    want = steps_from_ranges(orc.calc_range_many(q), g.resolution_f64, orc.M).reshape(n, len(angles))
    window = None
    if use_window:
        ww, wh = min(em.PW, 640), min(em.PH, 640)
        wx0 = max(0, min(em.PW - ww, em.PW // 2 - ww // 2)) & ~31
        wy0 = max(0, min(em.PH - wh, em.PH // 2 - wh // 2))
        window = (wx0, wy0, ww, wh)
    got, replays = em.range_steps(x, y, th, angles, window)
    assert em.M == orc.M
    assert np.array_equal(got.astype(np.int64), want), "%d rays differ" % int((got != want).sum())
    assert replays > 0   # the stress poses do exercise the exact-replay rule


def test_skip_map_codes_are_conservative():
    """Every code's promise holds: no blocked cell within adv cells of a cell of code 1+adv."""
    g = maps.load_named_map("sibal1")
    em = EmuMap(g)
    v8 = em.v8().astype(np.int32)
    blocked = v8 == 0
    from scipy import ndimage
    dil = ndimage.binary_dilation(blocked, structure=np.ones((3, 3), bool), border_value=1)
    d = ndimage.distance_transform_edt(~dil)
    free = v8 >= 2
    adv = v8 - 1
    assert (adv[free] < d[free] + 1.0 + 1e-9).all()
    assert ((v8 == 1) == (dil & ~blocked)).all()
    v4 = em.v4()
    lo, hi = v4 & 15, v4 >> 4
    assert np.array_equal(lo, np.minimum(v8[:, 0::2], 15)) and np.array_equal(hi, np.minimum(v8[:, 1::2], 15))


def _oracle_steps(g, angles, x, y, th):
    n = len(x)
    orc = ob.Oracle(g, angles, max_particles=n)
    q = np.stack([np.repeat(x, len(angles)), np.repeat(y, len(angles)),
                  (th[:, None] + angles[None, :].astype(np.float64)).reshape(-1)])
    return steps_from_ranges(orc.calc_range_many(q), g.resolution_f64, orc.M).reshape(n, len(angles))


@pytest.mark.parametrize("name,buckets", [("sibal1", 2048), ("sibal1", 4096), ("first_map", 2048), ("first_map", 4096),
                                          ("basement_fixed", 4096), ("Spielberg_map", 2048)])
def test_directional_march_equals_reference_march(name, buckets):
    """k_raycast_dir's logic (heading bucket -> sector map -> march_ray_dir) on stress poses:
    cell corners, axis-aligned rays, poses outside the map; every sector gets used."""
    g = maps.load_named_map(name)
    angles = synth.beam_angles()
    em = EmuMap(g)
    n = 4000
    x, y, th = _stress_poses(g, n, 11, angles)
    want = _oracle_steps(g, angles, x, y, th)
    got, replays = em.range_steps_dir(x, y, th, angles, buckets=buckets)
    assert np.array_equal(got.astype(np.int64), want), "%d rays differ" % int((got != want).sum())
    assert replays > 0


def test_directional_window_covers_a_tracking_cloud():
    """Tracking cloud on Spielberg: particles inside the box march a bounds-checked copy of each
    sector's window (a read outside it fails the test), stragglers the whole sector map; the
    directional maps need far fewer lookups than the isotropic skip map."""
    full = maps.load_named_map("Spielberg_map")
    angles = synth.beam_angles()
    gt, _ = synth.trajectory(full, 60, 8.0)
    pose = gt[30]
    # crop 640 x 640 cells around the pose (the CPU build of the sector maps of the whole map is slow)
    res = full.resolution_f64
    c0 = max(0, int((pose[0] - full.origin[0]) / res) - 320)
    r0 = max(0, int((pose[1] - full.origin[1]) / res) - 320)
    g = maps.OccupancyGrid(np.ascontiguousarray(full.data[r0:r0 + 640, c0:c0 + 640]), full.resolution,
                           (full.origin[0] + c0 * res, full.origin[1] + r0 * res, 0.0))
    em = EmuMap(g)
    rng = np.random.default_rng(5)
    n = 1500
    x = pose[0] + rng.normal(0, 0.15, n)
    y = pose[1] + rng.normal(0, 0.15, n)
    th = pose[2] + rng.normal(0, 0.25, n)
    x[:20] += rng.normal(0, 6.0, 20)          # stragglers far outside the box
    y[:20] += rng.normal(0, 6.0, 20)
    want = _oracle_steps(g, angles, x, y, th)
    got, replays, lk = em.range_steps_dir(x, y, th, angles, buckets=4096, window_box=64, want_lookups=True)
    assert replays >= 0, "%d reads fell outside a sector window" % -replays
    assert np.array_equal(got.astype(np.int64), want), "%d rays differ" % int((got != want).sum())
    cbar = np.where(want < em.M, want + 1, em.M).mean()
    assert lk.mean() < 0.1 * cbar     # an order of magnitude fewer lookups than cells sampled


def test_directional_codes_respect_their_promise():
    """Brute force on sibal1: from any sub-cell position and any direction of the sector, the
    adv-1 samples after a cell of advance adv are not blocked."""
    g = maps.load_named_map("sibal1")
    em = EmuMap(g)
    v8 = em.v8()
    rng = np.random.default_rng(9)
    from emu_bindings import dir_constants
    n_sec = dir_constants()[0]
    width = 2 * np.pi / n_sec
    for s in sorted({0, 3, 8 % n_sec, 13 % n_sec, 21 % n_sec, n_sec - 1}):
        d = em.dir_map(s)
        assert ((d == 0x80) == (v8 == 0)).all()
        assert (((d & 0x80) != 0) == (v8 < 2)).all()
        assert ((d & 0x7f)[v8 > 0] >= 1).all()
        ys, xs = np.nonzero((v8 >= 1) & ((d & 0x7f) > 1))
        pick = rng.choice(len(ys), size=min(3000, len(ys)), replace=False)
        for cy, cx in zip(ys[pick], xs[pick]):
            adv = int(d[cy, cx] & 0x7f)
            a = rng.uniform(s * width - 0.002, (s + 1) * width + 0.002, 6)
            px = cx + rng.random(6)
            py = cy + rng.random(6)
            t = np.arange(1, adv)[:, None]
            sx = np.floor(px[None, :] + t * np.cos(a)[None, :]).astype(int)
            sy = np.floor(py[None, :] + t * np.sin(a)[None, :]).astype(int)
            assert (v8[sy, sx] != 0).all(), (s, cx, cy, adv)


def test_sector_of_a_ray_contains_its_direction():
    """Host arithmetic of the directional stage: the sector chosen from (heading bucket, beam)
    by integer arithmetic contains the ray's real direction up to the map's validity margin, for
    every admissible bucket count, headings on the +-pi seam and arbitrary beam tables."""
    from emu_bindings import dir_constants, dir_sector
    S, margin, min_buckets = dir_constants()
    assert S in (16, 32) and min_buckets == 2048
    width = 2 * np.pi / S
    rng = np.random.default_rng(21)
    thetas = np.concatenate([rng.uniform(-np.pi, np.pi, 4000), [np.pi, -np.pi, 0.0, np.nextafter(np.pi, 0), -np.nextafter(np.pi, 0)],
                             (np.arange(64) - 32) * (2 * np.pi / 64)])
    alphas = np.concatenate([synth.beam_angles(), rng.uniform(-3.2, 3.2, 40).astype(np.float32), np.float32([0.0, 3.1415927, -3.1415927])])
    for B in (2048, 4096):
        worst = 0.0
        for th in thetas:
            for al in alphas[rng.integers(0, len(alphas), 12)]:
                s = dir_sector(th, al, B)
                a = (th + float(al)) % (2 * np.pi)
                lo, hi = s * width, (s + 1) * width
                # cyclic distance of a to the interval [lo, hi]
                d = 0.0 if lo <= a <= hi else min(abs(a - lo), abs(a - hi), abs(a - lo - 2 * np.pi), abs(a - hi + 2 * np.pi),
                                                  abs(a - lo + 2 * np.pi), abs(a - hi - 2 * np.pi))
                worst = max(worst, d)
        assert worst <= np.pi / B + 1e-9 < margin, (B, worst)


def test_sector_windows_fit_shared_memory_and_stay_inside_the_grid():
    from emu_bindings import dir_windows
    for name in ("sibal1", "Spielberg_map", "basement_fixed"):
        em = EmuMap(maps.load_named_map(name))
        cap = 112 * 1024
        for bx0, by0 in ((-500, -500), (0, 0), (em.PW // 2, em.PH // 2), (em.PW - 30, em.PH - 30), (em.PW + 400, 7)):
            box, w = dir_windows(em, bx0, by0, cap)
            assert box >= 64 and box % 16 == 0
            assert (w[:, 2] * w[:, 3] <= cap).all()
            assert (w[:, 0] % 16 == 0).all() and (w[:, 2] % 16 == 0).all()
            assert (w[:, 0] >= 0).all() and (w[:, 1] >= 0).all()
            assert (w[:, 0] + w[:, 2] <= em.PW).all() and (w[:, 1] + w[:, 3] <= em.PH).all()


@pytest.mark.parametrize("fixture,name", [("update_sibal1_4000.npz", "sibal1"), ("update_basement_fixed_1000.npz", "basement_fixed"),
                                          ("update_Spielberg_map_2000.npz", "Spielberg_map")])
def test_directional_march_reproduces_the_golden_range_steps(fixture, name):
    """The committed golden fixtures hold the step indices the UNMODIFIED reference produced for
    its own proposal particles; the directional march (CPU build of the kernel source, window
    path included) reproduces them exactly for every update of every fixture."""
    from helpers import load_golden
    z = load_golden(fixture)
    g = maps.load_named_map(name)
    em = EmuMap(g)
    for t in range(len(z["u"])):
        p = z["particles"][t]          # the proposal the reference cast its rays from (:540)
        for box in ((0, 128) if name == "sibal1" else (128,)):
            got, replays = em.range_steps_dir(p[0], p[1], p[2], z["angles"], buckets=2048, window_box=box)
            assert replays >= 0, "%d reads fell outside a sector window" % -replays
            assert np.array_equal(got, z["steps"][t]), "update %d box %d: %d rays differ" % (t, box, int((got != z["steps"][t]).sum()))


def test_coarse_key_layout_spreads_every_search_step_over_the_banks():
    """The coarse level of the resampling search is a binary search over a power-of-two table in shared memory.  Stored
    at their index, the candidates of the first steps are multiples of large powers of two -- all in bank 0 (ncu found
    216 wavefronts per warp and search, 88 % replays).  coarse_slot pads one word per 32 keys and one per 1024: the
    layout stays injective and monotone, the table grows by ~3 %, and the candidates of EVERY step fall into as many
    banks as a random access pattern would reach."""
    from emu_bindings import coarse_slot, coarse_slots
    for nc in (16384, 8192, 4096, 1000):
        slots = np.array([coarse_slot(k) for k in range(nc)])
        assert (np.diff(slots) >= 1).all() and slots[0] == 0                 # injective, monotone
        assert coarse_slots(nc) == slots[-1] + 1 <= nc + nc // 32 + nc // 1024 + 1
        if nc & (nc - 1):
            continue
        # the probes of a lower_bound over [0, nc): step s looks at the midpoints of the 2^s intervals
        lo, hi = np.zeros(1, dtype=np.int64), np.full(1, nc, dtype=np.int64)
        step = 0
        while len(lo) <= 4096 and (hi - lo).min() >= 1:
            mid = (lo + hi) // 2
            cand = np.unique(mid)
            per_bank = np.bincount(slots[cand] % 32, minlength=32)
            assert (per_bank > 0).sum() == min(len(cand), 32)                # as many banks as candidates (up to all 32)
            assert per_bank.max() == -(-len(cand) // 32)                     # and evenly filled
            if nc == 16384 and 1 <= step <= 9:
                assert len(np.unique(cand % 32)) == 1                        # unpadded: all of them in ONE bank
            lo, hi = np.concatenate([lo, mid + 1]), np.concatenate([mid, hi])
            step += 1
