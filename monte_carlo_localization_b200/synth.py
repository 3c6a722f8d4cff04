"""Synthetic inputs for the MCL update: scan geometry, trajectories, scans, odometry.

SURVEY 8(d) common inputs: 1080 beams, angle_min -2.35, increment 4.7/1079 in float32
arithmetic like lidarCB (src/particle_filter.cpp:303), angle_step 18 -> 60 beams; scan =
ray cast from the ground-truth pose (12 m cap) + N(0, 0.01 m); action = [v dt, 0, w dt]
(:764-766), dt = 0.025 s.  The ray caster is passed in (the product's
``MclContext.calc_range_many`` in bench.py, the oracle's in CPU tests).
"""
from __future__ import annotations

import numpy as np

NUM_BEAMS_FULL = 1080
ANGLE_MIN = np.float32(-2.35)
ANGLE_INC = np.float32(4.7 / 1079)
DT = 0.025


def laser_angles(n: int = NUM_BEAMS_FULL, angle_min=ANGLE_MIN, angle_inc=ANGLE_INC) -> np.ndarray:
    """laser_angles_[i] = angle_min + i * angle_increment, evaluated in float32 (:303)."""
    i = np.arange(n).astype(np.float32)
    return (np.float32(angle_min) + i * np.float32(angle_inc)).astype(np.float32)


def downsample(a: np.ndarray, angle_step: int = 18) -> np.ndarray:
    """every ANGLE_STEP-th beam (:307-310, :317-320)."""
    return np.ascontiguousarray(a[::angle_step])


def beam_angles(angle_step: int = 18) -> np.ndarray:
    return downsample(laser_angles(), angle_step)


def free_component_mask(grid, min_cells: int = 2000, pick: str = "smallest_big"):
    """Mask of the drivable free component: the smallest connected free region above
    min_cells (on race maps the track ring is smaller than infield/exterior)."""
    from scipy import ndimage
    lab, n = ndimage.label(grid.data == 0)
    if n == 0:
        raise ValueError("map has no free space")
    sizes = ndimage.sum(grid.data == 0, lab, range(1, n + 1))
    order = np.argsort(sizes)
    big = [i + 1 for i in order if sizes[i] >= min_cells]
    if not big:
        big = [int(order[-1]) + 1]
    chosen = big[0] if pick == "smallest_big" else big[-1]
    return lab == chosen


def centreline_loop(grid, n_points: int = 400, min_clearance_cells: float = 4.0, mask=None):
    """A closed ground-truth path along the ridge of maximum clearance of the drivable
    region, ordered by angle around the region's centroid and smoothed.  Returns
    (x, y, theta) arrays of length n_points."""
    from scipy import ndimage
    if mask is None:
        mask = free_component_mask(grid)
    dist = ndimage.distance_transform_edt(mask)
    rows, cols = np.nonzero(mask)
    cy, cx = rows.mean(), cols.mean()
    ang = np.arctan2(rows - cy, cols - cx)
    bins = np.linspace(-np.pi, np.pi, n_points + 1)
    which = np.digitize(ang, bins) - 1
    px = np.full(n_points, np.nan)
    py = np.full(n_points, np.nan)
    d = dist[rows, cols]
    for b in range(n_points):
        sel = np.nonzero(which == b)[0]
        if sel.size == 0:
            continue
        k = sel[np.argmax(d[sel])]
        if d[k] >= min_clearance_cells:
            px[b], py[b] = cols[k] + 0.5, rows[k] + 0.5
    ok = ~np.isnan(px)
    if ok.sum() < 8:
        raise ValueError("could not trace a centreline on this map")
    idx = np.arange(n_points)
    px = np.interp(idx, idx[ok], px[ok], period=n_points)
    py = np.interp(idx, idx[ok], py[ok], period=n_points)
    for _ in range(3):   # circular smoothing
        px = (np.roll(px, 1) + px + np.roll(px, -1)) / 3.0
        py = (np.roll(py, 1) + py + np.roll(py, -1)) / 3.0
    res = grid.resolution_f64
    x = px * res + grid.origin[0]
    y = py * res + grid.origin[1]
    th = np.arctan2(np.roll(y, -1) - y, np.roll(x, -1) - x)
    return x, y, th


def trajectory(grid, n_steps: int, speed: float, dt: float = DT, mask=None):
    """Ground-truth poses gt[n_steps+1, 3] driven along the centreline at `speed`, and the
    odometry actions [n_steps, 3] = [v dt, 0, w dt] the node would synthesise (:761-766)."""
    x, y, th = centreline_loop(grid, mask=mask)
    seg = np.hypot(np.roll(x, -1) - x, np.roll(y, -1) - y)
    s = np.concatenate([[0.0], np.cumsum(seg)])
    total = s[-1]
    xs = np.concatenate([x, x[:1]])
    ys = np.concatenate([y, y[:1]])
    t = (np.arange(n_steps + 1) * speed * dt) % total
    gx = np.interp(t, s, xs)
    gy = np.interp(t, s, ys)
    ahead = (t + 0.25) % total
    gth = np.arctan2(np.interp(ahead, s, ys) - gy, np.interp(ahead, s, xs) - gx)
    gt = np.stack([gx, gy, gth], axis=1)
    dth = np.diff(gth)
    dth = (dth + np.pi) % (2 * np.pi) - np.pi
    dist = np.hypot(np.diff(gx), np.diff(gy))
    actions = np.stack([dist, np.zeros(n_steps), dth], axis=1)
    return gt, actions


def scan_from_pose(cast_many, pose, angles_full: np.ndarray, rng: np.random.Generator | None,
                   sigma: float = 0.01, max_range: float = 12.0) -> np.ndarray:
    """Full scan (float32 metres) at `pose`: ray cast + N(0, sigma), clipped to [0, max_range]."""
    n = len(angles_full)
    q = np.empty((3, n), dtype=np.float64)
    q[0] = pose[0]
    q[1] = pose[1]
    q[2] = pose[2] + angles_full.astype(np.float64)
    r = np.asarray(cast_many(q), dtype=np.float64)
    if rng is not None and sigma > 0:
        r = r + rng.normal(0.0, sigma, n)
    return np.clip(r, 0.0, max_range).astype(np.float32)
