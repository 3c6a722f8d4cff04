"""Shared helpers for the parity tests (TEST INFRASTRUCTURE)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerances (BASELINE.json): indices bit-exact, ranges within 1 cell (we expect
# equality), normalised weights 1e-5 relative, pose 1 mm / 1e-4 rad.
WEIGHT_RTOL = 1e-5
POSE_XY_TOL = 1e-3
POSE_TH_TOL = 1e-4


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def steps_from_ranges(r, res, M):
    """The reference's range -> table index conversion (src/particle_filter.cpp:556-574)."""
    r = np.asarray(r, dtype=np.float32)
    px = (r.astype(np.float64) / res).astype(np.float32)
    px = np.minimum(px, np.float32(M))
    return np.clip(np.round(px).astype(np.int64), 0, M)


def ang_diff(a, b):
    return np.abs((np.asarray(a) - np.asarray(b) + np.pi) % (2 * np.pi) - np.pi)


def assert_pose_close(p, q):
    p, q = np.asarray(p), np.asarray(q)
    assert abs(p[0] - q[0]) <= POSE_XY_TOL and abs(p[1] - q[1]) <= POSE_XY_TOL, (p, q)
    assert ang_diff(p[2], q[2]) <= POSE_TH_TOL, (p, q)


def assert_weights_close(w, w_ref):
    w, w_ref = np.asarray(w), np.asarray(w_ref)
    rel = np.abs(w - w_ref) / np.maximum(np.abs(w_ref), 1e-300)
    assert rel.max() <= WEIGHT_RTOL, "max relative weight error %.3e" % rel.max()
