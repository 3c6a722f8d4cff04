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


class EmuRanks:
    """`world` ranks of ONE particle-sharded filter emulated on one GPU (test infrastructure).

    Kernels of different ranks must never spin on each other on one device, so the contexts run the
    host-ordered exchange (mcl_shard_set_exchange(fused=0, hook)): every rank runs in its own host
    thread, and the library calls the hook -- a thread barrier -- wherever all ranks must have
    published before any rank consumes.  Everything else is the product path: the same kernels, the
    same mailboxes, the same routed pushes (into the other context's memory on the same device).
    """

    def __init__(self, grid, angles, n_global, world, keep_ranges=True, two_hop=None, **params):
        import threading

        from monte_carlo_localization_b200 import MclContext
        self.world = world
        self.n_local = n_global // world
        self.ctxs = [MclContext(device=0, shard=(world, r), max_particles=n_global, **params) for r in range(world)]
        for c in self.ctxs:
            c.set_map(grid)
            c.set_beam_angles(angles)
            c.set_keep_ranges(keep_ranges)
        for c in self.ctxs:
            c.shard_connect_local(self.ctxs)
        self._barrier = threading.Barrier(world)
        for c in self.ctxs:
            c.shard_set_exchange(False, hook=self._hook)
            if two_hop is not None:      # None: the library's default (two-hop request routing)
                c.shard_set_route(two_hop)

    def _hook(self):
        self._barrier.wait(timeout=120)
        return 0

    def run(self, fn):
        """fn(rank, ctx) on every rank concurrently; returns the results in rank order."""
        import threading
        out, err = [None] * self.world, [None] * self.world

        def work(r):
            try:
                out[r] = fn(r, self.ctxs[r])
            except BaseException as e:   # noqa: BLE001 -- re-raised below
                err[r] = e
                self._barrier.abort()

        th = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out

    def sl(self, r):
        return slice(r * self.n_local, (r + 1) * self.n_local)

    def set_state(self, particles, weights):
        for r, c in enumerate(self.ctxs):
            c.set_particles(None if particles is None else np.ascontiguousarray(particles[:, self.sl(r)]),
                            None if weights is None else np.ascontiguousarray(weights[self.sl(r)]))

    def update(self, action, obs, u=None, z=None):
        """One update on all ranks; returns the poses (one per rank; all must be equal)."""
        return self.run(lambda r, c: c.update(action, obs, u, z))

    def gather(self, what):
        """Whole-filter array from the ranks' slices: what(ctx) -> array whose LAST axis is the particle."""
        return np.concatenate([what(c) for c in self.ctxs], axis=-1)

    def close(self):
        for c in self.ctxs:
            c.close()
