"""The node's update shell against the reference's OWN timer_update / odomCB /
clicked_pose / get_current_pose, tick by tick on CPU.

The unmodified reference (oracle/_ref, clocks scripted through oracle/shim/ref_clock_prelude.hpp) and
monte_carlo_localization_b200/host/update_shell.hpp (through the pfhost_shell_* hooks of libpf_host.so)
receive the same scripted odometry / scan / clock sequence.  The MCL update itself is the reference's
(its pose and its TimingStats time are handed to the shell under test): what is compared is everything
AROUND the hot path -- the synthesised action (through its bit-exact effect on the particles), the
start-up jitter, the odometry-tracking state, the delay compensation including the 200-iteration
TimingStats window (:814-827), the re-anchoring and the get_current_pose priority chain (:892-916).
"""
import ctypes as C
import os

import numpy as np
import pytest

from monte_carlo_localization_b200 import maps, synth
from oracle import bindings as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SO = os.path.join(ROOT, "monte_carlo_localization_b200", "host", "libpf_host.so")
pytestmark = pytest.mark.skipif(not ob.have_reference(), reason="oracle/_ref (the compiled reference) is absent")

_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
_clock = {"t": 1000.0}   # the reference keeps ONE timer timeline per process (function-local statics, :735-746)


class Shell:
    """UpdateShell of the host mirror, driven through the C hooks."""

    def __init__(self, factor=1.5, max_pose_range=10000.0):
        L = self.L = C.CDLL(HOST_SO)
        L.pfhost_shell_create.restype = C.c_void_p
        L.pfhost_shell_create.argtypes = [C.c_double, C.c_double]
        L.pfhost_shell_destroy.argtypes = [C.c_void_p]
        L.pfhost_shell_odom.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int]
        L.pfhost_shell_clicked_pose.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.pfhost_shell_timer_update.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pfhost_shell_state.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.pfhost_shell_current_pose.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.pfhost_shell_set_inferred.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self.h = L.pfhost_shell_create(factor, max_pose_range)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pfhost_shell_destroy(self.h)
            self.h = None

    def odom(self, pose, v, w, map_init=True):
        self.L.pfhost_shell_odom(self.h, _dp(np.asarray(pose, dtype=np.float64)), v, w, int(map_init))

    def clicked_pose(self, pose):
        self.L.pfhost_shell_clicked_pose(self.h, _dp(np.asarray(pose, dtype=np.float64)))

    def timer_update(self, dt, mcl_pose, mcl_ms, normals, map_init=True, lidar_init=True, nranges=60, ok=True):
        act = np.zeros(3)
        z = np.asarray(normals, dtype=np.float64)
        ran = self.L.pfhost_shell_timer_update(self.h, dt, int(map_init), int(lidar_init), nranges,
                                               _dp(np.asarray(mcl_pose, dtype=np.float64)), mcl_ms, int(ok), _dp(z), _dp(act))
        return bool(ran), act

    def state(self):
        out = np.empty(23)
        self.L.pfhost_shell_state(self.h, _dp(out))
        return ob.unpack_shell_state(out)

    def current_pose(self, mean=None, map_init=True):
        out = np.empty(3)
        m = np.zeros(3) if mean is None else np.asarray(mean, dtype=np.float64)
        self.L.pfhost_shell_current_pose(self.h, int(map_init), _dp(m), int(mean is not None), _dp(out))
        return out


def _assert_state_equal(a, b, tick):
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert np.array_equal(a[k], b[k]), "tick %d: %s differs: %s vs %s" % (tick, k, a[k], b[k])
        elif isinstance(a[k], float):
            assert a[k] == pytest.approx(b[k], rel=1e-12, abs=1e-12), "tick %d: %s differs: %r vs %r" % (tick, k, a[k], b[k])
        else:
            assert a[k] == b[k], "tick %d: %s differs: %r vs %r" % (tick, k, a[k], b[k])


def _setup(seed, n=128, **params):
    g = maps.load_named_map("sibal1")
    # no motion noise: the particles after a tick then show the synthesised action bit for bit
    ref = ob.Reference(g, seed, max_particles=n, motion_dispersion_x=0.0, motion_dispersion_y=0.0,
                       motion_dispersion_theta=0.0, num_threads=2, **params)
    ref.seed(seed)
    angles_full = synth.laser_angles()
    orc = ob.Oracle(g, synth.downsample(angles_full), max_particles=n, motion_dispersion_x=0.0,
                    motion_dispersion_y=0.0, motion_dispersion_theta=0.0)
    scan = synth.scan_from_pose(orc.calc_range_many, [-3.3, 1.6, 0.3], angles_full, None)
    return g, ref, orc, scan


def _tick(ref, shell, orc, twin, dt, n, tick, pose0):
    """One timer tick on both sides; returns the reference's state after it."""
    # all particles on one pose with uniform weights: MCL's proposal is motion_model(pose0, action) for every particle
    ref.set_state(np.repeat(np.asarray(pose0, dtype=np.float64)[:, None], n, axis=1), np.full(n, 1.0 / n))
    before = ref.shell_state()
    jitter = (not before["odom_initialized"]) and (not before["pose_initialized_from_rviz"]) and before["iters"] + 1 < 15
    z3 = twin.normal(3) if jitter else np.zeros(3)       # the draws timer_update takes from rng_ (:769-771) ...
    twin.update_noise(n)                                   # ... before MCL's own draws
    _clock["t"] += dt
    ob.Reference.clock_set_steady(_clock["t"])
    ref.timer_update()
    after = ref.shell_state()
    # MCL time of this tick as TimingStats saw it; on a tick that reset the window (:826) it is the same
    # scripted figure as on the tick before (the high-resolution clock advances by a fixed quantum per read)
    if after["window_count"] == before["window_count"] + 1:
        mcl_ms = after["window_total_ms"] - before["window_total_ms"]
        _tick.last_ms = mcl_ms
    else:
        mcl_ms = _tick.last_ms
    ran, action = shell.timer_update(dt, after["inferred"], mcl_ms, z3)
    assert ran == (after["iters"] == before["iters"] + 1), "tick %d: one side skipped the tick" % tick
    if ran:
        # the synthesised action, through its effect: every particle of the reference moved by exactly this action
        want = orc.motion_model(np.repeat(np.asarray(pose0, dtype=np.float64)[:, None], n, axis=1), action, np.zeros(3 * n))
        got, _ = ref.get_state()
        assert np.array_equal(got, want), "tick %d: action %s does not reproduce the reference's motion" % (tick, action)
    _assert_state_equal(shell.state(), after, tick)
    return after


def test_timer_update_shell_matches_reference_tick_by_tick():
    n = 128
    g, ref, orc, scan = _setup(4242, n)
    twin = ob.NoiseStream(4242)
    shell = Shell()
    ob.Reference.clock_fake(True)
    try:
        ob.Reference.clock_set_hr_quantum(0.25)     # every clock read inside MCL() advances TimingStats' clock by 0.25 ms
        _clock["t"] += 0.5
        ob.Reference.clock_set_steady(_clock["t"])
        ref.timer_update()                          # the reference's first call only initialises its timer (:742-747)
        pose0 = [-3.3, 1.6, 0.3]
        # no scan yet: both sides skip (:758)
        _clock["t"] += 0.01
        ob.Reference.clock_set_steady(_clock["t"])
        ref.timer_update()
        ran, _ = shell.timer_update(0.01, [0, 0, 0], 0.0, np.zeros(3), lidar_init=False, nranges=0)
        assert not ran and ref.shell_state()["iters"] == 0
        ref.lidar(float(synth.ANGLE_MIN), float(synth.ANGLE_INC), scan)
        # (1) start-up without odometry: decaying jitter for the first ticks (:767-772), no tracking
        tick = 0
        for _ in range(17):
            tick += 1
            st = _tick(ref, shell, orc, twin, 0.01, n, tick, pose0)
            assert not st["odom_tracking_active"]
            assert np.array_equal(shell.current_pose(), ref.current_pose())       # priority 2: the filter estimate
        # (2) odometry arrives: action = [v dt, 0, w dt] (:764-766), tracking initialised from the estimate (:785-788),
        #     delay compensation from the TimingStats window (:791-802), re-anchoring every tick (:804-806)
        odom = np.array([100.0, -50.0, 0.3])
        for k in range(230):                        # crosses iteration 200: TimingStats window reset (:814-827)
            tick += 1
            v, w = (2.0 + 0.01 * k, 0.3 * np.sin(0.1 * k)) if k % 37 else (0.0, 0.0)   # incl. standstill ticks: zero action
            if k == 120:
                ob.Reference.clock_set_hr_quantum(0.9)   # MCL gets slower: the windowed mean must follow
            odom = odom + np.array([v * 0.01 * np.cos(odom[2]), v * 0.01 * np.sin(odom[2]), w * 0.01])
            ref.odom(odom, v, w)
            # the node receives the heading as a quaternion (nav_msgs/Odometry): hand the shell the pose as the
            # reference decoded it, so that the comparison is about the shell and not about the ROS transport
            shell.odom(ref.shell_state()["last_pose"], v, w)
            _assert_state_equal(shell.state(), ref.shell_state(), tick)
            dt = 0.01 if k % 50 else 0.00005        # a tick below the 0.1 ms threshold applies no motion (:754)
            st = _tick(ref, shell, orc, twin, dt, n, tick, pose0)
            assert st["odom_tracking_active"]
            assert np.array_equal(shell.current_pose(), ref.current_pose())       # priority 1: odometry tracking
        assert ref.shell_state()["iters"] > 200 and ref.shell_state()["window_count"] < 100   # the window was reset
        # (3) /initialpose: re-initialise, tracking from the clicked pose (:355-374)
        ref.clicked_pose([-3.0, 1.5, 0.1])
        shell.clicked_pose(ref.shell_state()["inferred"])     # (decoded from the message's quaternion)
        _assert_state_equal(shell.state(), ref.shell_state(), tick)
        assert np.array_equal(shell.current_pose(), ref.current_pose())
        for k in range(5):
            tick += 1
            ref.odom(odom + 0.01 * k, 1.0, 0.0)
            shell.odom(ref.shell_state()["last_pose"], 1.0, 0.0)
            _tick(ref, shell, orc, twin, 0.02, n, tick, [-3.0, 1.5, 0.1])
            assert np.array_equal(shell.current_pose(), ref.current_pose())
    finally:
        ob.Reference.clock_fake(False)


def test_get_current_pose_priority_chain_matches_reference():
    """Priorities 2-4 of get_current_pose (:899-916): estimate, particle mean, last odometry pose, origin."""
    n = 64
    g, ref, orc, scan = _setup(7, n, max_pose_range=50.0)
    shell = Shell(max_pose_range=50.0)
    p, _ = ref.get_state()
    mean = p.mean(axis=1)                 # particles_.colwise().mean(): a plain sum / N in column order
    for inferred in ([1.0, 2.0, 0.5], [float("nan"), 0.0, 0.0], [60.0, 0.0, 0.0]):
        ref.set_inferred(inferred)
        shell.L.pfhost_shell_set_inferred(shell.h, _dp(np.asarray(inferred, dtype=np.float64)))
        want = ref.current_pose()
        got = shell.current_pose(mean=mean)
        assert np.allclose(got, want, rtol=0, atol=1e-12), (inferred, got, want)
    # invalid estimate AND invalid particle mean: the last odometry pose (priority 4), else the origin
    far = np.full((3, n), 1e6)
    ref.set_state(far, None)
    ref.set_inferred([float("inf"), 0.0, 0.0])
    shell.L.pfhost_shell_set_inferred(shell.h, _dp(np.asarray([float("inf"), 0.0, 0.0])))
    assert np.array_equal(shell.current_pose(mean=far.mean(axis=1)), ref.current_pose())       # origin
    ref.odom([3.0, 4.0, 0.2], 0.0, 0.0)
    shell.odom(ref.shell_state()["last_pose"], 0.0, 0.0)
    assert np.array_equal(shell.current_pose(mean=far.mean(axis=1)), ref.current_pose())       # last_pose_
    assert np.allclose(ref.current_pose(), [3.0, 4.0, 0.2], atol=1e-12)


def test_failed_update_leaves_tracking_state_untouched():
    """when the update cannot run (e.g. a scan of another length) the tick must not re-anchor the
    odometry tracking from a stale pose."""
    shell = Shell()
    shell.odom([1.0, 2.0, 0.1], 1.0, 0.0)
    ran, _ = shell.timer_update(0.01, [0.5, 0.5, 0.0], 1.0, np.zeros(3))
    assert ran
    before = shell.state()
    ran, _ = shell.timer_update(0.01, [9.0, 9.0, 9.0], 1.0, np.zeros(3), ok=False)
    after = shell.state()
    assert not ran
    for k in ("inferred", "odom_pose", "odom_reference_pose", "odom_reference_odom", "window_count", "window_total_ms"):
        assert np.array_equal(np.asarray(before[k]), np.asarray(after[k])), k
