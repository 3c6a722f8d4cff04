"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE ONLY).

``Oracle``     -> oracle/liboracle.so      (CPU restatement, oracle/mcl_oracle.cpp)
``Reference``  -> oracle/_ref/libref_pf.so (the unmodified reference sources behind shims)
``NoiseStream``-> the libstdc++ RNG twin that produces the injected noise arrays

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_pf.so")

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int8_p = C.POINTER(C.c_int8)


def build(force: bool = False) -> None:
    """Compile the checkers (oracle always; oracle/_ref only where /root/reference exists)."""
    if force or not os.path.exists(_ORACLE_SO) or (
            os.path.getmtime(_ORACLE_SO) < os.path.getmtime(os.path.join(_HERE, "mcl_oracle.cpp"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liboracle.so")])
    if os.path.exists("/root/reference/src/particle_filter.cpp"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


class OrcParams(C.Structure):
    _fields_ = [("max_particles", C.c_int), ("num_threads", C.c_int),
                ("use_parallel_raycasting", C.c_int),
                ("squash_factor", C.c_double), ("max_range", C.c_double),
                ("z_short", C.c_double), ("z_max", C.c_double), ("z_rand", C.c_double),
                ("z_hit", C.c_double), ("sigma_hit", C.c_double),
                ("motion_dispersion_x", C.c_double), ("motion_dispersion_y", C.c_double),
                ("motion_dispersion_theta", C.c_double)]


class OrcTiming(C.Structure):
    _fields_ = [("total_ms", C.c_double), ("resample_ms", C.c_double), ("motion_ms", C.c_double),
                ("query_ms", C.c_double), ("raycast_ms", C.c_double), ("sensor_ms", C.c_double),
                ("pose_ms", C.c_double), ("count", C.c_int)]


_lib = None


def _oracle():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_ORACLE_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_map.argtypes = [C.c_void_p, c_int8_p, C.c_int, C.c_int, C.c_float, C.c_double,
                                  C.c_double, C.c_double]
        L.orc_max_range_px.argtypes = [C.c_void_p]
        L.orc_get_sensor_table.argtypes = [C.c_void_p, c_double_p]
        L.orc_set_beam_angles.argtypes = [C.c_void_p, c_float_p, C.c_int]
        L.orc_set_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_get_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_init_pose.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.orc_init_global.argtypes = [C.c_void_p, c_int32_p, c_double_p]
        L.orc_num_free_cells.argtypes = [C.c_void_p]
        L.orc_cast_ray.restype = C.c_float
        L.orc_cast_ray.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.orc_calc_range_many.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_float_p]
        L.orc_update.argtypes = [C.c_void_p, c_double_p, c_float_p, C.c_int, c_double_p, c_double_p,
                                 c_int32_p]
        L.orc_expected_pose.argtypes = [C.c_void_p, c_double_p]
        L.orc_get_ranges.argtypes = [C.c_void_p, c_float_p]
        L.orc_get_raw_weights.argtypes = [C.c_void_p, c_double_p]
        L.orc_mean_cells_per_ray.restype = C.c_double
        L.orc_mean_cells_per_ray.argtypes = [C.c_void_p]
        L.orc_get_timing.argtypes = [C.c_void_p, C.POINTER(OrcTiming)]
        L.orc_reset_timing.argtypes = [C.c_void_p]
        L.orc_motion_model.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
        L.orc_sensor_weights.argtypes = [C.c_void_p, c_double_p, c_float_p, C.c_int, c_double_p]
        L.orc_resample_indices.argtypes = [c_double_p, C.c_int, c_double_p, C.c_int, c_int32_p,
                                           c_double_p]
        L.orc_normalize_angle.restype = C.c_double
        L.orc_normalize_angle.argtypes = [C.c_double]
        L.orc_rng_create.restype = C.c_void_p
        L.orc_rng_create.argtypes = [C.c_uint32]
        L.orc_rng_destroy.argtypes = [C.c_void_p]
        L.orc_rng_canonical.argtypes = [C.c_void_p, C.c_int64, c_double_p]
        L.orc_rng_normal.argtypes = [C.c_void_p, C.c_int64, c_double_p]
        L.orc_rng_uniform_int.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, c_int32_p]
        L.orc_rng_uniform_real.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, c_double_p]
        L.orc_rng_global_init.argtypes = [C.c_void_p, C.c_int64, C.c_int32, c_int32_p, c_double_p]
        L.orc_rng_raw.restype = C.c_uint32
        L.orc_rng_raw.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def default_params(**kw) -> OrcParams:
    p = OrcParams()
    _oracle().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class NoiseStream:
    """Twin of the reference's ``rng_``/``normal_dist_`` (particle_filter.hpp:165-167)."""

    def __init__(self, seed: int):
        self._L = _oracle()
        self._h = self._L.orc_rng_create(C.c_uint32(seed))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_rng_destroy(self._h)
            self._h = None

    def canonical(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        self._L.orc_rng_canonical(self._h, n, _dp(out))
        return out

    def normal(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        self._L.orc_rng_normal(self._h, n, _dp(out))
        return out

    def global_init(self, n: int, n_free: int):
        cell = np.empty(n, dtype=np.int32)
        th = np.empty(n, dtype=np.float64)
        self._L.orc_rng_global_init(self._h, n, n_free, _ip(cell), _dp(th))
        return cell, th

    def update_noise(self, n: int):
        """Noise one MCL() consumes, in the reference's draw order: N discrete draws
        (:661-665) then 3N normals (:496-498)."""
        u = self.canonical(n)
        z = self.normal(3 * n)
        return u, z


class Oracle:
    """CPU restatement of the reference ParticleFilter update."""

    def __init__(self, grid, angles, **params):
        self._L = _oracle()
        self.params = default_params(**params)
        self.N = self.params.max_particles
        self._h = self._L.orc_create(C.byref(self.params))
        d = np.ascontiguousarray(grid.data, dtype=np.int8)
        rc = self._L.orc_set_map(self._h, d.ctypes.data_as(c_int8_p), grid.width, grid.height,
                                 C.c_float(float(grid.resolution)), grid.origin[0], grid.origin[1],
                                 grid.origin[2])
        if rc != 0:
            raise RuntimeError("orc_set_map failed %d" % rc)
        self.M = self._L.orc_max_range_px(self._h)
        self.angles = np.ascontiguousarray(angles, dtype=np.float32)
        self.R = len(self.angles)
        self._L.orc_set_beam_angles(self._h, _fp(self.angles), self.R)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    def sensor_table(self) -> np.ndarray:
        t = np.empty((self.M + 1) * (self.M + 1), dtype=np.float64)
        self._L.orc_get_sensor_table(self._h, _dp(t))
        return t

    def set_state(self, particles_colmajor, weights=None):
        p = np.ascontiguousarray(particles_colmajor, dtype=np.float64).reshape(-1)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._L.orc_set_state(self._h, _dp(p), None if w is None else _dp(w))

    def get_state(self):
        p = np.empty(3 * self.N, dtype=np.float64)
        w = np.empty(self.N, dtype=np.float64)
        self._L.orc_get_state(self._h, _dp(p), _dp(w))
        return p.reshape(3, self.N), w

    def init_pose(self, pose, z3n):
        pose = np.asarray(pose, dtype=np.float64)
        z = np.ascontiguousarray(z3n, dtype=np.float64)
        self._L.orc_init_pose(self._h, _dp(pose), _dp(z))

    def init_global(self, cell, theta):
        cell = np.ascontiguousarray(cell, dtype=np.int32)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self._L.orc_init_global(self._h, _ip(cell), _dp(theta))
        if rc != 0:
            raise RuntimeError("orc_init_global failed %d" % rc)

    def num_free_cells(self) -> int:
        return self._L.orc_num_free_cells(self._h)

    def cast_ray(self, x, y, a) -> float:
        return float(self._L.orc_cast_ray(self._h, x, y, a))

    def calc_range_many(self, queries_colmajor) -> np.ndarray:
        q = np.ascontiguousarray(queries_colmajor, dtype=np.float64).reshape(-1)
        n = q.size // 3
        out = np.empty(n, dtype=np.float32)
        self._L.orc_calc_range_many(self._h, _dp(q), n, _fp(out))
        return out

    def update(self, action, obs, u, z3n):
        action = np.asarray(action, dtype=np.float64)
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        u = np.ascontiguousarray(u, dtype=np.float64)
        z = np.ascontiguousarray(z3n, dtype=np.float64)
        idx = np.empty(self.N, dtype=np.int32)
        self._L.orc_update(self._h, _dp(action), _fp(obs), len(obs), _dp(u), _dp(z), _ip(idx))
        return idx

    def expected_pose(self) -> np.ndarray:
        out = np.empty(3, dtype=np.float64)
        self._L.orc_expected_pose(self._h, _dp(out))
        return out

    def ranges(self) -> np.ndarray:
        out = np.empty(self.N * self.R, dtype=np.float32)
        self._L.orc_get_ranges(self._h, _fp(out))
        return out.reshape(self.N, self.R)

    def raw_weights(self) -> np.ndarray:
        out = np.empty(self.N, dtype=np.float64)
        self._L.orc_get_raw_weights(self._h, _dp(out))
        return out

    def mean_cells_per_ray(self) -> float:
        return float(self._L.orc_mean_cells_per_ray(self._h))

    def timing(self) -> dict:
        t = OrcTiming()
        self._L.orc_get_timing(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in OrcTiming._fields_}

    def reset_timing(self):
        self._L.orc_reset_timing(self._h)

    def motion_model(self, particles_colmajor, action, z3n) -> np.ndarray:
        p = np.array(particles_colmajor, dtype=np.float64).reshape(-1).copy()
        a = np.asarray(action, dtype=np.float64)
        z = np.ascontiguousarray(z3n, dtype=np.float64)
        self._L.orc_motion_model(self._h, _dp(p), _dp(a), _dp(z))
        return p.reshape(3, self.N)

    def sensor_weights(self, particles_colmajor, obs) -> np.ndarray:
        p = np.ascontiguousarray(particles_colmajor, dtype=np.float64).reshape(-1)
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        out = np.empty(self.N, dtype=np.float64)
        self._L.orc_sensor_weights(self._h, _dp(p), _fp(obs), len(obs), _dp(out))
        return out


def resample_indices(weights, u, want_cdf=False):
    L = _oracle()
    w = np.ascontiguousarray(weights, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    idx = np.empty(len(u), dtype=np.int32)
    cdf = np.empty(len(w), dtype=np.float64) if want_cdf else None
    L.orc_resample_indices(_dp(w), len(w), _dp(u), len(u), _ip(idx), None if cdf is None else _dp(cdf))
    return (idx, cdf) if want_cdf else idx


def normalize_angle(a: float) -> float:
    return float(_oracle().orc_normalize_angle(a))


# ------------------------------------------------------------------ reference (tier A)
_ref = None


def _reflib():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            build()
        if not os.path.exists(_REF_SO):
            raise FileNotFoundError("oracle/_ref/libref_pf.so not built (needs /root/reference)")
        L = C.CDLL(_REF_SO)
        L.ref_set_param_int.argtypes = [C.c_char_p, C.c_longlong]
        L.ref_set_param_double.argtypes = [C.c_char_p, C.c_double]
        L.ref_set_param_bool.argtypes = [C.c_char_p, C.c_int]
        L.ref_set_verbose.argtypes = [C.c_int]
        L.ref_install_map.argtypes = [c_int8_p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double,
                                      C.c_double]
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_uint]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_seed.argtypes = [C.c_void_p, C.c_uint]
        for fn in ("ref_num_particles", "ref_max_range_px", "ref_num_threads"):
            getattr(L, fn).argtypes = [C.c_void_p]
        L.ref_map_resolution.restype = C.c_double
        L.ref_map_resolution.argtypes = [C.c_void_p]
        L.ref_lidar.argtypes = [C.c_void_p, C.c_float, C.c_float, c_float_p, C.c_int]
        L.ref_get_beam_angles.argtypes = [C.c_void_p, c_float_p]
        L.ref_get_downsampled_ranges.argtypes = [C.c_void_p, c_float_p]
        L.ref_get_sensor_table.argtypes = [C.c_void_p, c_double_p]
        L.ref_set_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.ref_get_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
        L.ref_init_pose.argtypes = [C.c_void_p, c_double_p]
        L.ref_init_global.argtypes = [C.c_void_p]
        L.ref_cast_ray.restype = C.c_float
        L.ref_cast_ray.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.ref_mcl.argtypes = [C.c_void_p, c_double_p, c_float_p, C.c_int, c_double_p]
        L.ref_get_ranges.argtypes = [C.c_void_p, c_float_p]
        L.ref_num_ranges.restype = C.c_longlong
        L.ref_num_ranges.argtypes = [C.c_void_p]
        L.ref_get_timing.argtypes = [C.c_void_p, c_double_p, C.POINTER(C.c_int)]
        L.ref_reset_timing.argtypes = [C.c_void_p]
        L.ref_clock_fake.argtypes = [C.c_int]
        L.ref_clock_set_steady.argtypes = [C.c_double]
        L.ref_clock_set_hr_quantum.argtypes = [C.c_double]
        L.ref_odom.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.ref_clicked_pose.argtypes = [C.c_void_p] + [C.c_double] * 3
        L.ref_timer_update.argtypes = [C.c_void_p]
        L.ref_current_pose.argtypes = [C.c_void_p, c_double_p]
        L.ref_set_inferred.argtypes = [C.c_void_p, c_double_p]
        L.ref_shell_state.argtypes = [C.c_void_p, c_double_p]
        L.ref_viz_sample.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        _ref = L
    return _ref


def unpack_shell_state(v) -> dict:
    """Layout shared by ref_shell_state (oracle/ref_harness.cpp) and pfhost_shell_state (host/host_capi.cpp)."""
    v = np.asarray(v, dtype=np.float64)
    d = {k: v[3 * i:3 * i + 3].copy() for i, k in enumerate(Reference.SHELL_KEYS)}
    d.update(iters=int(v[15]), odom_initialized=bool(v[16]), pose_initialized_from_rviz=bool(v[17]),
             odom_tracking_active=bool(v[18]), window_total_ms=float(v[19]), window_count=int(v[20]),
             velocity=float(v[21]), angular_velocity=float(v[22]))
    return d


class Reference:
    """The unmodified ``particle_filter_cpp::ParticleFilter`` (compiled against shims)."""

    def __init__(self, grid, seed: int, **params):
        L = self._L = _reflib()
        L.ref_clear_params()
        for k, v in params.items():
            if isinstance(v, bool):
                L.ref_set_param_bool(k.encode(), int(v))
            elif isinstance(v, int):
                L.ref_set_param_int(k.encode(), v)
            else:
                L.ref_set_param_double(k.encode(), float(v))
        d = np.ascontiguousarray(grid.data, dtype=np.int8)
        L.ref_install_map(d.ctypes.data_as(c_int8_p), grid.width, grid.height,
                          C.c_float(float(grid.resolution)), grid.origin[0], grid.origin[1],
                          grid.origin[2])
        self._h = L.ref_create(C.c_uint(seed))
        self.N = L.ref_num_particles(self._h)
        self.M = L.ref_max_range_px(self._h)
        self.R = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_destroy(self._h)
            self._h = None

    def seed(self, s: int):
        self._L.ref_seed(self._h, C.c_uint(s))

    def num_threads(self) -> int:
        return self._L.ref_num_threads(self._h)

    def map_resolution(self) -> float:
        return float(self._L.ref_map_resolution(self._h))

    def lidar(self, angle_min, angle_increment, ranges) -> int:
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        self.R = self._L.ref_lidar(self._h, C.c_float(angle_min), C.c_float(angle_increment), _fp(r), len(r))
        return self.R

    def beam_angles(self) -> np.ndarray:
        out = np.empty(self.R, dtype=np.float32)
        self._L.ref_get_beam_angles(self._h, _fp(out))
        return out

    def downsampled_ranges(self) -> np.ndarray:
        out = np.empty(self.R, dtype=np.float32)
        self._L.ref_get_downsampled_ranges(self._h, _fp(out))
        return out

    def sensor_table(self) -> np.ndarray:
        t = np.empty((self.M + 1) * (self.M + 1), dtype=np.float64)
        self._L.ref_get_sensor_table(self._h, _dp(t))
        return t

    def set_state(self, particles_colmajor, weights=None):
        p = np.ascontiguousarray(particles_colmajor, dtype=np.float64).reshape(-1)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._L.ref_set_state(self._h, _dp(p), None if w is None else _dp(w))

    def get_state(self):
        p = np.empty(3 * self.N, dtype=np.float64)
        w = np.empty(self.N, dtype=np.float64)
        self._L.ref_get_state(self._h, _dp(p), _dp(w))
        return p.reshape(3, self.N), w

    def init_pose(self, pose):
        self._L.ref_init_pose(self._h, _dp(np.asarray(pose, dtype=np.float64)))

    def init_global(self):
        self._L.ref_init_global(self._h)

    def cast_ray(self, x, y, a) -> float:
        return float(self._L.ref_cast_ray(self._h, x, y, a))

    def mcl(self, action, obs) -> np.ndarray:
        a = np.asarray(action, dtype=np.float64)
        o = np.ascontiguousarray(obs, dtype=np.float32)
        pose = np.empty(3, dtype=np.float64)
        self._L.ref_mcl(self._h, _dp(a), _fp(o), len(o), _dp(pose))
        return pose

    def ranges(self) -> np.ndarray:
        n = self._L.ref_num_ranges(self._h)
        out = np.empty(n, dtype=np.float32)
        self._L.ref_get_ranges(self._h, _fp(out))
        return out

    # ---- the node's update shell (timer_update / odomCB / clicked_pose / get_current_pose) ----------
    SHELL_KEYS = ("inferred", "odom_pose", "odom_reference_pose", "odom_reference_odom", "last_pose")

    @staticmethod
    def clock_fake(on: bool):
        _reflib().ref_clock_fake(int(on))

    @staticmethod
    def clock_set_steady(seconds: float):
        _reflib().ref_clock_set_steady(float(seconds))

    @staticmethod
    def clock_set_hr_quantum(ms: float):
        _reflib().ref_clock_set_hr_quantum(float(ms))

    def odom(self, pose, v: float, w: float):
        self._L.ref_odom(self._h, float(pose[0]), float(pose[1]), float(pose[2]), float(v), float(w))

    def clicked_pose(self, pose):
        self._L.ref_clicked_pose(self._h, float(pose[0]), float(pose[1]), float(pose[2]))

    def timer_update(self):
        self._L.ref_timer_update(self._h)

    def current_pose(self) -> np.ndarray:
        out = np.empty(3, dtype=np.float64)
        self._L.ref_current_pose(self._h, _dp(out))
        return out

    def set_inferred(self, pose):
        self._L.ref_set_inferred(self._h, _dp(np.asarray(pose, dtype=np.float64)))

    def shell_state(self) -> dict:
        return unpack_shell_state(self._shell_raw())

    def _shell_raw(self) -> np.ndarray:
        out = np.empty(23, dtype=np.float64)
        self._L.ref_shell_state(self._h, _dp(out))
        return out

    def viz_sample(self, k: int) -> np.ndarray:
        """visualize()'s weighted sub-sample (:946-958): k draws of discrete_distribution(weights_) from rng_."""
        out = (C.c_int * k)()
        self._L.ref_viz_sample(self._h, k, out)
        return np.asarray(list(out), dtype=np.int32)

    def timing(self) -> dict:
        buf = np.zeros(6, dtype=np.float64)
        cnt = C.c_int(0)
        self._L.ref_get_timing(self._h, _dp(buf), C.byref(cnt))
        keys = ["total_ms", "resample_ms", "motion_ms", "query_ms", "raycast_ms", "sensor_ms"]
        d = dict(zip(keys, buf.tolist()))
        d["count"] = cnt.value
        return d

    def reset_timing(self):
        self._L.ref_reset_timing(self._h)
