"""ctypes binding of the C ABI in ``include/mcl_b200.h`` (libmcl_b200.so).

This is the binding a Python host would use; the C++ host mirror
(``host/particle_filter.hpp``) binds the same symbols.  There is no fallback: if the CUDA
library is missing or no device is visible, loading / ``mcl_create`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int8_p = C.POINTER(C.c_int8)
c_uint8_p = C.POINTER(C.c_uint8)

MCL_OK = 0
MCL_ERR_INVALID = -1
MCL_ERR_NO_DEVICE = -2
MCL_ERR_CUDA = -3
MCL_ERR_NO_MAP = -4
MCL_ERR_UNSUPPORTED = -5
MCL_ERR_NO_FREE_SPACE = -6


class MclParams(C.Structure):
    _fields_ = [("max_particles", C.c_int32), ("max_viz_particles", C.c_int32), ("angle_step", C.c_int32),
                ("squash_factor", C.c_double), ("max_range", C.c_double),
                ("z_short", C.c_double), ("z_max", C.c_double), ("z_rand", C.c_double),
                ("z_hit", C.c_double), ("sigma_hit", C.c_double),
                ("motion_dispersion_x", C.c_double), ("motion_dispersion_y", C.c_double),
                ("motion_dispersion_theta", C.c_double),
                ("seed", C.c_uint64), ("num_filters", C.c_int32)]


class MclNoise(C.Structure):
    _fields_ = [("u_resample", c_double_p), ("z_motion", c_double_p)]


class MclStageMs(C.Structure):
    _fields_ = [("cdf", C.c_float), ("resample_motion", C.c_float), ("raycast_weight", C.c_float),
                ("normalize_pose", C.c_float), ("total", C.c_float), ("ray_march", C.c_float), ("exchange", C.c_float)]


# every symbol include/mcl_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mcl_default_params": (None, [C.POINTER(MclParams)]),
    "mcl_last_error": (C.c_char_p, []),
    "mcl_status_str": (C.c_char_p, [C.c_int]),
    "mcl_abi_version": (C.c_int, []),
    "mcl_device_count": (C.c_int, []),
    "mcl_create": (C.c_int, [C.POINTER(MclParams), C.c_int, C.POINTER(C.c_void_p)]),
    "mcl_destroy": (C.c_int, [C.c_void_p]),
    "mcl_set_map": (C.c_int, [C.c_void_p, c_int8_p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_double]),
    "mcl_max_range_px": (C.c_int, [C.c_void_p]),
    "mcl_get_sensor_table": (C.c_int, [C.c_void_p, c_double_p]),
    "mcl_set_sensor_table": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "mcl_set_beam_angles": (C.c_int, [C.c_void_p, c_float_p, C.c_int]),
    "mcl_init_pose": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "mcl_init_global": (C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_double_p]),
    "mcl_num_free_cells": (C.c_int, [C.c_void_p]),
    "mcl_set_particles": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    "mcl_get_particles": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mcl_get_weights": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mcl_update": (C.c_int, [C.c_void_p, c_double_p, c_float_p, C.c_int, C.POINTER(MclNoise), c_double_p]),
    "mcl_update_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mcl_read_pose": (C.c_int, [C.c_void_p, c_double_p]),
    "mcl_synchronize": (C.c_int, [C.c_void_p]),
    "mcl_expected_pose": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mcl_calc_range_many": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_float_p]),
    "mcl_cast_ray": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, c_float_p]),
    "mcl_get_resample_indices": (C.c_int, [C.c_void_p, C.c_int, c_int32_p]),
    "mcl_get_ranges": (C.c_int, [C.c_void_p, C.c_int, c_float_p]),
    "mcl_get_range_steps": (C.c_int, [C.c_void_p, C.c_int, c_uint8_p]),
    "mcl_get_range_steps16": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint16)]),
    "mcl_get_raw_weights": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mcl_get_cdf": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "mcl_sample_particles": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "mcl_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "mcl_get_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(MclStageMs)]),
    "mcl_set_keep_ranges": (C.c_int, [C.c_void_p, C.c_int]),
    "mcl_kernel_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "mcl_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mcl_set_graphs": (C.c_int, [C.c_void_p, C.c_int]),
    "mcl_set_ray_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "mcl_ray_stage_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4),
    "mcl_get_dir_map": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mcl_microbench_gather": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_int, c_double_p]),
    "mcl_update_dev_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mcl_sample_particles_u": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_int32_p]),
    "mcl_get_kernel_ms": (C.c_int, [C.c_void_p, C.c_char_p, c_float_p, C.c_int, C.POINTER(C.c_int)]),
    "mcl_debug_pass_cycles": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
    "mcl_shard_create": (C.c_int, [C.POINTER(MclParams), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mcl_shard_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mcl_shard_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mcl_shard_connect_local": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mcl_shard_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mcl_shard_set_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mcl_shard_set_route": (C.c_int, [C.c_void_p, C.c_int]),
    "mcl_set_pdl": (C.c_int, [C.c_void_p, C.c_int]),
    "mcl_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mcl_create_sharded": (C.c_int, [C.POINTER(MclParams), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mcl_sharded_gather": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
}

BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)

_lib = None


class MclError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str):
        super().__init__("%s failed: status %d (%s)" % (where, status, detail))
        self.status = status


def load_library(path: str | None = None):
    """dlopen libmcl_b200.so and bind every declared symbol.  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # MCL_B200_LIB: development aid (A/B builds of the same library, scripts/build_variant.sh); still CUDA-only
    p = path or os.environ.get("MCL_B200_LIB") or _build.LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            "%s is missing: build it with `python -m monte_carlo_localization_b200.build` "
            "(the MCL update is CUDA-only; there is no CPU fallback)" % p)
    L = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _fp(a):
    return None if a is None else a.ctypes.data_as(c_float_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def default_params(**kw) -> MclParams:
    p = MclParams()
    load_library().mcl_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError("mcl_params has no field %r" % k)
        setattr(p, k, v)
    return p


def nccl_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and hands the 128 bytes to the others)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.mcl_nccl_unique_id(buf, 128)
    if rc != MCL_OK:
        raise MclError(rc, "mcl_nccl_unique_id", L.mcl_last_error().decode(errors="replace"))
    return bytes(buf.raw)


def microbench_gather(shared: bool, array_bytes: int = 4 << 20, iters: int = 4096, device: int = 0) -> float:
    """Random single-byte gathers per second from shared memory or an L2-resident array."""
    L = load_library()
    out = C.c_double(0.0)
    rc = L.mcl_microbench_gather(device, int(shared), array_bytes, iters, C.byref(out))
    if rc != MCL_OK:
        raise MclError(rc, "mcl_microbench_gather", L.mcl_last_error().decode(errors="replace"))
    return float(out.value)


class MclContext:
    """One ``mcl_ctx``: a ParticleFilter's device state (or a batch of independent ones)."""

    def __init__(self, device: int = 0, shard=None, nccl_id: bytes | None = None, **params):
        """shard = (world, rank): one rank of a particle-sharded filter (``max_particles`` = particles of
        the WHOLE filter); with ``nccl_id`` (mcl_nccl_unique_id of rank 0) the library owns the NCCL
        communicator and connects the ranks itself, otherwise connect with ``shard_connect*``."""
        self._L = load_library()
        self.params = default_params(**params)
        h = C.c_void_p()
        self.world, self.rank = (1, 0) if shard is None else (int(shard[0]), int(shard[1]))
        if shard is None:
            self._check(self._L.mcl_create(C.byref(self.params), device, C.byref(h)), "mcl_create")
        elif nccl_id is not None:
            self._check(self._L.mcl_create_sharded(C.byref(self.params), device, self.world, self.rank,
                                                   C.c_char_p(nccl_id), C.byref(h)), "mcl_create_sharded")
        else:
            self._check(self._L.mcl_shard_create(C.byref(self.params), device, self.world, self.rank, C.byref(h)),
                        "mcl_shard_create")
        self._h = h
        self.NG = int(self.params.max_particles)            # particles of the whole filter
        self.N = self.NG // self.world                      # particles this context holds
        self.glo = self.N * self.rank
        self._hook = None
        self.F = int(self.params.num_filters)
        self.R = 0
        self.M = 0
        self.device = device

    def _check(self, rc: int, where: str):
        if rc != MCL_OK:
            raise MclError(rc, where, self._L.mcl_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._L.mcl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup -------------------------------------------------------------------------
    def set_map(self, grid):
        d = np.ascontiguousarray(grid.data, dtype=np.int8)
        self._check(self._L.mcl_set_map(self._h, d.ctypes.data_as(c_int8_p), grid.width, grid.height,
                                        C.c_float(float(grid.resolution)), grid.origin[0], grid.origin[1],
                                        grid.origin[2]), "mcl_set_map")
        self.M = self._L.mcl_max_range_px(self._h)

    def sensor_table(self) -> np.ndarray:
        t = np.empty((self.M + 1) * (self.M + 1), dtype=np.float64)
        self._check(self._L.mcl_get_sensor_table(self._h, _dp(t)), "mcl_get_sensor_table")
        return t

    def set_sensor_table(self, table_colmajor):
        t = np.ascontiguousarray(table_colmajor, dtype=np.float64).reshape(-1)
        tw = int(round(len(t) ** 0.5))
        self._check(self._L.mcl_set_sensor_table(self._h, _dp(t), tw), "mcl_set_sensor_table")

    def set_beam_angles(self, angles):
        a = np.ascontiguousarray(angles, dtype=np.float32)
        self._check(self._L.mcl_set_beam_angles(self._h, _fp(a), len(a)), "mcl_set_beam_angles")
        self.R = len(a)

    def num_free_cells(self) -> int:
        return self._L.mcl_num_free_cells(self._h)

    def init_pose(self, pose, normals_3n=None, filter: int = 0):
        p = np.asarray(pose, dtype=np.float64)
        z = None if normals_3n is None else np.ascontiguousarray(normals_3n, dtype=np.float64)
        self._check(self._L.mcl_init_pose(self._h, filter, _dp(p), _dp(z)), "mcl_init_pose")

    def init_global(self, cell_ordinal=None, theta=None, filter: int = 0):
        c = None if cell_ordinal is None else np.ascontiguousarray(cell_ordinal, dtype=np.int32)
        t = None if theta is None else np.ascontiguousarray(theta, dtype=np.float64)
        self._check(self._L.mcl_init_global(self._h, filter, _ip(c), _dp(t)), "mcl_init_global")

    def set_particles(self, particles_colmajor=None, weights=None, filter: int = 0):
        p = None if particles_colmajor is None else np.ascontiguousarray(particles_colmajor, dtype=np.float64).reshape(-1)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self._check(self._L.mcl_set_particles(self._h, filter, _dp(p), _dp(w)), "mcl_set_particles")

    def get_particles(self, filter: int = 0) -> np.ndarray:
        p = np.empty(3 * self.N, dtype=np.float64)
        self._check(self._L.mcl_get_particles(self._h, filter, _dp(p)), "mcl_get_particles")
        return p.reshape(3, self.N)

    def get_weights(self, filter: int = 0) -> np.ndarray:
        w = np.empty(self.N, dtype=np.float64)
        self._check(self._L.mcl_get_weights(self._h, filter, _dp(w)), "mcl_get_weights")
        return w

    # ---- the path ----------------------------------------------------------------------
    def update(self, action, obs, u=None, z3n=None) -> np.ndarray:
        """MCL(action, obs) + expected_pose().  Batch: action [F,3], obs [F,R], u [F,N], z [F,3N]."""
        a = np.ascontiguousarray(action, dtype=np.float64).reshape(-1)
        o = np.ascontiguousarray(obs, dtype=np.float32).reshape(-1)
        if a.size != 3 * self.F or o.size != self.R * self.F:
            raise ValueError("action/obs shape does not match num_filters=%d, beams=%d" % (self.F, self.R))
        noise = None
        keep = []
        if u is not None or z3n is not None:
            arr = (MclNoise * self.F)()
            # injected noise covers the WHOLE filter (a sharded rank indexes it by the global slot)
            uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64).reshape(self.F, self.NG)
            zz = None if z3n is None else np.ascontiguousarray(z3n, dtype=np.float64).reshape(self.F, 3 * self.NG)
            keep = [uu, zz]
            for f in range(self.F):
                arr[f].u_resample = None if uu is None else uu[f].ctypes.data_as(c_double_p)
                arr[f].z_motion = None if zz is None else zz[f].ctypes.data_as(c_double_p)
            noise = arr
        pose = np.empty(3 * self.F, dtype=np.float64)
        self._check(self._L.mcl_update(self._h, _dp(a), _fp(o), self.R, noise, _dp(pose)), "mcl_update")
        del keep
        return pose.reshape(self.F, 3) if self.F > 1 else pose

    def update_dev(self, action_dev_ptr: int, obs_dev_ptr: int):
        self._check(self._L.mcl_update_dev(self._h, C.c_void_p(action_dev_ptr), C.c_void_p(obs_dev_ptr), self.R),
                    "mcl_update_dev")

    def update_dev_noise(self, action_dev_ptr: int, obs_dev_ptr: int, u_dev_ptr: int = 0, z_dev_ptr: int = 0):
        self._check(self._L.mcl_update_dev_noise(self._h, C.c_void_p(action_dev_ptr), C.c_void_p(obs_dev_ptr), self.R,
                                                 C.c_void_p(u_dev_ptr or None), C.c_void_p(z_dev_ptr or None)),
                    "mcl_update_dev_noise")

    def read_pose(self) -> np.ndarray:
        pose = np.empty(3 * self.F, dtype=np.float64)
        self._check(self._L.mcl_read_pose(self._h, _dp(pose)), "mcl_read_pose")
        return pose.reshape(self.F, 3) if self.F > 1 else pose

    def synchronize(self):
        self._check(self._L.mcl_synchronize(self._h), "mcl_synchronize")

    def expected_pose(self, filter: int = 0) -> np.ndarray:
        pose = np.empty(3, dtype=np.float64)
        self._check(self._L.mcl_expected_pose(self._h, filter, _dp(pose)), "mcl_expected_pose")
        return pose

    def calc_range_many(self, queries_colmajor) -> np.ndarray:
        q = np.ascontiguousarray(queries_colmajor, dtype=np.float64).reshape(-1)
        n = q.size // 3
        out = np.empty(n, dtype=np.float32)
        self._check(self._L.mcl_calc_range_many(self._h, _dp(q), n, _fp(out)), "mcl_calc_range_many")
        return out

    def cast_ray(self, x: float, y: float, angle: float) -> float:
        out = C.c_float(0)
        self._check(self._L.mcl_cast_ray(self._h, x, y, angle, C.byref(out)), "mcl_cast_ray")
        return float(out.value)

    # ---- stage read-backs -------------------------------------------------------------
    def resample_indices(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N, dtype=np.int32)
        self._check(self._L.mcl_get_resample_indices(self._h, filter, _ip(out)), "mcl_get_resample_indices")
        return out

    def ranges(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N * self.R, dtype=np.float32)
        self._check(self._L.mcl_get_ranges(self._h, filter, _fp(out)), "mcl_get_ranges")
        return out.reshape(self.N, self.R)

    def range_steps(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N * self.R, dtype=np.uint8)
        self._check(self._L.mcl_get_range_steps(self._h, filter, out.ctypes.data_as(c_uint8_p)), "mcl_get_range_steps")
        return out.reshape(self.N, self.R)

    def range_steps16(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N * self.R, dtype=np.uint16)
        self._check(self._L.mcl_get_range_steps16(self._h, filter, out.ctypes.data_as(C.POINTER(C.c_uint16))), "mcl_get_range_steps16")
        return out.reshape(self.N, self.R)

    def raw_weights(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N, dtype=np.float64)
        self._check(self._L.mcl_get_raw_weights(self._h, filter, _dp(out)), "mcl_get_raw_weights")
        return out

    def cdf(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N, dtype=np.float64)
        self._check(self._L.mcl_get_cdf(self._h, filter, _dp(out)), "mcl_get_cdf")
        return out

    def sample_particles(self, k: int, filter: int = 0, u=None, return_indices: bool = False):
        """visualize()'s weighted sub-sample (:946-958).  u: injected canonical uniforms (k of them)."""
        out = np.empty(3 * k, dtype=np.float64)
        idx = np.empty(k, dtype=np.int32)
        uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        if uu is not None and uu.size != k:
            raise ValueError("need %d uniforms" % k)
        self._check(self._L.mcl_sample_particles_u(self._h, filter, k, _dp(uu), _dp(out), _ip(idx)), "mcl_sample_particles_u")
        return (out.reshape(3, k), idx) if return_indices else out.reshape(3, k)

    # ---- options ----------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._check(self._L.mcl_set_profiling(self._h, int(on)), "mcl_set_profiling")

    def stage_ms(self) -> dict:
        s = MclStageMs()
        self._check(self._L.mcl_get_stage_ms(self._h, C.byref(s)), "mcl_get_stage_ms")
        return {k: getattr(s, k) for k, _ in MclStageMs._fields_}

    def kernel_ms(self) -> list:
        """[(kernel name, device ms)] of the last profiled update, in launch order."""
        cap = 64
        names = C.create_string_buffer(48 * cap)
        ms = np.zeros(cap, dtype=np.float32)
        n = C.c_int(0)
        self._check(self._L.mcl_get_kernel_ms(self._h, names, _fp(ms), cap, C.byref(n)), "mcl_get_kernel_ms")
        return [(names.raw[48 * i:48 * i + 48].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(min(n.value, cap))]

    def debug_pass_cycles(self, pass_kind: int, read: bool = False):
        out = (C.c_uint64 * 8)() if read else None
        self._check(self._L.mcl_debug_pass_cycles(self._h, pass_kind, out), "mcl_debug_pass_cycles")
        return [int(v) for v in out] if read else None

    def set_keep_ranges(self, on: bool):
        self._check(self._L.mcl_set_keep_ranges(self._h, int(on)), "mcl_set_keep_ranges")

    def set_graphs(self, on: bool):
        """CUDA-graph replay of the steady-state host-facing update (default on)."""
        self._check(self._L.mcl_set_graphs(self._h, int(on)), "mcl_set_graphs")

    def set_pdl(self, on: bool):
        """Programmatic dependent launches between the kernels of an update (default on)."""
        self._check(self._L.mcl_set_pdl(self._h, int(on)), "mcl_set_pdl")

    def set_ray_mode(self, mode: int):
        """0 auto, 1 isotropic skip-map kernel only, 2 directional stage always."""
        self._check(self._L.mcl_set_ray_mode(self._h, int(mode)), "mcl_set_ray_mode")

    def ray_stage_info(self) -> dict:
        v = [C.c_int(0) for _ in range(4)]
        self._check(self._L.mcl_ray_stage_info(self._h, *[C.byref(x) for x in v]), "mcl_ray_stage_info")
        return dict(zip(("directional_ready", "last_mode", "box_cells", "units"), (int(x.value) for x in v)))

    def dir_map(self, sector: int) -> np.ndarray:
        pw, ph = C.c_int(0), C.c_int(0)
        self._check(self._L.mcl_get_dir_map(self._h, sector, None, C.byref(pw), C.byref(ph)), "mcl_get_dir_map")
        out = np.empty((ph.value, pw.value), dtype=np.uint8)
        self._check(self._L.mcl_get_dir_map(self._h, sector, out.ctypes.data_as(C.POINTER(C.c_uint8)), None, None),
                    "mcl_get_dir_map")
        return out

    def kernel_launches(self) -> int:
        n = C.c_int64(0)
        self._check(self._L.mcl_kernel_launches(self._h, C.byref(n)), "mcl_kernel_launches")
        return int(n.value)

    # ---- particle sharding -------------------------------------------------------------
    SHARD_BLOB = 512

    def shard_export(self) -> bytes:
        buf = C.create_string_buffer(self.SHARD_BLOB)
        self._check(self._L.mcl_shard_export(self._h, buf, self.SHARD_BLOB), "mcl_shard_export")
        return bytes(buf.raw)

    def shard_connect(self, blobs: bytes):
        if len(blobs) != self.world * self.SHARD_BLOB:
            raise ValueError("expected %d bytes of exchange-arena handles" % (self.world * self.SHARD_BLOB))
        self._check(self._L.mcl_shard_connect(self._h, C.c_char_p(blobs)), "mcl_shard_connect")

    def shard_connect_local(self, ranks):
        """ranks: the `world` MclContext objects of this process, in rank order."""
        arr = (C.c_void_p * len(ranks))(*[r._h for r in ranks])
        self._check(self._L.mcl_shard_connect_local(self._h, arr), "mcl_shard_connect_local")

    def shard_set_exchange(self, fused: bool, hook=None):
        """fused: the kernels wait for their peers themselves (one rank per GPU).  Otherwise the library
        calls hook() -> int wherever all ranks must have published (hook None: NCCL barrier)."""
        cb = None
        if hook is not None:
            cb = BARRIER_FN(lambda _user: int(hook() or 0))
        self._hook = cb   # keep the trampoline alive
        self._check(self._L.mcl_shard_set_exchange(self._h, int(fused), C.cast(cb, C.c_void_p) if cb else None, None),
                    "mcl_shard_set_exchange")

    def shard_set_route(self, two_hop):
        """Two-hop request routing (True), every rank testing all draws (False) or the library's choice by world
        size (None); same result bit for bit."""
        self._check(self._L.mcl_shard_set_route(self._h, -1 if two_hop is None else int(bool(two_hop))), "mcl_shard_set_route")

    def sharded_gather(self):
        """(particles [3, NG], weights [NG]) of the whole filter, on every rank (NCCL, collective)."""
        p = np.empty(3 * self.NG, dtype=np.float64)
        w = np.empty(self.NG, dtype=np.float64)
        self._check(self._L.mcl_sharded_gather(self._h, _dp(p), _dp(w)), "mcl_sharded_gather")
        return p.reshape(3, self.NG), w

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._L.mcl_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "mcl_set_stream")
