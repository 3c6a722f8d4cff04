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
    "mcl_set_shard": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "mcl_update_local_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mcl_exchange_buffers_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mcl_update_finish_dev": (C.c_int, [C.c_void_p]),
    "mcl_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mcl_ipc_import": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mcl_set_peer_pointers": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mcl_state_pointers_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "mcl_p2p_buffers_dev": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
}

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
    p = path or _build.LIB_PATH
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

    def __init__(self, device: int = 0, **params):
        self._L = load_library()
        self.params = default_params(**params)
        h = C.c_void_p()
        self._check(self._L.mcl_create(C.byref(self.params), device, C.byref(h)), "mcl_create")
        self._h = h
        self.N = int(self.params.max_particles)
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
            uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64).reshape(self.F, self.N)
            zz = None if z3n is None else np.ascontiguousarray(z3n, dtype=np.float64).reshape(self.F, 3 * self.N)
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

    def raw_weights(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N, dtype=np.float64)
        self._check(self._L.mcl_get_raw_weights(self._h, filter, _dp(out)), "mcl_get_raw_weights")
        return out

    def cdf(self, filter: int = 0) -> np.ndarray:
        out = np.empty(self.N, dtype=np.float64)
        self._check(self._L.mcl_get_cdf(self._h, filter, _dp(out)), "mcl_get_cdf")
        return out

    def sample_particles(self, k: int, filter: int = 0) -> np.ndarray:
        out = np.empty(3 * k, dtype=np.float64)
        self._check(self._L.mcl_sample_particles(self._h, filter, k, _dp(out)), "mcl_sample_particles")
        return out.reshape(3, k)

    # ---- options ----------------------------------------------------------------------
    def set_profiling(self, on: bool):
        self._check(self._L.mcl_set_profiling(self._h, int(on)), "mcl_set_profiling")

    def stage_ms(self) -> dict:
        s = MclStageMs()
        self._check(self._L.mcl_get_stage_ms(self._h, C.byref(s)), "mcl_get_stage_ms")
        return {k: getattr(s, k) for k, _ in MclStageMs._fields_}

    def set_keep_ranges(self, on: bool):
        self._check(self._L.mcl_set_keep_ranges(self._h, int(on)), "mcl_set_keep_ranges")

    def set_graphs(self, on: bool):
        """CUDA-graph replay of the steady-state host-facing update (default on)."""
        self._check(self._L.mcl_set_graphs(self._h, int(on)), "mcl_set_graphs")

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
    def set_shard(self, lo: int, count: int):
        self._check(self._L.mcl_set_shard(self._h, lo, count), "mcl_set_shard")

    def update_local_dev(self, action_dev_ptr: int, obs_dev_ptr: int, u_dev_ptr: int = 0, z_dev_ptr: int = 0):
        self._check(self._L.mcl_update_local_dev(self._h, C.c_void_p(action_dev_ptr), C.c_void_p(obs_dev_ptr), self.R,
                                                 C.c_void_p(u_dev_ptr or None), C.c_void_p(z_dev_ptr or None)),
                    "mcl_update_local_dev")

    def exchange_buffers_dev(self):
        """(ptrs[4], n_total, lo, count): device addresses of x, y, theta, raw weight."""
        ptrs = (C.c_void_p * 4)()
        n, lo, cnt = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self._check(self._L.mcl_exchange_buffers_dev(self._h, ptrs, C.byref(n), C.byref(lo), C.byref(cnt)),
                    "mcl_exchange_buffers_dev")
        return [int(p) for p in ptrs], int(n.value), int(lo.value), int(cnt.value)

    def update_finish_dev(self):
        self._check(self._L.mcl_update_finish_dev(self._h), "mcl_update_finish_dev")

    IPC_BLOB = 8 * 64   # eight cudaIpcMemHandle_t: x, y, theta of both state buffers + their packed copies

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(self.IPC_BLOB)
        self._check(self._L.mcl_ipc_export(self._h, buf, self.IPC_BLOB), "mcl_ipc_export")
        return bytes(buf.raw)

    def ipc_import(self, world: int, rank: int, blobs: bytes):
        if len(blobs) != world * self.IPC_BLOB:
            raise ValueError("expected %d bytes of IPC handles" % (world * self.IPC_BLOB))
        self._check(self._L.mcl_ipc_import(self._h, world, rank, C.c_char_p(blobs)), "mcl_ipc_import")

    def state_pointers_dev(self):
        ptrs = (C.c_void_p * 8)()
        self._check(self._L.mcl_state_pointers_dev(self._h, ptrs), "mcl_state_pointers_dev")
        return [int(p) for p in ptrs]

    def set_peer_pointers(self, world: int, rank: int, ptrs):
        arr = (C.c_void_p * (8 * world))(*[C.c_void_p(p) for p in ptrs])
        self._check(self._L.mcl_set_peer_pointers(self._h, world, rank, arr), "mcl_set_peer_pointers")

    def p2p_buffers_dev(self):
        w, part = C.c_void_p(), C.c_void_p()
        self._check(self._L.mcl_p2p_buffers_dev(self._h, C.byref(w), C.byref(part)), "mcl_p2p_buffers_dev")
        return int(w.value), int(part.value)

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._L.mcl_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "mcl_set_stream")
