"""ctypes binding of tests/emu/libemu.so -- the MCL_HD algorithm headers of the CUDA path
compiled for the CPU (TEST HARNESS ONLY; see tests/emu/emu_harness.cpp)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_lib = None


def lib():
    global _lib
    if _lib is None:
        import __graft_entry__ as ge
        path = ge.build_emu()
        L = C.CDLL(path)
        L.emu_map_create.restype = C.c_void_p
        L.emu_map_create.argtypes = [C.POINTER(C.c_int8), C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_double]
        L.emu_map_destroy.argtypes = [C.c_void_p]
        L.emu_map_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.emu_map_v8.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.emu_map_v4.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.emu_range_steps.restype = C.c_longlong
        L.emu_range_steps.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3 + [
            C.c_longlong, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.POINTER(C.c_uint8), C.POINTER(C.c_longlong)]
        L.emu_dir_map.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint8)]
        L.emu_dir_constants.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.emu_dir_sector.argtypes = [C.c_double, C.c_float, C.c_int]
        L.emu_dir_windows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_int)]
        L.emu_range_steps_dir.restype = C.c_longlong
        L.emu_range_steps_dir.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3 + [
            C.c_longlong, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
            C.POINTER(C.c_uint8), C.POINTER(C.c_int32)]
        L.emu_exact_scan.restype = C.c_longlong
        L.emu_exact_scan.argtypes = [C.POINTER(C.c_double), C.c_longlong, C.c_int, C.c_double,
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.emu_coarse_slot.argtypes = [C.c_int]
        L.emu_coarse_slots.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class EmuMap:
    def __init__(self, grid, max_range=12.0):
        L = lib()
        d = np.ascontiguousarray(grid.data, dtype=np.int8)
        self._h = L.emu_map_create(d.ctypes.data_as(C.POINTER(C.c_int8)), grid.width, grid.height,
                                   C.c_float(float(grid.resolution)), grid.origin[0], grid.origin[1], max_range)
        assert self._h, "emu_map_create failed"
        pw, ph, m = C.c_int(), C.c_int(), C.c_int()
        L.emu_map_dims(self._h, C.byref(pw), C.byref(ph), C.byref(m))
        self.PW, self.PH, self.M = pw.value, ph.value, m.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().emu_map_destroy(self._h)
            self._h = None

    def v8(self):
        out = np.empty((self.PH, self.PW), dtype=np.uint8)
        lib().emu_map_v8(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def v4(self):
        out = np.empty((self.PH, self.PW // 2), dtype=np.uint8)
        lib().emu_map_v4(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def range_steps(self, x, y, th, angles, window=None):
        """Step indices [n, R] the kernel logic produces; window=(wx0, wy0, ww, wh) selects the
        4-bit window path for window-safe particles."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        th = np.ascontiguousarray(th, dtype=np.float64)
        a = np.ascontiguousarray(angles, dtype=np.float32)
        n, R = len(x), len(a)
        out = np.empty((n, R), dtype=np.uint8)
        mode, (wx0, wy0, ww, wh) = (1, window) if window else (0, (0, 0, 0, 0))
        rep = lib().emu_range_steps(self._h, _dp(x), _dp(y), _dp(th), n, a.ctypes.data_as(C.POINTER(C.c_float)), R,
                                    mode, wx0, wy0, ww, wh, out.ctypes.data_as(C.POINTER(C.c_uint8)), None)
        return out, int(rep)


    def dir_map(self, sector):
        out = np.empty((self.PH, self.PW), dtype=np.uint8)
        lib().emu_dir_map(self._h, sector, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out

    def range_steps_dir(self, x, y, th, angles, buckets=4096, window_box=0, want_lookups=False):
        """Step indices [n, R] of the directional march (k_raycast_dir's logic); window_box > 0
        routes particles inside that box around the cloud centre through a window copy."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        th = np.ascontiguousarray(th, dtype=np.float64)
        a = np.ascontiguousarray(angles, dtype=np.float32)
        n, R = len(x), len(a)
        out = np.empty((n, R), dtype=np.uint8)
        lk = np.empty((n, R), dtype=np.int32) if want_lookups else None
        rep = lib().emu_range_steps_dir(self._h, _dp(x), _dp(y), _dp(th), n, a.ctypes.data_as(C.POINTER(C.c_float)), R,
                                        buckets, int(window_box > 0), int(window_box),
                                        out.ctypes.data_as(C.POINTER(C.c_uint8)),
                                        None if lk is None else lk.ctypes.data_as(C.POINTER(C.c_int32)))
        return (out, int(rep), lk) if want_lookups else (out, int(rep))


def dir_constants():
    s, m = C.c_int(0), C.c_double(0)
    min_buckets = lib().emu_dir_constants(C.byref(s), C.byref(m))
    return s.value, m.value, min_buckets


def dir_sector(theta, alpha, buckets):
    return lib().emu_dir_sector(float(theta), C.c_float(float(alpha)), int(buckets))


def dir_windows(em, bx0, by0, capacity):
    out = np.zeros(4 * dir_constants()[0], dtype=np.int32)
    box = lib().emu_dir_windows(em._h, int(bx0), int(by0), int(capacity), out.ctypes.data_as(C.POINTER(C.c_int)))
    return box, out.reshape(-1, 4)


def coarse_slot(k):
    return lib().emu_coarse_slot(int(k))


def coarse_slots(nc):
    return lib().emu_coarse_slots(int(nc))


def exact_scan(src, div=None, want_prefix=True, force_last_one=False):
    src = np.ascontiguousarray(src, dtype=np.float64)
    n = len(src)
    tot = C.c_double(0)
    out = np.empty(n, dtype=np.float64) if want_prefix else None
    nop = lib().emu_exact_scan(_dp(src), n, int(div is not None), float(div if div is not None else 1.0),
                               C.byref(tot), None if out is None else _dp(out), int(force_last_one))
    return tot.value, out, int(nop)
