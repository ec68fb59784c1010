"""oracle_py — ctypes loader of the CPU oracle (oracle/_build/liboracle.so, built from oracle/rt_oracle.cpp by
__graft_entry__.build()). TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg. The product never imports it."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.oracle_load.restype = C.c_void_p
        L.oracle_load.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int]
        L.oracle_load2.restype = C.c_void_p
        L.oracle_load2.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int, C.c_int]
        L.oracle_primary_ids_brute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_n_top.argtypes = [C.c_void_p]
        L.oracle_n_nodes.argtypes = [C.c_void_p]
        L.oracle_leaf_order.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_primary_ids.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_render.restype = C.c_ulonglong
        L.oracle_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.oracle_xorwow.argtypes = [C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def load_ppm(path):
    """Binary P6 -> (h, w, 3) uint8."""
    with open(path, "rb") as f:
        data = f.read()
    parts = data.split(None, 4)
    assert parts[0] == b"P6", path
    w, h, mx = int(parts[1]), int(parts[2]), int(parts[3])
    off = len(data) - w * h * 3
    return np.frombuffer(data, dtype=np.uint8, offset=off).reshape(h, w, 3).copy()


class Oracle:
    def __init__(self, sd_bytes, images=(), bvh=True):
        """bvh=False: skip the reference's O(n^2)-per-level BVH constructor (large scenes); only primary_ids_brute works."""
        self._sd = bytes(sd_bytes)
        self._imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        arr = (C.c_void_p * max(len(self._imgs), 1))(*[im.ctypes.data for im in self._imgs])
        self._h = lib().oracle_load2(self._sd, len(self._sd), arr, len(self._imgs), 1 if bvh else 0)
        if not self._h:
            raise ValueError("oracle_load: bad scene description")
        self.n_top = lib().oracle_n_top(self._h)
        self.n_nodes = lib().oracle_n_nodes(self._h)

    def close(self):
        if self._h:
            lib().oracle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def leaf_order(self):
        rank = np.zeros(self.n_top, dtype=np.int32)
        lib().oracle_leaf_order(self._h, rank.ctypes.data)
        return rank

    def primary_ids(self, nx, ny):
        obj = np.empty((ny, nx), dtype=np.int32)
        mat = np.empty((ny, nx), dtype=np.int32)
        t = np.empty((ny, nx), dtype=np.float32)
        lib().oracle_primary_ids(self._h, nx, ny, obj.ctypes.data, mat.ctypes.data, t.ctypes.data)
        return obj, mat, t

    def primary_ids_brute(self, nx, ny):
        """Same query as primary_ids with no hierarchy at all: every object, in creation order, behind its own box test."""
        obj = np.empty((ny, nx), dtype=np.int32)
        mat = np.empty((ny, nx), dtype=np.int32)
        t = np.empty((ny, nx), dtype=np.float32)
        lib().oracle_primary_ids_brute(self._h, nx, ny, obj.ctypes.data, mat.ctypes.data, t.ctypes.data)
        return obj, mat, t

    def render(self, nx, ny, spp, background=(0, 0, 0), gradient=False, max_depth=50, gamma=2.2):
        fb = np.empty((ny, nx, 3), dtype=np.float32)
        bg = np.asarray(background, dtype=np.float32)
        rays = lib().oracle_render(self._h, nx, ny, spp, max_depth, bg.ctypes.data, int(bool(gradient)), gamma, fb.ctypes.data)
        return fb, int(rays)


def primary_ids(sd_bytes, nx, ny, images=()):
    o = Oracle(sd_bytes, images)
    try:
        return o.primary_ids(nx, ny)
    finally:
        o.close()


def xorwow(seed, n):
    raw = np.empty(n, dtype=np.uint32)
    uni = np.empty(n, dtype=np.float32)
    lib().oracle_xorwow(seed, n, raw.ctypes.data, uni.ctypes.data)
    return raw, uni
