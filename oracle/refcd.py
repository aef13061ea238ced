"""ctypes view of oracle/_ref/libref_cd.so — the reference's own host functions
(oracle/ref_driver.cu). TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

The library is compiled in the build container (where /root/reference exists) and
travels to the GPU box as a built file; available() says whether it is there.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_cd.so")
_LIB = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_PATH)
        _LIB.ref_load_obj.restype = C.c_void_p
        _LIB.ref_from_arrays.restype = C.c_void_p
        _LIB.ref_from_arrays_keyed.restype = C.c_void_p
        _LIB.ref_collide.restype = C.c_uint64
        _LIB.ref_collide_mt.restype = C.c_uint64
        _LIB.ref_morton3D.restype = C.c_uint64
        _LIB.ref_morton3D.argtypes = [C.c_double] * 3
        for f in ("ref_free", "ref_num_verts", "ref_num_tris", "ref_get_mesh", "ref_get_sorted", "ref_build",
                  "ref_get_nodes", "ref_collide", "ref_get_pairs", "ref_get_timing", "ref_gpu_run"):
            getattr(_LIB, f).argtypes = None
        _LIB.ref_num_verts.restype = C.c_uint32
        _LIB.ref_num_tris.restype = C.c_uint32
        _LIB.ref_build.restype = C.c_uint
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefMesh:
    """Owns one reference-side mesh (vectors of vec3f / Triangle / morton + Node arrays)."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        self.n = lib().ref_num_tris(self.h)
        self.nv = lib().ref_num_verts(self.h)

    @classmethod
    def from_obj(cls, path):
        return cls(lib().ref_load_obj(os.fsencode(path)))

    @classmethod
    def from_arrays(cls, xyz, idx):
        xyz = np.ascontiguousarray(xyz, np.float32)
        idx = np.ascontiguousarray(idx, np.uint32)
        return cls(lib().ref_from_arrays(_p(xyz, C.c_float), C.c_uint32(xyz.shape[0]), _p(idx, C.c_uint32),
                                         C.c_uint32(idx.shape[0])))

    @classmethod
    def from_arrays_keyed(cls, xyz, idx, keys):
        """caller-supplied UNIQUE sort keys instead of morton3D (whose box is hard-coded, morton.h:43-58); everything
        after the key is the reference's own code. The pair set does not depend on the keys."""
        xyz = np.ascontiguousarray(xyz, np.float32)
        idx = np.ascontiguousarray(idx, np.uint32)
        keys = np.ascontiguousarray(keys, np.uint64)
        assert keys.shape[0] == idx.shape[0]
        return cls(lib().ref_from_arrays_keyed(_p(xyz, C.c_float), C.c_uint32(xyz.shape[0]), _p(idx, C.c_uint32),
                                               C.c_uint32(idx.shape[0]), _p(keys, C.c_uint64)))

    def close(self):
        if self.h:
            lib().ref_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def mesh(self):
        xyz = np.empty((self.nv, 3), np.float32)
        idx = np.empty((self.n, 3), np.uint32)
        lib().ref_get_mesh(self.h, _p(xyz, C.c_float), _p(idx, C.c_uint32))
        return xyz, idx

    def sorted(self):
        keys = np.empty(self.n, np.uint64)
        ids = np.empty(self.n, np.uint32)
        lib().ref_get_sorted(self.h, _p(keys, C.c_uint64), _p(ids, C.c_uint32))
        return keys, ids

    def build(self):
        return int(lib().ref_build(self.h))

    def nodes(self):
        n = self.n
        left = np.empty(n - 1, np.int32)
        right = np.empty(n - 1, np.int32)
        bounded = np.empty(n - 1, np.uint32)
        parent = np.empty(2 * n - 1, np.int32)
        bounds = np.empty((2 * n - 1, 6), np.float64)
        lib().ref_get_nodes(self.h, _p(left, C.c_int32), _p(right, C.c_int32), _p(parent, C.c_int32),
                            _p(bounded, C.c_uint32), _p(bounds, C.c_double))
        return dict(left=left, right=right, parent=parent, bounded=bounded, bounds=bounds)

    def collide(self, nthreads=1):
        """nthreads > 1: the reference's per-query function over host threads (our parallel loop)"""
        cnt = lib().ref_collide(self.h) if nthreads <= 1 else lib().ref_collide_mt(self.h, C.c_int(nthreads))
        out = np.empty((cnt, 2), np.uint32)
        if cnt:
            lib().ref_get_pairs(self.h, _p(out, C.c_uint32))
        return out

    def gpu_run(self, pair_cap=None, repeats=3):
        """the reference's own GPU kernels (main.cu:92,99,107,142 launch configurations) on this mesh, on the current
        CUDA device -> (pairs (count, 2) uint32 in discovery order, {stage: ms}); raises without a device"""
        cap = int(pair_cap or max(500, self.n))
        ms = (C.c_float * 4)()
        cnt = C.c_uint64()
        rc = lib().ref_gpu_run(self.h, C.c_uint32(cap), C.c_int(repeats), ms, C.byref(cnt))
        if rc != 0:
            raise RuntimeError({-1: "no CUDA device", -2: "CUDA error in the reference kernels",
                                -3: "more pairs than pair_cap"}.get(rc, f"ref_gpu_run: {rc}"))
        out = np.empty((cnt.value, 2), np.uint32)
        if cnt.value:
            lib().ref_get_pairs(self.h, _p(out, C.c_uint32))
        return out, dict(zip(("fillLeafNodes", "generateHierarchyParallel", "calBoundingBox", "findCollisions"),
                             [float(x) for x in ms]))

    def timing(self):
        t = np.zeros(5, np.float64)
        lib().ref_get_timing(self.h, _p(t, C.c_double))
        return dict(zip(("load", "fill", "hierarchy", "refit", "query"), t.tolist()))


def morton3D(x, y, z):
    return int(lib().ref_morton3D(x, y, z))


def tri_contact(t18):
    t = np.ascontiguousarray(t18, np.float64).reshape(18)
    return int(lib().ref_tri_contact(_p(t, C.c_double)))


def box_overlap(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return int(lib().ref_box_overlap(_p(a, C.c_double), _p(b, C.c_double)))


def range_split(keys, i):
    keys = np.ascontiguousarray(keys, np.uint64)
    out = np.zeros(3, np.int32)
    lib().ref_range_split(_p(keys, C.c_uint64), C.c_int(keys.shape[0]), C.c_int(i), _p(out, C.c_int32))
    return tuple(int(v) for v in out)
