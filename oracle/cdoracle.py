"""ctypes view of oracle/libcd_oracle.so (the plain-C restatement, oracle/cd_oracle.c).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Params(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("extent", C.c_double * 3), ("key_bits", C.c_int)]


class Counters(C.Structure):
    _fields_ = [("internal_visits", C.c_uint64), ("box_tests", C.c_uint64), ("leaf_hits", C.c_uint64),
                ("narrow_calls", C.c_uint64), ("max_stack", C.c_uint32)]


class Timing(C.Structure):
    _fields_ = [("ms_morton", C.c_double), ("ms_sort", C.c_double), ("ms_hierarchy", C.c_double),
                ("ms_refit", C.c_double), ("ms_query", C.c_double), ("pairs", C.c_uint64), ("ctr", Counters)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "libcd_oracle.so")])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcd_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.cdo_morton_of_centroid.restype = C.c_uint64
        _LIB.cdo_morton_of_centroid.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(Params)]
        _LIB.cdo_self_collide.restype = C.c_uint64
        _LIB.cdo_brute_force.restype = C.c_uint64
        _LIB.cdo_run.restype = C.c_uint64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def default_params(key_bits=63):
    p = Params()
    lib().cdo_default_params(C.byref(p))
    p.key_bits = key_bits
    return p


def make_params(origin, extent, key_bits=63):
    p = Params()
    p.origin[:] = list(origin)
    p.extent[:] = list(extent)
    p.key_bits = key_bits
    return p


def auto_params(xyz, key_bits=63):
    xyz = np.ascontiguousarray(xyz, np.float32)
    p = Params()
    lib().cdo_auto_params(_p(xyz, C.c_float), C.c_uint32(xyz.shape[0]), C.c_int(key_bits), C.byref(p))
    return p


def morton_keys(xyz, idx, params):
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    n = idx.shape[0]
    keys = np.empty(n, np.uint64)
    lib().cdo_morton_keys(_p(xyz, C.c_float), _p(idx, C.c_uint32), C.c_uint32(n), C.byref(params), _p(keys, C.c_uint64))
    return keys


def sort_keys(keys):
    keys = np.ascontiguousarray(keys, np.uint64)
    n = keys.shape[0]
    sk = np.empty(n, np.uint64)
    si = np.empty(n, np.uint32)
    lib().cdo_sort(_p(keys, C.c_uint64), C.c_uint32(n), _p(sk, C.c_uint64), _p(si, C.c_uint32))
    return sk, si


def hierarchy(skeys):
    skeys = np.ascontiguousarray(skeys, np.uint64)
    n = skeys.shape[0]
    m = max(n - 1, 0)
    first, last, split, left, right = (np.empty(m, np.int32) for _ in range(5))
    parent = np.empty(2 * n - 1, np.int32)
    lib().cdo_hierarchy(_p(skeys, C.c_uint64), C.c_uint32(n), _p(first, C.c_int32), _p(last, C.c_int32),
                        _p(split, C.c_int32), _p(left, C.c_int32), _p(right, C.c_int32), _p(parent, C.c_int32))
    return dict(first=first, last=last, split=split, left=left, right=right, parent=parent)


def refit(xyz, idx, sids, h):
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    n = idx.shape[0]
    bounds = np.zeros((2 * n - 1, 6), np.float64)
    lib().cdo_refit(_p(xyz, C.c_float), _p(idx, C.c_uint32), _p(sids, C.c_uint32), C.c_uint32(n),
                    _p(h["left"], C.c_int32), _p(h["right"], C.c_int32), _p(h["parent"], C.c_int32),
                    _p(bounds, C.c_double))
    return bounds


def _take_pairs(ptr, n):
    if n == 0:
        out = np.empty((0, 2), np.uint32)
    else:
        out = np.ctypeslib.as_array(ptr, shape=(n, 2)).copy()
    lib().cdo_free(ptr)
    return out


def self_collide(xyz, idx, sids, h, bounds):
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    ptr = C.POINTER(C.c_uint32)()
    ctr = Counters()
    n = lib().cdo_self_collide(_p(xyz, C.c_float), _p(idx, C.c_uint32), _p(sids, C.c_uint32),
                               C.c_uint32(idx.shape[0]), _p(h["left"], C.c_int32), _p(h["right"], C.c_int32),
                               _p(bounds, C.c_double), C.byref(ptr), C.byref(ctr))
    return _take_pairs(ptr, n), ctr


def brute_force(xyz, idx):
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    ptr = C.POINTER(C.c_uint32)()
    n = lib().cdo_brute_force(_p(xyz, C.c_float), _p(idx, C.c_uint32), C.c_uint32(idx.shape[0]), C.byref(ptr))
    return sort_pairs(_take_pairs(ptr, n))


def sort_pairs(pairs):
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    if pairs.shape[0]:
        lib().cdo_sort_pairs(_p(pairs, C.c_uint32), C.c_uint64(pairs.shape[0]))
    return pairs


def run(xyz, idx, params):
    """whole pipeline, one thread; returns (sorted pairs, Timing)"""
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    ptr = C.POINTER(C.c_uint32)()
    tm = Timing()
    n = lib().cdo_run(_p(xyz, C.c_float), C.c_uint32(xyz.shape[0]), _p(idx, C.c_uint32), C.c_uint32(idx.shape[0]),
                      C.byref(params), C.byref(ptr), C.byref(tm))
    return _take_pairs(ptr, n), tm


def tri_contact(t18):
    t = np.ascontiguousarray(t18, np.float64).reshape(18)
    return int(lib().cdo_tri_contact(_p(t, C.c_double)))


def box_overlap(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    return int(lib().cdo_box_overlap(_p(a, C.c_double), _p(b, C.c_double)))


def morton_of_centroid(x, y, z, params):
    return int(lib().cdo_morton_of_centroid(x, y, z, C.byref(params)))
