"""Synthetic workloads (BASELINE.json configs C1..C5) — ctypes view of csrc/meshgen.c.

Host-side workload utilities only; the collision library never calls these.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "lib", "libb200cd_meshgen.so")
_LIB = None

# the reference's hard-coded Morton normalisation box (reference morton.h:45,51,57)
REF_ORIGIN = (0.004501, -0.476622, -0.381965)
REF_EXTENT = (3.08, 0.76, 2.36)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} missing - run __graft_entry__.build() / make -C gpu-computing-course_b200/csrc")
        _LIB = C.CDLL(_PATH)
        _LIB.mg_grid_num_verts.restype = C.c_uint64
        _LIB.mg_grid_num_tris.restype = C.c_uint64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def soup(n, h=None, seed=1234, origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0), out=None):
    """C4: n private triangles, centroids uniform in the box, vertex jitter +-h.

    Default h keeps SURVEY §8(d)'s density: h = 0.68 * (box volume / n)^(1/3)."""
    if h is None:
        h = 0.68 * (extent[0] * extent[1] * extent[2] / n) ** (1.0 / 3.0)
    if out is None:
        xyz = np.empty((3 * n, 3), np.float32)
        idx = np.empty((n, 3), np.uint32)
    else:
        xyz, idx = out
    o = (C.c_double * 3)(*origin)
    e = (C.c_double * 3)(*extent)
    lib().mg_soup(C.c_uint32(n), C.c_double(h), C.c_uint64(seed), o, e, _p(xyz, C.c_float), _p(idx, C.c_uint32))
    return xyz, idx


def grid_sizes(nx, ny, sheets=1):
    """(nverts, ntris) of the grid workloads"""
    return (nx + 1) * (ny + 1) * sheets, 2 * nx * ny * sheets


def _grid_alloc(nx, ny, sheets=1, out=None):
    nv, nt = grid_sizes(nx, ny, sheets)
    if out is not None:  # caller-owned (e.g. pinned) buffers
        xyz, idx = out
        assert xyz.shape == (nv, 3) and xyz.dtype == np.float32 and idx.shape == (nt, 3) and idx.dtype == np.uint32
        return xyz, idx
    return np.empty((nv, 3), np.float32), np.empty((nt, 3), np.uint32)


def cloth_fold(nx=708, ny=708, layers=8, gap_edges=0.5, amp=1.5, wl_edges=12.0, seed=1, out=None):
    """C3: accordion-folded sheet inside the reference Morton box, dense contacts."""
    xyz, idx = _grid_alloc(nx, ny, out=out)
    lib().mg_cloth_fold(C.c_uint32(nx), C.c_uint32(ny), C.c_uint32(layers), C.c_double(gap_edges), C.c_double(amp),
                        C.c_double(wl_edges), C.c_uint64(seed), _p(xyz, C.c_float), _p(idx, C.c_uint32))
    return xyz, idx


def two_sheets(nq=4096, seed=7, out=None):
    """C5: two nq x nq-quad sheets intersecting along curves, unit cube."""
    xyz, idx = _grid_alloc(nq, nq, 2, out=out)
    lib().mg_two_sheets(C.c_uint32(nq), C.c_uint64(seed), _p(xyz, C.c_float), _p(idx, C.c_uint32))
    return xyz, idx


def flag(nx=795, nz=794, nfold=3, seed=2021, out=None):
    """C1/C2 stand-in for the missing flag-2000-changed.obj (1 262 460 triangles by default)."""
    xyz, idx = _grid_alloc(nx, nz, out=out)
    lib().mg_flag(C.c_uint32(nx), C.c_uint32(nz), C.c_uint32(nfold), C.c_uint64(seed), _p(xyz, C.c_float),
                  _p(idx, C.c_uint32))
    return xyz, idx


def write_obj(path, xyz, idx):
    xyz = np.ascontiguousarray(xyz, np.float32)
    idx = np.ascontiguousarray(idx, np.uint32)
    rc = lib().mg_write_obj(os.fsencode(path), _p(xyz, C.c_float), C.c_uint32(xyz.shape[0]), _p(idx, C.c_uint32),
                            C.c_uint32(idx.shape[0]))
    if rc != 0:
        raise OSError(f"mg_write_obj({path}) failed: {rc}")


def edge_cases(copies=10, seed=3, origin=REF_ORIGIN, extent=REF_EXTENT):
    """Pairs of private triangles in the configurations where the narrow phase (reference tri_contact.cuh:19-78) and
    the strict box test (box.cuh:40-43) are at their limits: piercing, coplanar overlap, exactly shared edges and
    vertices, a vertex lying exactly in the other triangle's plane, crossing edges, near misses tens of fp32 ulps apart,
    and degenerate triangles (points, segments, collinear). Every pattern is instantiated `copies` times, each copy at
    its own lattice cell of the box with its own random rotation (so boxes have volume; one copy per pattern stays
    axis-aligned, where flat boxes fail the STRICT overlap test and the reference reports nothing).
    Plain numpy, deterministic; fp32 coordinates; centroids all distinct (the reference needs unique Morton codes)."""
    rng = np.random.default_rng(seed)

    def rot(k):
        if k == 0:
            return np.eye(3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    A = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    patterns = [
        ("pierce", [[0.25, 0.25, -0.5], [0.25, 0.25, 0.5], [0.6, 0.1, 0.4]]),
        ("coplanar_overlap", [[0.2, 0.2, 0.0], [1.2, 0.3, 0.0], [0.3, 1.1, 0.0]]),
        ("coplanar_disjoint", [[1.5, 1.5, 0.0], [2.5, 1.5, 0.0], [1.5, 2.5, 0.0]]),
        ("shared_edge_coords", [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.5]]),
        ("shared_vertex_coords", [[1.0, 0.0, 0.0], [2.0, 0.5, 0.5], [1.5, -0.5, 0.7]]),
        ("vertex_in_plane", [[0.25, 0.25, 0.0], [0.5, 0.5, 1.0], [0.1, 0.6, 0.8]]),
        ("vertex_on_edge", [[0.5, 0.0, 0.0], [0.5, -1.0, 0.6], [0.9, -0.8, -0.4]]),
        ("edges_cross", [[0.5, -0.5, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 1.0]]),
        ("near_miss_above", [[0.15, 0.1, 2 ** -12], [0.7, 0.15, 2 ** -12], [0.2, 0.6, 2 ** -12]]),
        ("near_miss_side", [[1.0 + 2 ** -12, 0.0, -0.5], [1.0 + 2 ** -12, 1.0, 0.5], [1.6, 0.5, 0.0]]),
        ("graze_below", [[0.15, 0.1, -(2 ** -12)], [0.7, 0.15, 2 ** -12], [0.2, 0.6, -(2 ** -12)]]),
        ("point_inside", [[0.3, 0.3, 0.0]] * 3),
        ("point_outside", [[0.3, 0.3, 0.25]] * 3),
        ("segment_through", [[0.3, 0.3, -0.5], [0.3, 0.3, 0.5], [0.3, 0.3, -0.5]]),
        ("collinear_in_plane", [[-0.5, 0.4, 0.0], [0.5, 0.4, 0.0], [1.5, 0.4, 0.0]]),
        ("parallel_close", [[0.0, 0.0, 0.01], [1.0, 0.0, 0.01], [0.0, 1.0, 0.01]]),
        ("contained_coplanar", [[0.1, 0.1, 0.0], [0.4, 0.1, 0.0], [0.1, 0.4, 0.0]]),
    ]
    ncell = len(patterns) * copies
    side = int(np.ceil(ncell ** (1.0 / 3.0)))
    o, e = np.asarray(origin, np.float64), np.asarray(extent, np.float64)
    cell = 0.9 * e / side
    s = 0.18 * cell.min()  # pattern coordinates span about [-1, 2.5]: everything stays inside its cell
    verts = []
    k = 0
    for name, B in patterns:
        for c in range(copies):
            ijk = np.array([k % side, (k // side) % side, k // (side * side)], np.float64)
            centre = o + 0.05 * e + (ijk + 0.5) * cell + rng.uniform(-0.01, 0.01, 3) * cell
            R = rot(c)
            for tri in (A, np.asarray(B, np.float64)):
                verts.append((centre + s * (tri - 0.5) @ R.T))
            k += 1
    xyz = np.concatenate(verts).astype(np.float32)
    idx = np.arange(len(xyz), dtype=np.uint32).reshape(-1, 3)
    names = [n for n, _ in patterns for _ in range(copies)]
    return xyz, idx, names
