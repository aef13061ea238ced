"""The arithmetic of the quantised traversal nodes (csrc/common.cuh: qcell_lo / qcell_hi / qnode_word / qquery_word /
qoverlap), restated in numpy float32 / uint32 (the library is built with -fmad=false, so (x - o) * s rounds twice like
numpy does). CPU only: the kernels themselves are checked against the oracle in tests/test_gpu_parity.py.

What the traversal relies on:
  1. the packed test - ONE 32-bit subtraction per axis, both strict comparisons in its two guard bits - equals the
     plain comparison of the cells;
  2. the cells are conservative: boxes that strictly overlap in floats (box.cuh:40-43) always overlap on the grid, for
     any frame, including values outside it (clamped) and a degenerate frame (scale 0).
"""
import numpy as np

Q = np.uint32(32767)
GUARD = np.uint32(0x80008000)
f32 = np.float32


def cell_lo(x, o, s):
    u = np.floor((x.astype(f32) - f32(o)) * f32(s))
    return np.minimum(np.maximum(u, f32(0)), f32(Q - 1)).astype(np.int64).astype(np.uint32)


def cell_hi(x, o, s):
    u = np.floor((x.astype(f32) - f32(o)) * f32(s)) + f32(1)
    return np.minimum(np.maximum(u, f32(1)), f32(Q)).astype(np.int64).astype(np.uint32)


def node_word(lo, hi, o, s):
    return cell_lo(lo, o, s) | ((Q - cell_hi(hi, o, s)) << np.uint32(16))


def query_word(lo, hi, o, s):
    return ((cell_hi(hi, o, s) | np.uint32(0x8000)) | (((Q - cell_lo(lo, o, s)) | np.uint32(0x8000)) << np.uint32(16))) - np.uint32(0x00010001)


def packed_overlap(qw, nw):
    with np.errstate(over="ignore"):
        return ((qw - nw) & GUARD) == GUARD


def test_packed_subtraction_is_the_two_strict_cell_comparisons():
    rng = np.random.default_rng(1)
    n = 200000
    # all cell values incl. the extremes, as raw cells (bypassing the float map)
    nlo = rng.integers(0, 32767, n, dtype=np.uint32)          # [0, Q-1]
    nhi = rng.integers(1, 32768, n, dtype=np.uint32)          # [1, Q]
    qlo = rng.integers(0, 32767, n, dtype=np.uint32)
    qhi = rng.integers(1, 32768, n, dtype=np.uint32)
    for arr in (nlo, qlo):
        arr[:4] = (0, 0, 32766, 32766)
    for arr in (nhi, qhi):
        arr[:4] = (1, 32767, 1, 32767)
    nw = nlo | ((Q - nhi) << np.uint32(16))
    qw = ((qhi | np.uint32(0x8000)) | (((Q - qlo) | np.uint32(0x8000)) << np.uint32(16))) - np.uint32(0x00010001)
    assert np.all((nw & GUARD) == 0)                          # node words keep both guard bits clear
    want = (nlo < qhi) & (qlo < nhi)
    assert np.array_equal(packed_overlap(qw, nw), want)
    # and the AND over three axes is the AND of the three answers
    with np.errstate(over="ignore"):
        t = (qw - nw) & (np.roll(qw, 1) - np.roll(nw, 1)) & (np.roll(qw, 2) - np.roll(nw, 2))
    assert np.array_equal((t & GUARD) == GUARD, want & np.roll(want, 1) & np.roll(want, 2))


def _frames():
    yield 0.0, 32767.0                      # unit cube
    yield 0.004501, 32767.0 / 3.08          # the reference's x axis (morton.h:45)
    yield -0.476622, 32767.0 / 0.76         # ... y axis
    yield 0.4, 32767.0 / 0.2                # most values outside the frame: clamped
    yield -40.0, 32767.0 / 100.0            # cells far larger than the boxes
    yield 0.0, 0.0                          # degenerate frame: one cell


def test_cells_are_conservative_for_every_frame():
    rng = np.random.default_rng(2)
    n = 400000
    for o, s in _frames():
        a_lo = rng.random(n, dtype=f32)
        b_hi = rng.random(n, dtype=f32)
        # near-ties and exact ties are where a quantiser goes wrong
        b_hi[: n // 4] = np.nextafter(a_lo[: n // 4], f32(2))          # b.hi one ulp above a.lo: overlaps
        b_hi[n // 4: n // 2] = a_lo[n // 4: n // 2]                     # touching: no overlap in floats (either answer is allowed)
        overlap = a_lo < b_hi
        on_grid = cell_lo(a_lo, o, s) < cell_hi(b_hi, o, s)
        assert np.all(on_grid[overlap]), (o, s)
        # whole boxes through the packed words
        lo = rng.random((n, 2), dtype=f32)
        ext = rng.random((n, 2), dtype=f32) * f32(0.01)
        hi = lo + ext
        exact = (lo[:, 0] < hi[:, 1]) & (lo[:, 1] < hi[:, 0])
        got = packed_overlap(query_word(lo[:, 0], hi[:, 0], o, s), node_word(lo[:, 1], hi[:, 1], o, s))
        assert np.all(got[exact]), (o, s)
        if s == 32767.0:   # a fitted frame admits few extras: the quantised test is not vacuous
            assert got.sum() < 1.2 * exact.sum() + 100


def test_cells_are_monotone():
    rng = np.random.default_rng(3)
    x = np.sort(rng.normal(0.5, 1.0, 200000).astype(f32))
    for o, s in _frames():
        assert np.all(np.diff(cell_lo(x, o, s).astype(np.int64)) >= 0)
        assert np.all(np.diff(cell_hi(x, o, s).astype(np.int64)) >= 0)
        assert np.all(cell_lo(x, o, s) < cell_hi(x, o, s))
