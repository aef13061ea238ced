"""Checksum of a SORTED colliding-pair list, used to compare full-size results
(16-64 M triangles) without shipping the lists: bench.py prints it on every line, tests/golden/checksums.json
holds the values the CPU oracle and the reference's own host functions produce (tests/golden/make_checksums.py).

A pair is uint32[2] = (lower ID, higher ID) (reference main.cu:145-152 prints them in that order); packed
little-endian into one 64-bit word w = lower | higher << 32. The checksum is [sum(w), sum(w*w + (w >> 7))],
both modulo 2^64, together with the pair count. bench.py computes the same two sums on the device."""
import numpy as np

MASK = (1 << 64) - 1


def pairs_checksum_np(pairs):
    """pairs: (count, 2) uint32 array, lower ID first -> [sum1, sum2] as python ints"""
    p = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    if p.shape[0] == 0:
        return [0, 0]
    w = p.view(np.uint64).reshape(-1)
    with np.errstate(over="ignore"):
        s1 = int(w.sum(dtype=np.uint64))
        s2 = int((w * w + (w >> np.uint64(7))).sum(dtype=np.uint64))
    return [s1 & MASK, s2 & MASK]
