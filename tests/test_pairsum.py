"""CPU: the pair-list checksum (gpu-computing-course_b200/pairsum.py) that bench.py prints on every line and
tests/golden/checksums.json pins - known answers, wrap-around, and the file's schema."""
import importlib
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ps = importlib.import_module("gpu-computing-course_b200.pairsum")


def test_checksum_known_answers_and_wraparound():
    assert ps.pairs_checksum_np(np.empty((0, 2), np.uint32)) == [0, 0]
    p = np.array([[1, 2], [3, 0x80000001]], np.uint32)          # word = lower | higher << 32
    w = [1 | 2 << 32, 3 | 0x80000001 << 32]
    assert ps.pairs_checksum_np(p) == [sum(w) % 2 ** 64, sum(x * x + (x >> 7) for x in w) % 2 ** 64]
    big = np.full((1000, 2), 0xfffffffe, np.uint32)             # sums far beyond 2^64: must wrap, not overflow-error
    x = 0xfffffffe | 0xfffffffe << 32
    assert ps.pairs_checksum_np(big) == [(1000 * x) % 2 ** 64, (1000 * (x * x + (x >> 7))) % 2 ** 64]
    # order matters only through the multiset: the sums are order-independent, the COUNT is reported beside them
    q = p[::-1].copy()
    assert ps.pairs_checksum_np(q) == ps.pairs_checksum_np(p)


def test_golden_checksums_cover_every_bench_workload():
    import bench
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "checksums.json")))
    mg = importlib.import_module("gpu-computing-course_b200.meshgen")
    for name in bench.WORKLOADS:
        assert name in g, name
        nverts, ntris = bench.workload_sizes(mg, name)
        assert g[name]["triangles"] == ntris and g[name]["vertices"] == nverts
        assert g[name]["pairs"] > 0 and len(g[name]["checksum"]) == 2
        assert "reference host functions" in g[name]["source"] and "cd_oracle.c" in g[name]["source"]
        assert bench.matches_oracle(name, g[name]["pairs"], g[name]["checksum"]) is True
        assert bench.matches_oracle(name, g[name]["pairs"] + 1, g[name]["checksum"]) is False
    assert bench.matches_oracle("no-such-workload", 1, [0, 0]) is None
