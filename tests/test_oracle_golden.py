"""CPU: the oracle (oracle/cd_oracle.c) against the committed golden vectors.

tests/golden/*.npz were produced by the reference's OWN host functions (see
tests/golden/make_golden.py); this is what pins the oracle on machines without /root/reference.
"""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIPELINES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not p.endswith("kat.npz"))


def test_fixtures_present():
    assert PIPELINES == ["cloth_20x20", "edge_cases", "flag_40x40", "soup_1500_refbox", "two_sheets_24"]


@pytest.mark.parametrize("name", PIPELINES)
def test_oracle_pipeline_matches_reference_output(co, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    xyz, idx = g["xyz"], g["idx"]
    n = len(idx)
    op = co.default_params()  # the reference's hard-coded box, morton.h:45,51,57
    keys = co.morton_keys(xyz, idx, op)
    sk, si = co.sort_keys(keys)
    assert np.array_equal(sk, g["sorted_keys"])
    assert np.array_equal(si, g["sorted_ids"])
    h = co.hierarchy(sk)
    assert np.array_equal(h["left"], g["left"])
    assert np.array_equal(h["right"], g["right"])
    assert np.array_equal(h["parent"], g["parent"])
    b = co.refit(xyz, idx, si, h)
    assert np.array_equal(b, g["bounds"].astype(np.float64))
    pairs, ctr = co.self_collide(xyz, idx, si, h, b)
    assert np.array_equal(co.sort_pairs(pairs), g["pairs"])
    # whole-pipeline entry point (what bench.py times) gives the same list
    p2, tm = co.run(xyz, idx, op)
    assert np.array_equal(p2, g["pairs"]) and tm.pairs == len(g["pairs"])
    # and the set definition itself (brute force, check.cuh:117-141 + box filter)
    assert np.array_equal(co.brute_force(xyz, idx), g["pairs"])
    assert n == len(g["sorted_ids"])


def test_kat_range_split_check_cuh(co):
    """check.cuh:19-27: keys {1,2,4,5,19,24,25,30}; answers from determineRangeCpu/findSplitCpu"""
    g = np.load(os.path.join(GOLDEN, "kat.npz"))
    h = co.hierarchy(g["range_keys"])
    got = np.stack([h["first"], h["last"], h["split"]], axis=1)
    assert np.array_equal(got, g["range_split"])
    assert tuple(got[6]) == (5, 6, 5)  # the node testFunc probes
    assert tuple(got[0]) == (0, 7, 3)


def test_kat_morton3d(co):
    g = np.load(os.path.join(GOLDEN, "kat.npz"))
    op = co.default_params()
    got = np.array([co.morton_of_centroid(*p, op) for p in g["morton_pts"]], np.uint64)
    assert np.array_equal(got, g["morton_codes"])


def test_kat_tri_contact(co):
    g = np.load(os.path.join(GOLDEN, "kat.npz"))
    got = np.array([co.tri_contact(t) for t in g["tris"]], np.int32)
    assert np.array_equal(got, g["contact"])
    assert 0 < got.sum() < len(got)


def test_kat_box_overlap_is_strict(co):
    g = np.load(os.path.join(GOLDEN, "kat.npz"))
    got = np.array([co.box_overlap(b[0], b[1]) for b in g["boxes"]], np.int32)
    assert np.array_equal(got, g["overlap"])
    a = np.array([0, 0, 0, 1, 1, 1.0])
    assert co.box_overlap(a, a + np.array([1, 0, 0, 1, 0, 0.0])) == 0  # touching faces: product is 0, not > 0
    assert co.box_overlap(a, a + np.array([0.5, 0, 0, 0.5, 0, 0.0])) == 1


def test_oracle_edge_cases(co, mg):
    op = co.make_params((0, 0, 0), (1, 1, 1))
    # n = 1 and n = 2: no internal node / a single one
    for n in (1, 2, 3):
        xyz, idx = mg.soup(n, h=0.4, seed=n)
        p, tm = co.run(xyz, idx, op)
        assert np.array_equal(p, co.brute_force(xyz, idx))
    # duplicate keys (30-bit, tight cluster): tie-break keeps the tree well formed, set unchanged
    xyz, idx = mg.soup(3000, h=0.08, seed=4)
    op30 = co.make_params((0, 0, 0), (4096, 4096, 4096), 30)  # every centroid lands in Morton cell 0
    keys = co.morton_keys(xyz, idx, op30)
    assert len(np.unique(keys)) == 1
    p, _ = co.run(xyz, idx, op30)
    assert np.array_equal(p, co.brute_force(xyz, idx))
    sk, si = co.sort_keys(keys)
    assert np.array_equal(si, np.arange(3000, dtype=np.uint32))  # stable: ties keep face order
