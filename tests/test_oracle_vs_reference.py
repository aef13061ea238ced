"""CPU: the oracle restatement against the reference's own functions, run live.

Needs oracle/_ref/libref_cd.so (oracle/ref_driver.cu compiled against the headers under
/root/reference/CollisionDetection). It is built in the build container and travels to the GPU
box as a binary; if it is absent these tests skip and tests/test_oracle_golden.py (committed
vectors from the same functions) carries the pin.
"""
import os

import numpy as np
import pytest

from oracle import refcd

REFERENCE_SOURCES = "/root/reference/CollisionDetection"

if not refcd.available() and os.path.isdir(REFERENCE_SOURCES):
    # the build container: the reference is there, so its functions MUST have been compiled (make -C oracle);
    # a silent skip here would un-pin the oracle without anybody noticing
    raise RuntimeError("oracle/_ref/libref_cd.so is missing although /root/reference exists: run `make -C oracle` "
                       "(or __graft_entry__.build())")

# elsewhere (the GPU box normally receives the built file; a bare checkout does not): skip LOUDLY - the committed
# vectors made by the same functions (tests/test_oracle_golden.py) then carry the pin
pytestmark = pytest.mark.skipif(not refcd.available(),
                                reason="PARITY PIN REDUCED: oracle/_ref/libref_cd.so absent and no /root/reference to build it "
                                       "from - only the committed golden vectors check the oracle here")


def compare_all_stages(co, ref_mesh, xyz, idx):
    op = co.default_params()
    rk, ri = ref_mesh.sorted()
    sk, si = co.sort_keys(co.morton_keys(xyz, idx, op))
    assert np.array_equal(sk, rk), "sorted keys"
    assert np.array_equal(si, ri), "sorted ids"
    assert ref_mesh.build() == 0  # parentWrongNum, bvh.cuh:192-194
    nd = ref_mesh.nodes()
    h = co.hierarchy(sk)
    assert np.array_equal(h["left"], nd["left"]) and np.array_equal(h["right"], nd["right"]), "children"
    assert np.array_equal(h["parent"], nd["parent"]), "parents"
    assert np.all(nd["bounded"] == 2)  # check.cuh:73
    b = co.refit(xyz, idx, si, h)
    assert np.array_equal(b, nd["bounds"]), "bounds"
    rp = co.sort_pairs(ref_mesh.collide())
    p, ctr = co.self_collide(xyz, idx, si, h, b)
    assert np.array_equal(co.sort_pairs(p), rp), "pair set"
    assert np.all(rp[:, 0] < rp[:, 1])  # tri_contact.cuh:81
    return rp, ctr


def test_soup_in_reference_box(co, mg):
    xyz, idx = mg.soup(30000, seed=5, origin=(0.1, -0.4, -0.3), extent=(2.8, 0.6, 2.2))
    m = refcd.RefMesh.from_arrays(xyz, idx)
    rp, ctr = compare_all_stages(co, m, xyz, idx)
    assert len(rp) > 3000
    assert ctr.max_stack <= 32  # collision.cuh:21: the reference's fixed stack is enough here
    m.close()


def test_cloth_dense_contacts(co, mg):
    xyz, idx = mg.cloth_fold(120, 120)
    m = refcd.RefMesh.from_arrays(xyz, idx)
    rp, _ = compare_all_stages(co, m, xyz, idx)
    assert len(rp) > 0.1 * len(idx)
    m.close()


def test_edge_cases_touching_coplanar_degenerate(co, mg):
    """exact touching, coplanar, near-miss and degenerate triangle pairs (meshgen.edge_cases): the restatement follows
    the reference through every borderline compare of project3 / project6 and the strict box test"""
    xyz, idx, names = mg.edge_cases()
    m = refcd.RefMesh.from_arrays(xyz, idx)
    rp, _ = compare_all_stages(co, m, xyz, idx)
    hit = set(map(tuple, rp.tolist()))
    per = {}
    for c, n in enumerate(names):
        per.setdefault(n, []).append((2 * c, 2 * c + 1) in hit)
    assert all(per["pierce"]) and all(per["segment_through"]) and all(per["graze_below"])
    assert not any(per["coplanar_disjoint"]) and not any(per["parallel_close"]) and not any(per["point_outside"])
    assert not per["coplanar_overlap"][0]  # the axis-aligned copy: flat boxes fail the STRICT overlap test (box.cuh:40-43)
    assert any(per["coplanar_overlap"][1:])
    m.close()


def test_flag_standin_through_obj_parser(co, mg, tmp_path, capfd):
    xyz, idx = mg.flag(100, 100)
    path = os.path.join(tmp_path, "flag.obj")
    mg.write_obj(path, xyz, idx)
    m = refcd.RefMesh.from_obj(path)  # loadObj, load_obj.h:24
    capfd.readouterr()  # loadObj prints statistics (load_obj.h:117-122)
    pxyz, pidx = m.mesh()
    assert np.array_equal(pxyz, xyz) and np.array_equal(pidx, idx)
    rp, _ = compare_all_stages(co, m, xyz, idx)
    assert 0 < len(rp) < 1000
    m.close()


def test_predicates_random(co):
    rng = np.random.default_rng(7)
    o = np.array([0.004501, -0.476622, -0.381965])
    e = np.array([3.08, 0.76, 2.36])
    op = co.default_params()
    for p in o + e * rng.uniform(0.001, 0.999, size=(2000, 3)):
        assert co.morton_of_centroid(*p, op) == refcd.morton3D(*p)
    tris = rng.uniform(-1, 1, size=(3000, 18)).astype(np.float32).astype(np.float64)
    tris[1000:2000, 9:] = tris[1000:2000, :9] + rng.uniform(-0.2, 0.2, size=(1000, 9)).astype(np.float32)
    agree = [co.tri_contact(t) == refcd.tri_contact(t) for t in tris]
    assert all(agree)


def test_range_split_random_keys(co):
    rng = np.random.default_rng(11)
    keys = np.unique(rng.integers(0, 1 << 60, size=500, dtype=np.uint64))
    h = co.hierarchy(keys)
    for i in range(0, len(keys) - 1, 7):
        assert refcd.range_split(keys, i) == (h["first"][i], h["last"][i], h["split"][i])
