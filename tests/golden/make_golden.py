"""Generate tests/golden/*.npz from the REFERENCE'S OWN host functions.

Run in the build container only (needs oracle/_ref/libref_cd.so, which is compiled from
/root/reference/CollisionDetection by oracle/Makefile; the GPU box never runs this script):

    python tests/golden/make_golden.py

Every array below is produced by reference code (loadObj / morton3D / thrust::sort_by_key /
generateHierarchyParallelCpu / calBoundingBoxCpu / findCollisionIterativeCpu /
checkTriangleContact / checkBoxOverlap / determineRangeCpu / findSplitCpu) behind
oracle/ref_driver.cu. The meshes come from our own generators (the reference's bundled flag mesh
is missing from the checkout, .MISSING_LARGE_BLOBS:1) and all lie inside the reference's
hard-coded Morton box (morton.h:43-58), where its keys are valid.
"""
import importlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refcd  # noqa: E402

mg = importlib.import_module("gpu-computing-course_b200.meshgen")


def pipeline_fixture(name, xyz, idx, via_obj):
    if via_obj:  # through the reference's OBJ parser (load_obj.h:24-123)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, name + ".obj")
            mg.write_obj(path, xyz, idx)
            m = refcd.RefMesh.from_obj(path)
        pxyz, pidx = m.mesh()
        # the parser must give back exactly the arrays that were written ("%.9g" round-trips fp32)
        assert np.array_equal(pxyz, xyz) and np.array_equal(pidx, idx), "OBJ round trip changed the mesh"
    else:
        m = refcd.RefMesh.from_arrays(xyz, idx)
    skeys, sids = m.sorted()
    assert np.all(skeys[1:] > skeys[:-1]), "reference requires strictly increasing codes (load_obj.h:109-115)"
    wrong_parent = m.build()
    nd = m.nodes()
    pairs = m.collide()
    order = np.lexsort((pairs[:, 1], pairs[:, 0]))
    pairs = pairs[order]
    b32 = nd["bounds"].astype(np.float32)
    assert np.array_equal(b32.astype(np.float64), nd["bounds"]), "bounds are not fp32-exact"
    assert wrong_parent == 0 and np.all(nd["bounded"] == 2)
    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(out, xyz=xyz, idx=idx, sorted_keys=skeys, sorted_ids=sids, left=nd["left"],
                        right=nd["right"], parent=nd["parent"], bounds=b32, pairs=pairs,
                        via_obj=np.array(int(via_obj)))
    print(f"{name}: {len(idx)} tris, {len(pairs)} pairs -> {os.path.getsize(out)} bytes")
    m.close()


def kat_fixture():
    rng = np.random.default_rng(20261018)
    # check.cuh:19-27 known-answer input, every internal node
    keys = np.array([1, 2, 4, 5, 19, 24, 25, 30], np.uint64)
    rs = np.array([refcd.range_split(keys, i) for i in range(len(keys) - 1)], np.int32)
    # morton3D (morton.h:70-89) on points inside the reference box
    o = np.array(mg.REF_ORIGIN)
    e = np.array(mg.REF_EXTENT)
    pts = o + e * rng.uniform(1e-6, 1.0, size=(256, 3))
    pts[0] = o + e * 1e-9
    pts[1] = o + e * 1.0
    codes = np.array([refcd.morton3D(*p) for p in pts], np.uint64)
    # checkTriangleContact (tri_contact.cuh:19-78): random, near-touching, coplanar and degenerate pairs
    tris = rng.uniform(-1, 1, size=(600, 18))
    tris[100:200, 9:] = tris[100:200, :9] + rng.uniform(-0.3, 0.3, size=(100, 9))       # close pairs
    tris[200:300, [2, 5, 8, 11, 14, 17]] = 0.0                                             # coplanar
    tris[300:350, 9:12] = tris[300:350, 0:3]                                               # shared vertex position
    tris[350:400, 3:6] = tris[350:400, 0:3]                                                # degenerate P
    tris[400:450, 9:] = tris[400:450, :9]                                                  # identical triangles
    tris[450:500] = np.round(tris[450:500] * 4) / 4                                        # lattice: exact touching
    tris = tris.astype(np.float32).astype(np.float64)  # the pipeline only ever sees fp32-exact values
    contact = np.array([refcd.tri_contact(t) for t in tris], np.int32)
    # checkBoxOverlap (box.cuh:40-43): strict; touching faces do not overlap
    boxes = np.round(rng.uniform(0, 4, size=(400, 2, 2, 3)))  # [pair][box][corner][axis] on a lattice -> many ties
    boxes.sort(axis=2)
    boxes = boxes.reshape(400, 2, 6)
    overlap = np.array([refcd.box_overlap(b[0], b[1]) for b in boxes], np.int32)
    out = os.path.join(HERE, "kat.npz")
    np.savez_compressed(out, range_keys=keys, range_split=rs, morton_pts=pts, morton_codes=codes, tris=tris,
                        contact=contact, boxes=boxes, overlap=overlap)
    print(f"kat: contact {contact.sum()}/{len(contact)}, overlap {overlap.sum()}/{len(overlap)} -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    assert refcd.available(), "build oracle/_ref first (make -C oracle)"
    kat_fixture()
    pipeline_fixture("flag_40x40", *mg.flag(40, 40), via_obj=True)
    pipeline_fixture("cloth_20x20", *mg.cloth_fold(20, 20), via_obj=True)
    pipeline_fixture("soup_1500_refbox", *mg.soup(1500, seed=42, origin=(0.1, -0.4, -0.3), extent=(2.8, 0.6, 2.2)),
                     via_obj=False)
    # the multi-GPU workload's generator (two intersecting sheets), scaled uniformly into the reference box
    sx, si = mg.two_sheets(24, seed=7)
    o, e = np.array(mg.REF_ORIGIN), np.array(mg.REF_EXTENT)
    sx = (o + 0.05 * e + sx.astype(np.float64) * (0.9 * e.min())).astype(np.float32)
    pipeline_fixture("two_sheets_24", sx, si, via_obj=True)
    # narrow-phase / strict-box limit cases: exact touching, coplanar, near misses, degenerate triangles
    pipeline_fixture("edge_cases", *mg.edge_cases()[:2], via_obj=False)
