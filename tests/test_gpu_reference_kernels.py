"""GPU: the REFERENCE'S OWN CUDA kernels (fillLeafNodes, generateHierarchyParallel, calBoundingBox, findCollisions -
bvh.cuh:125,146,258, collision.cuh:73), built for sm_100a from /root/reference into oracle/_ref/libref_cd.so and
launched with main.cu:92-142's grids, give on this B200 the pair set their CPU twins gave in the build container
(tests/golden/*.npz) - and the set our library gives. This is the same-box GPU baseline bench.py prints for flag1m."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["flag_40x40", "cloth_20x20", "soup_1500_refbox"])
def test_reference_gpu_kernels_reproduce_the_golden_pairs(cd, co, ctx, name):
    from oracle import refcd
    if not refcd.available():
        pytest.skip("oracle/_ref was not built (needs /root/reference in the build container)")
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    m = refcd.RefMesh.from_arrays(g["xyz"], g["idx"])
    pairs, ms = m.gpu_run(repeats=2)
    m.close()
    assert set(ms) == {"fillLeafNodes", "generateHierarchyParallel", "calBoundingBox", "findCollisions"} and all(v > 0 for v in ms.values())
    assert np.array_equal(co.sort_pairs(pairs.copy()), g["pairs"])
    mesh = ctx.mesh_from_arrays(g["xyz"], g["idx"])
    bvh = ctx.bvh_build(mesh, cd.default_params())
    assert np.array_equal(ctx.self_collide(bvh, sorted=True), g["pairs"])
    bvh.destroy()
    mesh.destroy()
