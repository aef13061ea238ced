"""GPU parity: every stage of the CUDA path, called through the C ABI, against the CPU oracle.

Bit-exact bar (BASELINE.json north_star): keys, sorted order, tree topology, node bounds
(0 ulp — they are min/max of fp32 inputs) and the sorted colliding-pair set.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

UNIT = dict(origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0))


def oracle_stages(co, xyz, idx, op):
    keys = co.morton_keys(xyz, idx, op)
    sk, si = co.sort_keys(keys)
    h = co.hierarchy(sk)
    b = co.refit(xyz, idx, si, h)
    pairs, ctr = co.self_collide(xyz, idx, si, h, b)
    return dict(keys=keys, sk=sk, si=si, h=h, bounds=b, pairs=co.sort_pairs(pairs), ctr=ctr)


def check_all_stages(cd, co, ctx, xyz, idx, origin=None, extent=None, key_bits=63, auto_box=False):
    if auto_box:
        op = co.auto_params(xyz, key_bits)
    elif origin is None:
        op = co.default_params(key_bits)
    else:
        op = co.make_params(origin, extent, key_bits)
    ref = oracle_stages(co, xyz, idx, op)
    gp = cd.make_params(origin, extent, key_bits, auto_box)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, gp)
    nodes, skeys, sids = bvh.download()
    n = len(idx)
    assert np.array_equal(skeys, ref["sk"]), "sorted Morton keys differ"
    assert np.array_equal(sids, ref["si"]), "sorted triangle ids differ"
    if n >= 2:
        assert np.array_equal(nodes["left"][: n - 1], ref["h"]["left"]), "left children differ"
        assert np.array_equal(nodes["right"][: n - 1], ref["h"]["right"]), "right children differ"
    gb = np.concatenate([nodes["lo"], nodes["hi"]], axis=1).astype(np.float64)
    assert np.array_equal(gb, ref["bounds"]), "node bounds differ (expected 0 ulp)"
    chk = bvh.validate(mesh)
    expect_unsorted = int(np.sum(ref["sk"][1:] <= ref["sk"][:-1])) if n > 1 else 0
    assert chk == dict(null_parent_internal=1 if n >= 2 else 0, wrong_bound_count=0, null_child=0,
                       uninit_box_internal=0, null_parent_leaf=0, bad_triangle=0, uninit_box_leaf=0,
                       unsorted_keys=expect_unsorted, box_not_enclosing=0), chk
    pairs = ctx.self_collide(bvh, sorted=True)
    assert pairs.shape == ref["pairs"].shape, (pairs.shape, ref["pairs"].shape)
    assert np.array_equal(pairs, ref["pairs"]), "colliding-pair set differs"
    unsorted_pairs = ctx.self_collide(bvh, sorted=False)
    assert np.array_equal(co.sort_pairs(unsorted_pairs.copy()), ref["pairs"])
    st = ctx.stats()
    assert st["pairs"] == len(pairs)
    bvh.destroy()
    mesh.destroy()
    return ref, st


def test_soup_small_unit_cube(cd, co, ctx, mg):
    xyz, idx = mg.soup(20000, seed=3)
    ref, st = check_all_stages(cd, co, ctx, xyz, idx, **UNIT)
    assert len(ref["pairs"]) > 1000


def test_soup_reference_box_default_params(cd, co, ctx, mg):
    xyz, idx = mg.soup(50000, seed=11, origin=(0.1, -0.4, -0.3), extent=(2.8, 0.6, 2.2))
    check_all_stages(cd, co, ctx, xyz, idx)


def test_soup_256k(cd, co, ctx, mg):
    xyz, idx = mg.soup(1 << 18, seed=1234)
    ref, st = check_all_stages(cd, co, ctx, xyz, idx, **UNIT)
    assert st["candidates"] >= st["pairs"]


def test_cloth_dense_contacts_shared_vertices(cd, co, ctx, mg):
    xyz, idx = mg.cloth_fold(200, 200)
    ref, _ = check_all_stages(cd, co, ctx, xyz, idx)
    assert len(ref["pairs"]) > 0.1 * len(idx)


def test_flag_standin(cd, co, ctx, mg):
    xyz, idx = mg.flag(300, 300)
    ref, _ = check_all_stages(cd, co, ctx, xyz, idx)
    assert 0 < len(ref["pairs"]) < 2000


def test_two_sheets(cd, co, ctx, mg):
    xyz, idx = mg.two_sheets(200)
    check_all_stages(cd, co, ctx, xyz, idx, **UNIT)


def test_auto_box(cd, co, ctx, mg):
    xyz, idx = mg.soup(30000, seed=5, origin=(-7.0, 3.0, 100.0), extent=(2.0, 9.0, 0.5))
    check_all_stages(cd, co, ctx, xyz, idx, auto_box=True)


def test_30bit_keys_with_duplicates(cd, co, ctx, mg):
    # 60000 triangles in 2^30 cells still collide now and then; a tight cluster forces many ties
    xyz, idx = mg.soup(60000, seed=8)
    xyz2, idx2 = mg.soup(4000, h=1e-4, seed=9, origin=(0.5, 0.5, 0.5), extent=(0.002, 0.002, 0.002))
    xyz = np.concatenate([xyz, xyz2])
    idx = np.concatenate([idx, idx2 + 3 * 60000]).astype(np.uint32)
    ref, _ = check_all_stages(cd, co, ctx, xyz, idx, key_bits=30, **UNIT)
    assert len(np.unique(ref["keys"])) < len(idx), "test needs duplicate keys"


@pytest.mark.parametrize("scale", [1.0, 1 / 128, 1 / 180, 1 / 4096])
def test_hybrid_sort_runs_fixup_and_fallback(cd, co, ctx, mg, scale):
    """63-bit keys are sorted by radix passes over their top digits only; runs of equal high bits are put in order
    by the fix-up kernel, too long runs by the conditional full passes, and the number of digits adapts from one
    build to the next (api.cu run_build). Shrinking the mesh inside the Morton box makes the runs longer:
    1.0 -> runs of 1-2, 1/128 and 1/180 -> runs of 5-30 (insertion sort; one more digit next time),
    1/4096 -> every key in a handful of runs (fallback, then plain full sorts). Every build must give exactly the
    host sort's order (load_obj.h:107: ascending keys, ties in face order)."""
    xyz, idx = mg.soup(150_000, seed=5)
    xyz = (xyz * np.float32(scale) + np.float32(0.25)).astype(np.float32)
    op, gp = co.make_params(**UNIT), cd.make_params(**UNIT)
    rk, ri = co.sort_keys(co.morton_keys(xyz, idx, op))
    want, _ = co.run(xyz, idx, op)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, gp)
    seen = []
    for it in range(5):
        if it:
            ctx.bvh_rebuild(bvh, mesh, gp)
        _, sk, si = bvh.download()
        seen.append(ctx.stats()["sort_passes"])
        assert np.array_equal(sk, rk), f"build {it}: sorted keys differ (passes {seen})"
        assert np.array_equal(si, ri), f"build {it}: tie order differs (passes {seen})"
    assert np.array_equal(ctx.self_collide(bvh), want)
    for _ in range(3):  # back-to-back rebuilds with no synchronisation in between (the run statistics are not ready yet)
        ctx.bvh_rebuild(bvh, mesh, gp)
    _, sk, si = bvh.download()
    assert np.array_equal(sk, rk) and np.array_equal(si, ri)
    assert seen[0] == 5 and all(4 <= p <= 8 for p in seen), seen
    if scale == 1.0:
        assert seen[-1] == 4, seen      # the prefix got shorter
    if scale <= 1 / 4096:
        assert seen[-1] == 8, seen      # fell back to plain full sorts
    bvh.destroy()
    mesh.destroy()


def test_narrow_phase_edge_cases(cd, co, ctx, mg):
    """exact touching, coplanar, near-miss and degenerate triangle pairs: bit-exact pair set (and every build stage)"""
    xyz, idx, _ = mg.edge_cases(copies=24, seed=11)
    ref, _ = check_all_stages(cd, co, ctx, xyz, idx)
    assert 50 < len(ref["pairs"]) < len(idx)
    xyz, idx, _ = mg.edge_cases(copies=6, seed=5, origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0))
    check_all_stages(cd, co, ctx, xyz, idx, **UNIT)


def test_hybrid_sort_window_follows_growing_keys(cd, co, ctx, mg):
    """the window of sorted digits ends at the highest key bit the PREVIOUS build saw; when the mesh then grows and its
    keys reach above the window, that build falls back to the full passes and the next one moves the window up"""
    xyz, idx = mg.soup(120_000, seed=9)
    small = (xyz * np.float32(1 / 64)).astype(np.float32)   # keys below 2^42
    op, gp = co.make_params(**UNIT), cd.make_params(**UNIT)
    mesh = ctx.mesh_from_arrays(small, idx)
    bvh = ctx.bvh_build(mesh, gp)
    for frame in (small, small, xyz, xyz, small, xyz):
        mesh.update(xyz=frame)
        ctx.bvh_rebuild(bvh, mesh, gp)
        _, sk, si = bvh.download()
        rk, ri = co.sort_keys(co.morton_keys(frame, idx, op))
        assert np.array_equal(sk, rk) and np.array_equal(si, ri)
    assert np.array_equal(ctx.self_collide(bvh), co.run(xyz, idx, op)[0])
    bvh.destroy()
    mesh.destroy()


def test_all_keys_identical(cd, co, ctx, mg):
    # every centroid in one Morton cell: the tree is decided by the index tie-break alone
    xyz, idx = mg.soup(3000, h=0.08, seed=4)
    ref, _ = check_all_stages(cd, co, ctx, xyz, idx, key_bits=30, origin=(0, 0, 0), extent=(4096, 4096, 4096))
    assert len(np.unique(ref["keys"])) == 1


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 4095, 4096, 4097])
def test_tiny_and_tile_boundary_sizes(cd, co, ctx, mg, n):
    xyz, idx = mg.soup(n, h=0.3, seed=100 + n)
    check_all_stages(cd, co, ctx, xyz, idx, **UNIT)


def test_empty_mesh(cd, ctx):
    mesh = ctx.mesh_from_arrays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
    bvh = ctx.bvh_build(mesh, cd.default_params())
    assert len(ctx.self_collide(bvh)) == 0
    bvh.destroy()
    mesh.destroy()


def test_brute_force_small(cd, co, ctx, mg):
    xyz, idx = mg.soup(1500, h=0.08, seed=77)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    pairs = ctx.self_collide(bvh)
    assert np.array_equal(pairs, co.brute_force(xyz, idx))
    assert len(pairs) > 100


def test_sharded_query_union_equals_full(cd, co, ctx, mg):
    """each shard reports exactly the pairs whose smaller sorted position it owns (multigpu.owner_of_pairs)"""
    import importlib
    mgpu = importlib.import_module("gpu-computing-course_b200.multigpu")
    xyz, idx = mg.soup(40000, seed=21)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    _, _, sids = bvh.download(nodes=False)
    full = ctx.self_collide(bvh, sorted=True)
    for nshards, chunk in ((2, 0), (3, 1000), (8, 256), (4, 1 << 14)):
        owner = mgpu.owner_of_pairs(full, sids, nshards, chunk)
        parts = []
        for s in range(nshards):
            part = ctx.self_collide(bvh, sorted=True, shard=s, nshards=nshards, chunk=chunk)
            assert np.array_equal(part, full[owner == s]), (nshards, chunk, s)
            parts.append(part)
        assert np.array_equal(co.sort_pairs(np.concatenate(parts)), full)
    with pytest.raises(cd.B200cdError):
        ctx.self_collide(bvh, shard=2, nshards=2)
    bvh.destroy()
    mesh.destroy()


def test_rebuild_and_refit_after_vertex_update(cd, co, ctx, mg):
    xyz, idx = mg.cloth_fold(150, 150)
    p = cd.default_params()
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, p)
    first = ctx.self_collide(bvh)
    # move the vertices a little: refit keeps topology/order, boxes and pair set must follow the new positions
    rng = np.random.default_rng(3)
    xyz2 = (xyz + rng.normal(0, 2e-4, xyz.shape)).astype(np.float32)
    mesh.update(xyz=xyz2)
    ctx.bvh_refit(bvh, mesh)
    chk = bvh.validate(mesh)
    assert chk["box_not_enclosing"] == 0 and chk["wrong_bound_count"] == 0
    refit_pairs = ctx.self_collide(bvh)
    op = co.default_params()
    ref2, _ = co.run(xyz2, idx, op)
    assert np.array_equal(refit_pairs, ref2)  # the pair set does not depend on tree shape
    # a full rebuild in place agrees stage by stage with a fresh oracle run
    ctx.bvh_rebuild(bvh, mesh, p)
    nodes, sk, si = bvh.download()
    rk, ri = co.sort_keys(co.morton_keys(xyz2, idx, op))
    assert np.array_equal(sk, rk) and np.array_equal(si, ri)
    assert np.array_equal(ctx.self_collide(bvh), ref2)
    assert not np.array_equal(first, ref2)
    bvh.destroy()
    mesh.destroy()


def test_double_buffered_frames_async_upload(cd, co, ctx, mg):
    """b200cd_mesh_update_async / b200cd_mesh_wait: frame k+1 is uploaded on the copy stream while frame k is
    built and queried from the other mesh object; every frame's pair list equals the oracle's"""
    xyz0, idx = mg.cloth_fold(120, 120)
    rng = np.random.default_rng(11)
    frames_xyz = [xyz0] + [(xyz0 + rng.normal(0, 3e-4, xyz0.shape)).astype(np.float32) for _ in range(3)]
    op, p = co.default_params(), cd.default_params()
    want = [co.run(x, idx, op)[0] for x in frames_xyz]
    assert not np.array_equal(want[0], want[1])
    nv, nt = len(xyz0), len(idx)
    hx, hx_ptr = zip(*[cd.pinned_array((nv, 3), np.float32) for _ in range(2)])
    hi, hi_ptr = cd.pinned_array((nt, 3), np.uint32)
    hi[:] = idx
    meshes = [ctx.mesh_from_arrays(xyz0, idx) for _ in range(2)]
    bvh = ctx.bvh_build(meshes[0], p)
    hx[0][:] = frames_xyz[0]
    meshes[0].update_async_from_ptr(hx_ptr[0], hi_ptr)
    with pytest.raises(cd.B200cdError) as e:      # a mesh with an upload in flight cannot be used before wait()
        ctx.bvh_rebuild(bvh, meshes[0], p)
    assert e.value.status == cd.E_INVALID
    for k in range(len(frames_xyz)):
        cur = meshes[k % 2]
        if k + 1 < len(frames_xyz):
            hx[(k + 1) % 2][:] = frames_xyz[k + 1]
            meshes[(k + 1) % 2].update_async_from_ptr(hx_ptr[(k + 1) % 2], hi_ptr)
        cur.wait()
        ctx.bvh_rebuild(bvh, cur, p)
        assert np.array_equal(ctx.self_collide(bvh), want[k]), f"frame {k}"
    # the index check of an asynchronous upload is reported by wait()
    hi[0, 0] = nv
    meshes[0].update_async_from_ptr(0, hi_ptr)
    with pytest.raises(cd.B200cdError) as e:
        meshes[0].wait()
    assert e.value.status == cd.E_INVALID
    bvh.destroy()
    for m in meshes:
        m.destroy()


def test_async_slice_upload_without_peers(cd, co, ctx, mg):
    """b200cd_mesh_update_slice_async on one GPU (no peers set): the slice arithmetic of the chunked upload (four chunks,
    sizes that do not divide) and the index check; the multi-GPU pushes are covered by tests/run_multigpu_check.py"""
    xyz, idx = mg.soup(10_007, seed=17)
    nv, nt = len(xyz), len(idx)
    hx, hx_ptr = cd.pinned_array((nv, 3), np.float32)
    hi, hi_ptr = cd.pinned_array((nt, 3), np.uint32)
    hx[:] = xyz
    hi[:] = idx
    mesh = ctx.mesh_from_arrays(np.zeros_like(xyz), np.zeros_like(idx))
    cuts_v, cuts_t = [0, 1, 4099, nv], [0, 3333, 3334, nt]
    for a in range(3):
        v0, v1, t0, t1 = cuts_v[a], cuts_v[a + 1], cuts_t[a], cuts_t[a + 1]
        mesh.update_slice_async_from_ptr(hx_ptr + 12 * v0, v0, v1 - v0, hi_ptr + 12 * t0, t0, t1 - t0)
        mesh.wait()
    x2, i2 = mesh.download()
    assert np.array_equal(x2, xyz) and np.array_equal(i2, idx)
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    want, _ = co.run(xyz, idx, co.make_params(**UNIT))
    assert np.array_equal(ctx.self_collide(bvh), want)
    hi[5, 1] = nv + 3                                     # a bad index inside the second triangle slice
    mesh.update_slice_async_from_ptr(0, 0, 0, hi_ptr, 0, 100)
    with pytest.raises(cd.B200cdError) as e:
        mesh.wait()
    assert e.value.status == cd.E_INVALID
    bvh.destroy()
    mesh.destroy()


def test_capacity_error_reports_true_count(cd, ctx, mg):
    import ctypes as C
    xyz, idx = mg.soup(20000, seed=3)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    full = ctx.self_collide(bvh)
    out = np.full((10, 2), 0xFFFFFFFF, np.uint32)
    cnt = C.c_uint64()
    rc = cd.lib().b200cd_self_collide(ctx.h, bvh.h, out.ctypes.data_as(C.c_void_p), C.c_uint64(10), C.byref(cnt), C.c_int(1))
    assert rc == cd.E_CAPACITY and cnt.value == len(full)
    assert (out == 0xFFFFFFFF).all()  # nothing written (the reference overruns its 500-pair buffer, main.cu:81)
    rc = cd.lib().b200cd_self_collide(ctx.h, bvh.h, None, C.c_uint64(0), C.byref(cnt), C.c_int(0))
    assert rc in (cd.OK, cd.E_CAPACITY) and cnt.value == len(full)  # count-only call
    bvh.destroy()
    mesh.destroy()


def test_bad_vertex_index_is_rejected(cd, ctx):
    xyz = np.zeros((3, 3), np.float32)
    with pytest.raises(cd.B200cdError) as e:
        ctx.mesh_from_arrays(xyz, np.array([[0, 1, 3]], np.uint32))
    assert e.value.status == cd.E_INVALID


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partitioned_build_emulated_ranks_equal_single_gpu(cd, co, ctx, mg, world):
    """multi-GPU partitioned build (one Morton range per rank + ghost exchange) with the ranks emulated
    on one GPU: the merged pair list must equal the single-GPU list bit for bit"""
    import importlib
    mgpu = importlib.import_module("gpu-computing-course_b200.multigpu")
    for name, (xyz, idx), box in (("soup", mg.soup(60000, seed=17), UNIT), ("cloth", mg.cloth_fold(120, 120), None),
                                  ("sheets", mg.two_sheets(100), UNIT)):
        p = cd.make_params(**box) if box else cd.default_params()
        mesh = ctx.mesh_from_arrays(xyz, idx)
        bvh = ctx.bvh_build(mesh, p)
        full = ctx.self_collide(bvh, sorted=True)
        bvh.destroy()
        got, stats = mgpu.partitioned_self_collision_emulated(cd, ctx, mesh, p, world)
        assert sum(s["local_triangles"] for s in stats) == len(idx)
        assert np.array_equal(got, full), (name, world, stats)
        if world > 1:
            assert sum(s["ghosts"] for s in stats) > 0
        mesh.destroy()
    ctx.set_stream(None)


@pytest.mark.parametrize("world", [2, 3, 8, 16])
def test_partition_plan_kernel_equals_torch_twin(cd, ctx, world):
    """b200cd_partition_plan_device (splitters + per-owner counts in one launch) against the torch statement of the
    same rule (multigpu.PartitionedRank.splitters_from + the cumulative-sum differences)"""
    import torch
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    shift = 44
    g = torch.Generator(device="cpu").manual_seed(world)
    cases = []
    dense = torch.randint(0, 50, (65536,), generator=g, dtype=torch.int32)
    cases.append((dense * world, dense))                                        # every rank holds the same share
    sparse = torch.zeros(65536, dtype=torch.int32); sparse[torch.randint(0, 65536, (300,), generator=g)] = 7
    loc = torch.zeros(65536, dtype=torch.int32); loc[::2] = sparse[::2]
    cases.append((sparse, loc))                                                 # long runs of empty bins
    one = torch.zeros(65536, dtype=torch.int32); one[12345] = 1000
    cases.append((one, one // 2))                                               # everything in one bin
    last = torch.zeros(65536, dtype=torch.int32); last[65535] = 10; last[0] = 10
    cases.append((last, last))                                                  # the clamped last bin
    cases.append((torch.zeros(65536, dtype=torch.int32), torch.zeros(65536, dtype=torch.int32)))  # empty
    for gh, lh in cases:
        gh, lh = gh.to(dev), lh.to(dev)
        spl = torch.zeros(world - 1, dtype=torch.int64, device=dev)
        cnt = torch.zeros(world, dtype=torch.int32, device=dev)
        ctx.partition_plan_device(gh.data_ptr(), lh.data_ptr(), shift, world, spl.data_ptr(), cnt.data_ptr())
        csum = torch.cumsum(gh.to(torch.int64), 0)
        targets = (torch.arange(1, world, device=dev, dtype=torch.int64) * csum[-1]) // world
        bins = torch.clamp(torch.searchsorted(csum, targets), max=65534)
        lcs = torch.cumsum(lh.to(torch.int64), 0)
        bounds = torch.cat([lcs[bins], lcs[-1:]])
        want_cnt = torch.diff(bounds, prepend=bounds.new_zeros(1)).to(torch.int32)
        torch.cuda.synchronize()
        assert torch.equal(spl, (bins + 1) << shift), (world, spl.tolist()[:4], bins.tolist()[:4])
        assert torch.equal(cnt, want_cnt), (world, cnt.tolist(), want_cnt.tolist())
    ctx.set_stream(None)


def test_unique_triangle_set_on_device(cd, co, ctx, mg):
    """the second half of the reference's output (makeAndPrintSet, main.cu:33-45): sorted unique IDs of all colliding
    triangles, computed on the device - against numpy on the pair list, for sparse, dense and empty results"""
    for xyz, idx, p in ((*mg.cloth_fold(60, 60), cd.default_params()), (*mg.soup(50000, seed=3), cd.make_params(**UNIT)),
                        (*mg.flag(30, 30), cd.default_params()), (*mg.soup(64, h=1e-6, seed=1), cd.make_params(**UNIT))):
        mesh = ctx.mesh_from_arrays(xyz, idx)
        bvh = ctx.bvh_build(mesh, p)
        for srt in (True, False):
            pairs = ctx.self_collide(bvh, sorted=srt)
            ids = ctx.unique_triangles(bvh)
            assert ids.dtype == np.uint32 and np.array_equal(ids, np.unique(pairs))
        bvh.destroy()
        mesh.destroy()
    # capacity protocol of the host form, through the raw C ABI
    import ctypes as C
    xyz, idx = mg.cloth_fold(40, 40)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.default_params())
    pairs = ctx.self_collide(bvh)
    want = np.unique(pairs)
    cnt = C.c_uint64()
    small = np.zeros(4, np.uint32)
    rc = cd.lib().b200cd_unique_triangles(ctx.h, bvh.h, small.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint64(4), C.byref(cnt))
    assert rc == cd.E_CAPACITY and cnt.value == len(want) and not small.any()
    bvh.destroy()
    mesh.destroy()


def test_sort_fallback_without_cooperative_launch(tmp_path):
    """the hybrid sort's fallback (runs of equal high bits longer than the fix-up handles) is normally ONE cooperative
    launch; where that launch is refused the same passes run as eight conditional launches - forced here with
    B200CD_COOP=0 in a fresh process (the knob is read once): keys, tie order and pairs must equal the oracle's"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import importlib, sys
import numpy as np
sys.path.insert(0, %r)
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
from oracle import cdoracle as co
bx, bi = mg.soup(300, seed=11)
xyz = np.tile(bx, (64, 1))                      # 64 copies of every triangle: runs of 64 equal keys
idx = (np.tile(bi, (64, 1)) + (np.arange(64, dtype=np.uint32).repeat(len(bi)) * np.uint32(len(bx)))[:, None]).astype(np.uint32)
ctx = cd.Context(0)
mesh = ctx.mesh_from_arrays(xyz, idx)
p = cd.make_params((0, 0, 0), (1, 1, 1))
bvh = ctx.bvh_build(mesh, p)
for _ in range(2):
    ctx.bvh_rebuild(bvh, mesh, p)
_, sk, si = bvh.download(nodes=False)
op = co.make_params((0, 0, 0), (1, 1, 1))
rk, ri = co.sort_keys(co.morton_keys(xyz, idx, op))
assert np.array_equal(sk, rk) and np.array_equal(si, ri), "sorted keys / tie order"
ref, _ = co.run(xyz, idx, op)
assert np.array_equal(ctx.self_collide(bvh, sorted=True), ref), "pairs"
print("ok", len(ref))
""" % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, B200CD_COOP="0"))
    assert out.returncode == 0 and out.stdout.strip().startswith("ok"), out.stderr[-2000:]


@pytest.mark.parametrize("variant", ["1", "3"])
def test_traversal_variants_give_the_same_pairs(variant):
    """B200CD_TRAVERSAL=1 (one query per thread, per-thread stack) and =3 (one query per thread, STACKLESS: the escape
    pointer is implicit in the node numbering) against the oracle, in a fresh process (the knob is read once); the
    default (=2, persistent lanes) is what every other test runs"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import importlib, sys
import numpy as np
sys.path.insert(0, %r)
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
from oracle import cdoracle as co
ctx = cd.Context(0)
for (xyz, idx), box in ((mg.soup(60000, seed=2), ((0, 0, 0), (1, 1, 1))), (mg.cloth_fold(90, 90), None), (mg.two_sheets(64), ((0, 0, 0), (1, 1, 1))),
                        (mg.soup(3, h=0.4, seed=1), ((0, 0, 0), (1, 1, 1)))):
    p = cd.make_params(*box) if box else cd.default_params()
    op = co.make_params(*box) if box else co.default_params()
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, p)
    ref, _ = co.run(xyz, idx, op)
    assert np.array_equal(ctx.self_collide(bvh, sorted=True), ref), len(idx)
    bvh.destroy(); mesh.destroy()
print("ok")
""" % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, B200CD_TRAVERSAL=variant))
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


@pytest.mark.parametrize("knobs", [{"B200CD_BROAD_QUANT": "1"}, {"B200CD_BROAD_QUANT": "0"},
                                   {"B200CD_BROAD_QUANT": "1", "B200CD_BROAD_GRID": "persist"}])
def test_quantised_node_traversal_gives_the_same_pairs(knobs):
    """The traversal on 32-byte quantised nodes (default on soups) forced on for meshes too (=1: shared-vertex filter +
    quantised walk, touching boxes everywhere), off everywhere (=0), and with persistent warps - against the oracle, in a
    fresh process. Covers: Morton boxes much larger and much smaller than the mesh (cells clamp at the border), rebuilds
    (the grid moves to the previous root box), refits after the vertices moved out of that box, a degenerate flat mesh."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import importlib, sys
import numpy as np
sys.path.insert(0, %r)
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
from oracle import cdoracle as co
ctx = cd.Context(0)
flat = mg.soup(4000, seed=9)
flat[0][:, 2] = 0.5                                   # every vertex in one plane: zero extent along z
cases = ((mg.soup(60000, seed=2), ((0, 0, 0), (1, 1, 1))), (mg.cloth_fold(90, 90), None), (mg.two_sheets(64), ((0, 0, 0), (1, 1, 1))),
         (mg.soup(20000, seed=3), ((-40, -40, -40), (100, 100, 100))),      # grid far coarser than the triangles
         (mg.soup(20000, seed=4), ((0.4, 0.4, 0.4), (0.2, 0.2, 0.2))),      # most of the mesh outside the box
         (flat, ((0, 0, 0), (1, 1, 1))), (mg.soup(3, h=0.4, seed=1), ((0, 0, 0), (1, 1, 1))))
for (xyz, idx), box in cases:
    p = cd.make_params(*box) if box else cd.default_params()
    op = co.make_params(*box) if box else co.default_params()
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, p)
    ref, _ = co.run(xyz, idx, op)
    assert np.array_equal(ctx.self_collide(bvh, sorted=True), ref), ("build", len(idx))
    ctx.bvh_rebuild(bvh, mesh, p)                     # grid over the first build's root box
    assert np.array_equal(ctx.self_collide(bvh, sorted=True), ref), ("rebuild", len(idx))
    moved = (xyz * np.float32(1.5) + np.float32(0.25)).astype(np.float32)   # leaves the previous root box
    mesh.update(moved)
    ctx.bvh_refit(bvh, mesh)
    ref2, _ = co.run(moved, idx, op)
    got = ctx.self_collide(bvh, sorted=True)
    assert np.array_equal(got, ref2), ("refit", len(idx))
    bvh.destroy(); mesh.destroy()
print("ok")
""" % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, **knobs))
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
