"""GPU: BASELINE.json's full-size configurations through size-independent properties, plus a
mid-size bit-exact comparison with the oracle. (~1-2 minutes on a B200 box.)"""
import importlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

UNIT = dict(origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# full-size expected results: made once in the build container by tests/golden/make_checksums.py from the reference's
# own host functions AND the oracle restatement (full lists compared there)
CHECKSUMS = json.load(open(os.path.join(ROOT, "tests", "golden", "checksums.json")))
ps = importlib.import_module("gpu-computing-course_b200.pairsum")


def matches_golden(name, pairs, n):
    g = CHECKSUMS[name]
    assert g["triangles"] == n
    assert len(pairs) == g["pairs"], (len(pairs), g["pairs"])
    assert ps.pairs_checksum_np(pairs) == g["checksum"]


def pair_set_properties(pairs, n):
    assert pairs.dtype == np.uint32 and pairs.shape[1] == 2
    assert (pairs[:, 0] < pairs[:, 1]).all()          # tri_contact.cuh:81
    assert pairs.max() < n
    w = pairs[:, 0].astype(np.uint64) << np.uint64(32) | pairs[:, 1].astype(np.uint64)
    assert (w[1:] > w[:-1]).all()                     # sorted lexicographically, no duplicates


def test_soup_4m_bit_exact_vs_oracle(cd, co, ctx, mg):
    xyz, idx = mg.soup(1 << 22, seed=1234)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    pairs = ctx.self_collide(bvh, sorted=True)
    ref, tm = co.run(xyz, idx, co.make_params((0, 0, 0), (1, 1, 1)))
    assert np.array_equal(pairs, ref)
    _, sk, si = bvh.download(nodes=False)
    rk, ri = co.sort_keys(co.morton_keys(xyz, idx, co.make_params((0, 0, 0), (1, 1, 1))))
    assert np.array_equal(sk, rk) and np.array_equal(si, ri)
    bvh.destroy()
    mesh.destroy()


def test_soup_16m_bit_exact_vs_oracle(cd, co, ctx, mg):
    """C4 at full size: the whole pair list against the oracle, plus structure counters, pair-set invariants, shard union"""
    n = 1 << 24
    xyz, idx = mg.soup(n, seed=1234)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    chk = bvh.validate(mesh)
    assert chk["null_parent_internal"] == 1 and sum(chk.values()) == 1 + chk["unsorted_keys"], chk
    pairs = ctx.self_collide(bvh, sorted=True)
    st = ctx.stats()
    assert st["pairs"] == len(pairs) and st["candidates"] >= st["pairs"]
    assert len(pairs) > n // 8                         # ~0.19 contacts per triangle at this density
    pair_set_properties(pairs, n)
    # THE headline configuration, full list against the oracle (cpu.cuh:196-245 over every leaf, tri_contact.cuh:80-87)
    # and against the count + checksum the reference's own host functions gave for this mesh
    ref, _ = co.run(xyz, idx, co.make_params((0, 0, 0), (1, 1, 1)))
    assert np.array_equal(pairs, ref), "soup16m: GPU pair list differs from the oracle"
    del ref
    matches_golden("soup16m", pairs, n)
    # the unique-triangle set (main.cu:33-45) on the device
    assert np.array_equal(ctx.unique_triangles(bvh), np.unique(pairs))
    # union of 8 block-cyclic shards == the full list (what the 8-GPU path gathers)
    parts = [ctx.self_collide(bvh, sorted=False, shard=s, nshards=8, chunk=1 << 14) for s in range(8)]
    merged = np.concatenate(parts)
    assert len(merged) == len(pairs)
    w = np.sort(merged[:, 0].astype(np.uint64) << np.uint64(32) | merged[:, 1].astype(np.uint64))
    assert np.array_equal(w, pairs[:, 0].astype(np.uint64) << np.uint64(32) | pairs[:, 1].astype(np.uint64))
    # idempotence: a rebuild + second query reproduces the list bit for bit
    ctx.bvh_rebuild(bvh, mesh, cd.make_params(**UNIT))
    assert np.array_equal(ctx.self_collide(bvh, sorted=True), pairs)
    # every sampled reported pair is a contact for the oracle's SAT, called as (lower ID, higher ID)
    rng = np.random.default_rng(1)
    for a, b in pairs[rng.integers(0, len(pairs), 3000)]:
        t = np.concatenate([xyz[idx[a]].reshape(-1), xyz[idx[b]].reshape(-1)]).astype(np.float64)
        assert co.tri_contact(t) == 1
    # a spatial sub-block re-checked exhaustively: all triangles with centroid in a small cell
    cen = xyz[idx.reshape(-1)].reshape(n, 3, 3).mean(axis=1)
    sel = np.where((np.abs(cen - 0.5) < 0.03).all(axis=1))[0]
    assert 1000 < len(sel) < 20000
    sub_pairs = co.brute_force(xyz, idx[sel])  # indices are positions in sel
    sub = np.stack([sel[sub_pairs[:, 0]], sel[sub_pairs[:, 1]]], axis=1).astype(np.uint32)
    insel = np.zeros(n, bool)
    insel[sel] = True
    got = pairs[insel[pairs[:, 0]] & insel[pairs[:, 1]]]
    assert np.array_equal(got, sub)
    bvh.destroy()
    mesh.destroy()


def test_sheets_64m_matches_reference_checksum(cd, ctx, mg):
    """C5 (2^26 triangles) on one GPU: count + checksum of the sorted list equal what the reference's host functions
    and the oracle produced for this mesh (tests/golden/checksums.json; the oracle needs 100 s and 20 GB for it)"""
    xyz, idx = mg.two_sheets(4096, seed=7)
    n = len(idx)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    del xyz, idx
    bvh = ctx.bvh_build(mesh, cd.make_params(**UNIT))
    chk = bvh.validate(mesh)
    assert chk["null_parent_internal"] == 1 and sum(chk.values()) == 1 + chk["unsorted_keys"], chk
    pairs = ctx.self_collide(bvh, sorted=True)
    pair_set_properties(pairs, n)
    matches_golden("sheets64m", pairs, n)
    assert np.array_equal(ctx.unique_triangles(bvh), np.unique(pairs))
    bvh.destroy()
    mesh.destroy()


def test_cloth_1m_dense_contacts_bit_exact(cd, co, ctx, mg):
    """C3 at full size (1 002 528 triangles, shared vertices, reference Morton box)"""
    xyz, idx = mg.cloth_fold()
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.default_params())
    pairs = ctx.self_collide(bvh, sorted=True)
    ref, _ = co.run(xyz, idx, co.default_params())
    assert np.array_equal(pairs, ref) and len(pairs) > 200000
    matches_golden("cloth1m", pairs, len(idx))
    assert np.array_equal(ctx.unique_triangles(bvh), np.unique(pairs))
    bvh.destroy()
    mesh.destroy()


def test_flag_standin_full_size_bit_exact(cd, co, ctx, mg):
    """C1/C2 stand-in at the bundled mesh's size (1 262 460 triangles)"""
    xyz, idx = mg.flag()
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.default_params())
    pairs = ctx.self_collide(bvh, sorted=True)
    ref, _ = co.run(xyz, idx, co.default_params())
    assert np.array_equal(pairs, ref) and len(pairs) > 0
    matches_golden("flag1m", pairs, len(idx))
    bvh.destroy()
    mesh.destroy()
