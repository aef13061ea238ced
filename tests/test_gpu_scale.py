"""GPU: BASELINE.json's full-size configurations through size-independent properties, plus a
mid-size bit-exact comparison with the oracle. (~1-2 minutes on a B200 box.)"""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

UNIT = dict(origin=(0.0, 0.0, 0.0), extent=(1.0, 1.0, 1.0))


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


def test_soup_16m_full_size_properties(cd, co, ctx, mg):
    """C4 at full size: structure counters, pair-set invariants, shard union, sampled re-verification"""
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


def test_cloth_1m_dense_contacts_bit_exact(cd, co, ctx, mg):
    """C3 at full size (1 002 528 triangles, shared vertices, reference Morton box)"""
    xyz, idx = mg.cloth_fold()
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, cd.default_params())
    pairs = ctx.self_collide(bvh, sorted=True)
    ref, _ = co.run(xyz, idx, co.default_params())
    assert np.array_equal(pairs, ref) and len(pairs) > 200000
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
    bvh.destroy()
    mesh.destroy()
