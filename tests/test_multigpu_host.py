"""CPU: the multi-GPU host logic (shard arithmetic + the gather exchange) with world_size-2 gloo.

The per-rank pair lists are stand-ins computed by the CPU oracle (checker) and split with the same
partition the library applies on the device; the GPU twin of this test is
tests/test_gpu_parity.py::test_sharded_query_union_equals_full.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "gpu-computing-course_b200"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, chunk, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mgpu = importlib.import_module(f"{PKG}.multigpu")
        mg = importlib.import_module(f"{PKG}.meshgen")
        from oracle import cdoracle as co
        xyz, idx = mg.soup(6000, seed=21)
        op = co.make_params((0, 0, 0), (1, 1, 1))
        sk, si = co.sort_keys(co.morton_keys(xyz, idx, op))
        h = co.hierarchy(sk)
        b = co.refit(xyz, idx, si, h)
        pairs, _ = co.self_collide(xyz, idx, si, h, b)
        full = co.sort_pairs(pairs.copy())
        owner = mgpu.owner_of_pairs(pairs, si, world, chunk)
        mine = np.ascontiguousarray(pairs[owner == rank])
        words = torch.from_numpy(mine.view(np.int64).reshape(-1).copy())
        merged, counts = mgpu.gather_pairs(words, 0)
        assert counts[rank] == len(mine) and sum(counts) == len(full)
        if rank == 0:
            got = co.sort_pairs(mgpu.unpack_pairs(merged).copy())
            assert np.array_equal(got, full)
            np.save(os.path.join(tmpdir, "ok.npy"), np.array(counts))
        else:
            assert merged is None
        # the single-collective variant: padded all-gather, sentinels sort behind every real pair
        id_bits = max(1, int(len(idx) - 1).bit_length())
        words = torch.from_numpy(mine.view(np.int64).reshape(-1).copy())
        buf, counts2 = mgpu.gather_pairs_padded(words, id_bits)
        assert counts2 == counts and buf.numel() == world * max(counts)
        u = buf.numpy().view(np.uint32).reshape(-1, 2)
        order = np.lexsort((u[:, 1], u[:, 0]))  # what b200cd_sort_pairs_device does on the GPU: by lo id, then hi id
        assert np.array_equal(u[order][: len(full)], full)
        assert (u[order][len(full):] == (1 << id_bits) - 1).all()
        # nothing to gather anywhere
        buf, counts3 = mgpu.gather_pairs_padded(torch.zeros(0, dtype=torch.int64), id_bits)
        assert buf.numel() == 0 and counts3 == [0] * world
        # one rank empty: its slot is all sentinel
        buf, counts3 = mgpu.gather_pairs_padded(torch.arange(3 if rank == 0 else 0, dtype=torch.int64), id_bits)
        assert counts3 == [3, 0][:world] or world != 2
        assert buf.numel() == 3 * world and buf[:3].tolist() == [0, 1, 2] and (buf[3:] == (((1 << id_bits) - 1) << 32 | ((1 << id_bits) - 1))).all()
        # empty contribution from one rank
        words = torch.zeros(0 if rank == 1 else 5, dtype=torch.int64)
        merged, counts = mgpu.gather_pairs(words, 0)
        assert counts == [5, 0][:world] or world != 2
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunk", [0, 64])
def test_gather_pairs_world2_gloo(tmp_path, chunk):
    mp.spawn(_worker, args=(2, _free_port(), chunk, str(tmp_path)), nprocs=2, join=True)
    counts = np.load(os.path.join(tmp_path, "ok.npy"))
    assert counts.sum() > 500 and (counts > 0).all()


def test_shard_partition_covers_every_query_once():
    mgpu = importlib.import_module(f"{PKG}.multigpu")
    for n, nshards, chunk in [(1, 1, 0), (10, 3, 0), (1000, 8, 64), (4097, 4, 4096), (100, 8, 16), (5, 8, 0)]:
        seen = np.concatenate([mgpu.shard_positions(s, n, nshards, chunk) for s in range(nshards)])
        assert np.array_equal(np.sort(seen), np.arange(n)), (n, nshards, chunk)
    assert mgpu.resolve_chunk(10, 3, 0) == 256 and mgpu.resolve_chunk(1000, 3, 300) == 512
