"""Multi-GPU self-collision: one process (rank) per GPU, torch.distributed for the plumbing.

SURVEY.md §8(e): the query shards naturally (queries are independent, the BVH is read-only) and
has ONE exchange step, at the end. Each rank holds the whole BVH — built locally from the
replicated mesh (the build is deterministic, so all ranks hold bit-identical trees and nothing
crosses NVLink) or received from rank 0 (broadcast_bvh) — traverses its own subset of the
Morton-sorted query triangles (b200cd_self_collide_device with shard/nshards/chunk) and the
per-rank pair lists are gathered on rank 0 and sorted there (b200cd_sort_pairs_device).

The reference has no multi-GPU or collective code at all.

Nothing here computes on the CPU: `gather_pairs` moves tensors that live wherever the process
group's backend wants them (CUDA for nccl; CPU tensors for the gloo tests of the host logic).
"""
import numpy as np
import torch
import torch.distributed as dist

DEFAULT_CHUNK = 1 << 14  # sorted leaves per block-cyclic chunk: spatially coherent, load-balanced


# ---- shard arithmetic: the same partition b200cd_self_collide_shard applies (csrc/api.cu run_query)

QUERY_BLOCK = 256  # B200CD_QUERY_BLOCK: traversal blocks are 256 consecutive sorted leaves


def resolve_chunk(n, nshards, chunk):
    """chunk = 0 means one contiguous slice per shard; chunks are whole traversal blocks (rounded up)"""
    c = chunk if chunk else (n + nshards - 1) // max(nshards, 1)
    return max(1, (c + QUERY_BLOCK - 1) // QUERY_BLOCK) * QUERY_BLOCK


def shard_of_position(pos, n, nshards, chunk=DEFAULT_CHUNK):
    """which shard owns the query at sorted-leaf position `pos` (numpy array or int)"""
    c = resolve_chunk(n, nshards, chunk)
    return (np.asarray(pos) // c) % nshards


def shard_positions(shard, n, nshards, chunk=DEFAULT_CHUNK):
    """all sorted-leaf positions traversed by `shard`, ascending"""
    pos = np.arange(n, dtype=np.int64)
    return pos[shard_of_position(pos, n, nshards, chunk) == shard]


def owner_of_pairs(pairs, sorted_ids, nshards, chunk=DEFAULT_CHUNK):
    """shard that reports each colliding pair: a pair is discovered by the query with the SMALLER
    sorted position of its two leaves (collide.cu prunes subtrees that end at or before the query)"""
    n = len(sorted_ids)
    pos_of_id = np.empty(n, np.int64)
    pos_of_id[np.asarray(sorted_ids, np.int64)] = np.arange(n)
    p = np.asarray(pairs, np.int64).reshape(-1, 2)
    q = np.minimum(pos_of_id[p[:, 0]], pos_of_id[p[:, 1]])
    return shard_of_position(q, n, nshards, chunk)


# ---- zero-copy torch view of library-owned device memory

class _DeviceSpan:
    def __init__(self, ptr, nwords):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def device_pairs_as_tensor(ptr, count, device):
    """(count,) int64 view of a uint32[count][2] pair list: word = hi_id << 32 | lo_id"""
    if count == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DeviceSpan(ptr, count), device=device)


# ---- the one exchange step

def gather_pairs(local, dst=0, group=None):
    """Gather variable-length 1-D int64 tensors (packed pairs) on rank `dst`.

    all_gather of the counts, then grouped point-to-point send/recv straight into the destination
    buffer at each rank's offset (ncclSend/ncclRecv under the nccl backend). Returns
    (merged tensor on dst | None elsewhere, counts list)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts_t = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts_t, cnt, group=group)
    counts = [int(c.item()) for c in counts_t]
    if world == 1:
        return local, counts
    ops = []
    merged = None
    if rank == dst:
        merged = torch.empty(sum(counts), dtype=torch.int64, device=local.device)
        off = 0
        for r, c in enumerate(counts):
            if r == dst:
                merged[off:off + c].copy_(local)
            elif c:
                ops.append(dist.P2POp(dist.irecv, merged[off:off + c], r, group=group))
            off += c
    elif local.numel():
        ops.append(dist.P2POp(dist.isend, local.contiguous(), dst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return merged, counts


def unpack_pairs(words):
    """int64 packed words -> (count, 2) uint32 numpy array, lower ID first"""
    a = words.detach().cpu().numpy().astype(np.int64, copy=False)
    return np.ascontiguousarray(a).view(np.uint32).reshape(-1, 2)


def broadcast_bvh(cd, ctx, bvh, ntris, src=0, group=None):
    """Replicate a BVH built on rank `src`: NCCL broadcast of the three device blobs
    (traversal nodes, leaf records, sorted ids). Other ranks pass bvh=None and get a new handle."""
    rank = dist.get_rank(group)
    if rank != src:
        bvh = ctx.bvh_alloc_like(ntris)
    v = bvh.view()
    dev = torch.device("cuda", ctx.device)
    for ptr, nbytes in ((v.d_nodes, v.nodes_bytes), (v.d_leaves, v.leaves_bytes), (v.d_ids, v.ids_bytes)):
        if nbytes:
            t = torch.as_tensor(_DeviceSpan(ptr, nbytes // 8), device=dev) if nbytes % 8 == 0 else None
            if t is None:  # ids blob with an odd triangle count: 4-byte view
                span = _DeviceSpan(ptr, nbytes // 4)
                span.__cuda_array_interface__["typestr"] = "<i4"
                t = torch.as_tensor(span, device=dev)
            dist.broadcast(t, src=src, group=group)
    return bvh


class ShardedSelfCollision:
    """Per-rank driver: build (replicated) -> sharded query -> gather + sort on rank 0."""

    def __init__(self, cd, ctx, group=None, chunk=DEFAULT_CHUNK):
        self.cd, self.ctx, self.group, self.chunk = cd, ctx, group, chunk
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", ctx.device)
        self.counts = [0] * self.world
        # library kernels, NCCL transfers and the final sort are ordered on ONE stream: torch's current one
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, bvh, mesh, params, rebuild=True):
        """one build + sharded query + gather; returns the device tensor of all pairs on rank 0"""
        if rebuild:
            self.ctx.bvh_rebuild(bvh, mesh, params)
        # per-rank lists stay unsorted: the merged list is sorted once, on rank 0
        ptr, count = self.ctx.self_collide_device(bvh, sorted=(self.world == 1), shard=self.rank,
                                                  nshards=self.world, chunk=self.chunk if self.world > 1 else 0)
        local = device_pairs_as_tensor(ptr, count, self.device)
        if self.world == 1:
            self.counts = [count]
            return local
        merged, self.counts = gather_pairs(local, 0, self.group)
        if self.rank == 0 and merged.numel() > 1:
            self.ctx.sort_pairs_device(merged.data_ptr(), merged.numel(), id_bits=max(1, int(bvh.ntris - 1).bit_length()))
        return merged
