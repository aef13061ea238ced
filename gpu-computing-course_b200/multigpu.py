"""Multi-GPU self-collision: one process (rank) per GPU, torch.distributed for the plumbing.

SURVEY.md §8(e): the query shards naturally (queries are independent, the BVH is read-only) and
has ONE exchange step, at the end. Each rank holds the whole BVH — built locally from the
replicated mesh (the build is deterministic, so all ranks hold bit-identical trees and nothing
crosses NVLink) or received from rank 0 (BvhBroadcaster: NCCL broadcast by the library) — traverses its own subset of the
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
    def __init__(self, ptr, nwords, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TORCH_OF = {"<i8": torch.int64, "<i4": torch.int32, "<f4": torch.float32}


def device_view(ptr, nelem, typestr, device):
    """zero-copy 1-D torch view of library-owned device memory"""
    if nelem == 0:
        return torch.empty(0, dtype=_TORCH_OF[typestr], device=device)
    return torch.as_tensor(_DeviceSpan(ptr, nelem, typestr), device=device)


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
    counts = torch.cat(counts_t).cpu().tolist()  # one device -> host read for all ranks' counts
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


def gather_pairs_padded(local, id_bits, group=None):
    """The same exchange as ONE fixed-size collective: all-gather of the counts (one host read), then a single
    all-gather of every rank's list padded to the longest one. The padding is a sentinel word that sorts behind
    every real pair (both halves = 2^id_bits - 1; a real pair has lo < hi), so the caller sorts the whole
    buffer and keeps the first sum(counts) words - no per-rank offsets, no grouped send/recv launches.
    Returns (buffer of world * max(counts) words - identical on every rank -, counts list)."""
    world = dist.get_world_size(group)
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts_t = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts_t, cnt, group=group)
    counts = counts_t.cpu().tolist()
    cap = max(counts)
    if cap == 0:
        return torch.empty(0, dtype=torch.int64, device=local.device), counts
    top = (1 << id_bits) - 1
    mine = torch.full((cap,), (top << 32) | top, dtype=torch.int64, device=local.device)
    mine[:local.numel()] = local
    buf = torch.empty(world * cap, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(buf, mine, group=group)
    return buf, counts


def unpack_pairs(words):
    """int64 packed words -> (count, 2) uint32 numpy array, lower ID first"""
    a = words.detach().cpu().numpy().astype(np.int64, copy=False)
    return np.ascontiguousarray(a).view(np.uint32).reshape(-1, 2)


class BvhBroadcaster:
    """Replicated mode, "each GPU ... receives the BVH, broadcast via NCCL over NVLink" (BASELINE.json north_star): the
    LIBRARY owns an NCCL communicator (b200cd_dist_nccl_init; the 128-byte unique id travels through torch.distributed
    once) and b200cd_dist_broadcast_bvh sends the three device blobs of a BVH built on `src` - traversal nodes, leaf
    records, sorted ids - to BVHs of the same shape on the other ranks."""

    def __init__(self, cd, ctx, ntris, group=None):
        self.cd, self.ctx, self.group, self.ntris = cd, ctx, group, ntris
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ctx.set_stream(torch.cuda.current_stream(torch.device("cuda", ctx.device)).cuda_stream)
        self.dist = ctx.dist_create(self.rank, self.world, 0)       # only the communicator is used
        box = [cd.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        self.dist.nccl_init(box[0])

    def receive_or_send(self, bvh, src=0):
        """rank `src` passes its built BVH; the others pass None (a BVH is allocated) or a BVH from an earlier call"""
        if self.rank != src and bvh is None:
            bvh = self.ctx.bvh_alloc_like(self.ntris)
        self.dist.broadcast_bvh(bvh, src)
        return bvh

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.dist.destroy()


class ShardedSelfCollision:
    """Per-rank driver: build (replicated) -> sharded query -> gather + sort on rank 0."""

    def __init__(self, cd, ctx, group=None, chunk=DEFAULT_CHUNK):
        self.cd, self.ctx, self.group, self.chunk = cd, ctx, group, chunk
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", ctx.device)
        self.counts = [0] * self.world
        self.broadcaster = None  # set to a BvhBroadcaster: rank 0 alone builds and the others RECEIVE the BVH over NCCL
        # library kernels, NCCL transfers and the final sort are ordered on ONE stream: torch's current one
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, bvh, mesh, params, rebuild=True):
        """one build + sharded query + gather; returns the device tensor of all pairs on rank 0"""
        if rebuild and self.broadcaster is not None:
            if self.rank == 0:
                self.ctx.bvh_rebuild(bvh, mesh, params)
            self.broadcaster.receive_or_send(bvh, 0)
        elif rebuild:
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


# ================================================================ partitioned build (one Morton range per rank)
#
# Replicating the BVH caps the speed-up: every rank repeats the whole build (measured on 2 x B200,
# 64 M triangles: 27.5 ms vs 34.6 ms on one GPU). Here each rank owns ONE Morton range of the
# triangles: it sorts, builds and queries only that range and exchanges the thin layer of triangles
# whose boxes reach into a higher rank's range (ghost queries). The mesh itself (vertices + indices)
# is replicated input; what crosses NVLink per step is 12 B per triangle once ((key, id) to its
# owner), 6 KB of coarse boxes per rank and 64 B per ghost.

PARTITION_BOXES = 256  # coarse boxes per rank used for ghost selection


class PartitionedRank:
    """Everything one rank does between the communication steps (see PartitionedSelfCollision)."""

    def __init__(self, cd, ctx, mesh, params, rank, world, slack=1.5):
        self.cd, self.ctx, self.mesh, self.params, self.rank, self.world = cd, ctx, mesh, params, rank, world
        self.dev = torch.device("cuda", ctx.device)
        n = mesh.ntris
        self.n = n
        self.lo, self.hi = rank * n // world, (rank + 1) * n // world  # my slice of the INPUT triangles
        self.cnt = self.hi - self.lo
        self.cap = min(n, int(n / world * slack) + 65536)
        self.ghost_cap = max(n // world // 2, 65536)
        self.bvh = ctx.bvh_alloc_partial(self.cap, self.ghost_cap, world)
        self.shift = 44 if int(params.key_bits) == 63 else 14  # top 16 of the 60 / 30 bits in-box keys use
        self.keys = torch.empty(max(self.cnt, 1), dtype=torch.int64, device=self.dev)
        self.pkeys = torch.empty(max(self.cnt, 1), dtype=torch.int64, device=self.dev)
        self.pids = torch.empty(max(self.cnt, 1), dtype=torch.int32, device=self.dev)
        self.hist = torch.zeros(65536, dtype=torch.int32, device=self.dev)
        self.boxes = torch.empty(PARTITION_BOXES * 6, dtype=torch.float32, device=self.dev)
        kptr, iptr, cap = ctx.bvh_key_buffers(self.bvh)
        self.rkeys = device_view(kptr, cap, "<i8", self.dev)   # where my range's (key, id) arrive
        self.rids = device_view(iptr, cap, "<i4", self.dev)
        self.nlocal = 0
        self.nghost = 0

    # -- phase 1: keys of my input slice + histogram of their top 16 bits
    def keys_and_histogram(self):
        self.ctx.morton_keys_device(self.mesh, self.params, self.lo, self.cnt, self.keys.data_ptr())
        self.hist.zero_()
        self.ctx.key_histogram_device(self.keys.data_ptr(), self.cnt, self.shift, self.hist.data_ptr())
        return self.hist

    def splitters_from(self, global_hist):
        csum = torch.cumsum(global_hist.to(torch.int64), 0)
        targets = (torch.arange(1, self.world, device=self.dev, dtype=torch.int64) * csum[-1]) // self.world
        bins = torch.searchsorted(csum, targets)                     # first bin whose cumulative count reaches the target
        bins = torch.clamp(bins, max=65534)                          # (keys beyond the histogram's range share the last bin)
        self.split_bins = bins                                       # rank r owns the bins (bins[r-1], bins[r]]
        self.splitters = ((bins + 1) << self.shift).contiguous()     # keys >= splitter r-1 belong to rank >= r
        return self.splitters

    # -- phase 2: splitters from the GLOBAL histogram, bucket my (key, id) by owner
    def partition(self, global_hist):
        self.splitters_from(global_hist)
        counts = self.ctx.partition_keys_device(self.bvh, self.keys.data_ptr(), self.lo, self.cnt,
                                                self.splitters.data_ptr() if self.world > 1 else 0, self.world - 1,
                                                self.pkeys.data_ptr(), self.pids.data_ptr())
        pieces, off = [], 0
        for c in counts:
            pieces.append((self.pkeys[off:off + c], self.pids[off:off + c]))
            off += c
        return counts, pieces

    def key_recv_views(self, counts_from):
        total = sum(counts_from)
        if total > self.cap:
            raise RuntimeError(f"rank {self.rank}: {total} triangles in my Morton range, capacity {self.cap} "
                               f"(raise PartitionedSelfCollision(slack=...))")
        views, off = [], 0
        for c in counts_from:
            views.append((self.rkeys[off:off + c], self.rids[off:off + c]))
            off += c
        self.nlocal = total
        return views

    # -- phase 3: sort + tree over my range, then my coarse boxes
    def build(self):
        self.ctx.bvh_build_partial(self.bvh, self.mesh, self.params, self.nlocal)
        self.ctx.bvh_chunk_boxes_device(self.bvh, PARTITION_BOXES, self.boxes.data_ptr())
        return self.boxes

    # -- phase 4b (while the ghosts travel): pairs inside my range
    def collide_local(self):
        self.ctx.self_collide_device(self.bvh, sorted=False)

    # -- phase 4: my leaves that reach into a HIGHER rank's coarse boxes
    def select_ghosts(self, all_boxes):
        mask = 0
        for p in range(self.rank + 1, self.world):
            mask |= 1 << p
        ptr, stride, counts = self.ctx.select_ghosts_device(self.bvh, all_boxes.data_ptr(), self.world, PARTITION_BOXES, mask)
        pieces = [device_view(ptr + 64 * stride * p, 8 * counts[p], "<i8", self.dev) for p in range(self.world)]
        return counts, pieces

    def ghost_recv_views(self, counts_from):
        total = sum(counts_from)
        ptr, cap = self.ctx.bvh_ghost_buffer(self.bvh)
        if total > cap:
            raise RuntimeError(f"rank {self.rank}: {total} ghosts, room for {cap}")
        views, off = [], 0
        for c in counts_from:
            views.append(device_view(ptr + 64 * off, 8 * c, "<i8", self.dev))
            off += c
        self.nghost = total
        return views

    # -- phase 5: ghosts against my tree, appended to my local pairs; returns all of them (packed int64 words)
    def collide_ghosts(self):
        ptr, count = self.ctx.collide_ghosts_device(self.bvh, self.nghost, keep_pairs=True)
        return device_pairs_as_tensor(ptr, count, self.dev)


def _exchange_start(rank, world, send_lists, recv_lists, group=None):
    """variable all-to-all, several buffers at once: send_lists[k][p] -> rank p, recv_lists[k][p] <- rank p
    (tensors of matching sizes). ONE grouped launch of point-to-point operations (ncclSend/ncclRecv under
    the nccl backend); returns the work handles - the transfer proceeds while the caller enqueues kernels."""
    ops = []
    for send, recv in zip(send_lists, recv_lists):
        recv[rank].copy_(send[rank])
        for p in range(world):
            if p == rank:
                continue
            if send[p].numel():
                ops.append(dist.P2POp(dist.isend, send[p], p, group=group))
            if recv[p].numel():
                ops.append(dist.P2POp(dist.irecv, recv[p], p, group=group))
    return dist.batch_isend_irecv(ops) if ops else []


def _exchange_wait(works):
    for w in works:
        w.wait()


def _all_counts(counts, device, group=None):
    """all-gather every rank's per-destination count vector -> [world][world] python ints"""
    world = dist.get_world_size(group)
    mine = torch.tensor(counts, dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    return torch.stack(allc).cpu().tolist()


class PartitionedSelfCollision:
    """Per-rank driver of the partitioned build + query over torch.distributed (one process per GPU).

    peer_memory=True (default for world > 1): the two data exchanges are not NCCL calls but part of the
    kernels that produce the data - the range-partition kernel stores every (key, id) straight into the
    owning rank's buffer and the ghost-selection kernel appends records to the peers' ghost buffers, both
    through CUDA-IPC mapped peer memory over NVLink. NCCL carries only the small collectives (histogram
    all-reduce, count / box all-gathers, 1-word barriers) and the final gather of the pair lists."""

    def __init__(self, cd, ctx, mesh, params, group=None, slack=1.5, peer_memory=True):
        self.cd, self.ctx, self.group = cd, ctx, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", ctx.device)
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.part = PartitionedRank(cd, ctx, mesh, params, self.rank, self.world, slack)
        self.counts = [0] * self.world
        self.stats = {}
        self.peer_memory = bool(peer_memory) and self.world > 1
        self._mapped = []
        if self.peer_memory:
            # every rank must end up in the same mode: if CUDA IPC is unavailable anywhere (e.g. a container
            # without the needed permissions) all ranks fall back to the NCCL send/recv exchange
            ok, err = 1, ""
            try:
                handles, offsets = ctx.ipc_export(self.part.bvh)
            except Exception as e:  # noqa: BLE001
                ok, err, handles, offsets = 0, str(e), b"", []
            everyone = [None] * self.world
            dist.all_gather_object(everyone, (ok, handles, offsets), group=group)
            ok = int(all(e[0] for e in everyone))
            peers = [0] * (4 * self.world)
            if ok:
                try:
                    for r, (_, h, off) in enumerate(everyone):
                        if r == self.rank:
                            continue
                        for i in range(4):
                            base = ctx.ipc_open(h[64 * i:64 * i + 64])
                            self._mapped.append(base)
                            peers[4 * r + i] = base + off[i]
                    ctx.bvh_set_peers(self.part.bvh, self.world, self.rank, peers)
                except Exception as e:  # noqa: BLE001
                    ok, err = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                if self.rank == 0:
                    import sys
                    print(f"[b200cd] CUDA IPC peer memory unavailable ({err or 'on another rank'}): using NCCL send/recv",
                          file=sys.stderr)
                self.close()
                self.peer_memory = False
            else:
                self._counts_dev = torch.zeros(self.world, dtype=torch.int32, device=self.device)
                self._ghist = torch.zeros(65536, dtype=torch.int32, device=self.device)
                self._splitters = torch.zeros(max(self.world - 1, 1), dtype=torch.int64, device=self.device)
                self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
                dist.barrier(group=group)

    def _barrier(self):
        dist.all_reduce(self._token, group=self.group)  # stream-ordered: everybody's preceding kernels have finished

    def close(self):
        for base in self._mapped:
            self.ctx.ipc_close(base)
        self._mapped = []

    def _step_peer_memory(self, profile):
        import time
        p, r, w, g, ctx = self.part, self.rank, self.world, self.group, self.ctx
        marks = []

        def mark(name):
            if profile:
                torch.cuda.synchronize(self.device)
                marks.append((name, time.perf_counter()))
        mark("start")
        ctx.ghost_counter_reset(p.bvh)
        hist = p.keys_and_histogram()                                  # my own keys per top-bits bin
        ghist = self._ghist
        ghist.copy_(hist)
        dist.all_reduce(ghist, group=g)  # also orders every rank's counter reset before any ghost append of this step
        mark("keys+hist+allreduce")
        # splitters from the global histogram and - they sit on bin boundaries - how many of MY keys each rank owns from
        # my own histogram, in one launch (b200cd_partition_plan_device; PartitionedRank.splitters_from is the torch twin)
        splitters = self._splitters
        ctx.partition_plan_device(ghist.data_ptr(), hist.data_ptr(), p.shift, w, splitters.data_ptr(), self._counts_dev.data_ptr())
        allc = torch.empty(w * w, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(allc, self._counts_dev, group=g)
        allc = allc.view(w, w)                                         # [source rank][owner rank]
        recv_off = allc[:r].sum(0, dtype=torch.int32).contiguous()     # my segment's start in every owner's buffer
        totals = allc.sum(0).tolist()                                  # (every rank sees every range's size: all raise together)
        # read BEFORE the partition kernel is launched: the host then waits for the count all-gather only and the
        # partition kernel, the barrier and the build's kernels are enqueued back to back behind it
        p.nlocal = int(totals[r])
        if max(totals) > p.cap:
            raise RuntimeError(f"a Morton range holds {max(totals)} triangles, capacity {p.cap}: raise slack (very uneven mesh)")
        ctx.partition_to_peers_device(p.bvh, p.keys.data_ptr(), p.lo, p.cnt, splitters.data_ptr(), w - 1, recv_off.data_ptr())
        self._barrier()                                                # all (key, id) stores have landed
        mark("partition+exchange (fused)")
        boxes = p.build()
        mark("build")
        allb = torch.empty(w * boxes.numel(), dtype=boxes.dtype, device=self.device)
        dist.all_gather_into_tensor(allb, boxes, group=g)
        mask = 0
        for q in range(r + 1, w):
            mask |= 1 << q
        ctx.send_ghosts_to_peers_device(p.bvh, allb.data_ptr(), w, PARTITION_BOXES, mask)  # appends travel during the local query
        mark("ghost select+send (fused)")
        p.collide_local()
        mark("local query")
        self._barrier()                                                # all ghost appends have landed
        p.nghost = ctx.ghost_counter_read(p.bvh)
        local = p.collide_ghosts()
        mark("ghost query")
        self.stats = {"local_triangles": p.nlocal, "ghosts": p.nghost, "local_pairs": int(local.numel()), "peer_memory": True}
        id_bits = max(1, int(p.n - 1).bit_length())
        buf, self.counts = gather_pairs_padded(local, id_bits, g)
        merged = None
        if r == 0:
            if buf.numel() > 1:
                ctx.sort_pairs_device(buf.data_ptr(), buf.numel(), id_bits=id_bits)  # sentinels end up behind the pairs
            merged = buf[:sum(self.counts)]
        mark("gather+sort")
        if profile:
            self.stats["phase_ms"] = {b[0]: round(1e3 * (b[1] - a[1]), 3) for a, b in zip(marks, marks[1:])}
        return merged

    def step(self, profile=False):
        """one distributed build + query; rank 0 gets the sorted packed pair list (device tensor).
        profile=True synchronises after every phase and records wall-clock phase times in self.stats."""
        import time
        if self.peer_memory:
            return self._step_peer_memory(profile)
        p, r, w, g = self.part, self.rank, self.world, self.group
        marks = []

        def mark(name):
            if profile:
                torch.cuda.synchronize(self.device)
                marks.append((name, time.perf_counter()))
        mark("start")
        hist = p.keys_and_histogram()
        if w > 1:
            dist.all_reduce(hist, group=g)
        mark("keys+hist+allreduce")
        counts, pieces = p.partition(hist)
        mark("partition")
        allc = _all_counts(counts, self.device, g) if w > 1 else [counts]
        biggest = max(sum(allc[src][dst] for src in range(w)) for dst in range(w))
        if biggest > p.cap:  # same decision on every rank
            raise RuntimeError(f"a Morton range holds {biggest} triangles, capacity {p.cap}: raise slack (very uneven mesh)")
        views = p.key_recv_views([allc[src][r] for src in range(w)])
        if w > 1:
            _exchange_wait(_exchange_start(r, w, [[k for k, _ in pieces], [i for _, i in pieces]],
                                           [[k for k, _ in views], [i for _, i in views]], g))
        else:
            views[0][0].copy_(pieces[0][0])
            views[0][1].copy_(pieces[0][1])
        mark("key exchange")
        boxes = p.build()
        mark("build")
        works = []
        if w > 1:
            allb = torch.empty(w * boxes.numel(), dtype=boxes.dtype, device=self.device)
            dist.all_gather_into_tensor(allb, boxes, group=g)
            gcounts, gpieces = p.select_ghosts(allb)
            gall = _all_counts(gcounts, self.device, g)
            gviews = p.ghost_recv_views([gall[src][r] for src in range(w)])
            works = _exchange_start(r, w, [gpieces], [gviews], g)  # travels while the local query runs
        mark("ghost select")
        p.collide_local()
        mark("local query")
        _exchange_wait(works)
        local = p.collide_ghosts()
        mark("ghost query")
        self.stats = {"local_triangles": p.nlocal, "ghosts": p.nghost, "local_pairs": int(local.numel())}
        if w == 1:
            if local.numel() > 1:
                self.ctx.sort_pairs_device(local.data_ptr(), local.numel(), id_bits=max(1, int(p.n - 1).bit_length()))
            self.counts = [int(local.numel())]
            return local
        merged, self.counts = gather_pairs(local, 0, g)
        if r == 0 and merged.numel() > 1:
            self.ctx.sort_pairs_device(merged.data_ptr(), merged.numel(), id_bits=max(1, int(p.n - 1).bit_length()))
        mark("gather+sort")
        if profile:
            self.stats["phase_ms"] = {b[0]: round(1e3 * (b[1] - a[1]), 3) for a, b in zip(marks, marks[1:])}
        return merged


class DistSelfCollision:
    """The partitioned step driven by the LIBRARY (b200cd_dist_step, csrc/dist.cu): per frame one C call per rank; the
    ranks talk through CUDA-IPC peer memory and flag barriers. torch.distributed is used once, to move the export
    blobs between the processes, and for the caller's own barriers."""

    def __init__(self, cd, ctx, mesh, params, group=None, slack=1.5, pair_capacity=0, async_sort=False):
        self.cd, self.ctx, self.mesh, self.params, self.group = cd, ctx, mesh, params, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.device("cuda", ctx.device)
        ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.dist = ctx.dist_create(self.rank, self.world, mesh.ntris, slack, pair_capacity)
        if self.world > 1:
            blobs = [None] * self.world
            dist.all_gather_object(blobs, self.dist.export(), group=group)
            self.dist.connect(blobs)
            dist.barrier(group=group)  # nobody steps before everybody has mapped everybody
        self.counts = [0] * self.world
        # pipelined frames: rank 0 sorts frame k's list on a side stream while every rank starts on frame k + 1
        self.async_sort = bool(async_sort)
        if self.async_sort:
            self.dist.set_async_sort(True)

    def step(self, wait=False):
        """one distributed build + query; rank 0 gets the sorted packed pair list (device tensor), the others None.
        With async_sort the tensor's contents are valid after wait_sorted() (or pass wait=True)"""
        ptr, count = self.dist.step(self.mesh, self.params)
        if wait:
            self.dist.wait_sorted()
        if self.rank != 0:
            return None
        return device_pairs_as_tensor(ptr, count, self.device)

    def wait_sorted(self):
        self.dist.wait_sorted()

    @property
    def stats(self):
        return self.dist.stats()

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)  # peers may still be storing into my buffers before this
        self.dist.destroy()


def partitioned_self_collision_emulated(cd, ctx, mesh, params, world, slack=1.5):
    """The same algorithm with `world` ranks emulated one after the other on ONE GPU (exchanges are
    device copies). Used by the single-GPU tests of the multi-rank logic; returns (sorted (count, 2)
    uint32 pairs, per-rank stats)."""
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    ranks = [PartitionedRank(cd, ctx, mesh, params, r, world, slack) for r in range(world)]
    hist = torch.zeros(65536, dtype=torch.int32, device=dev)
    for p in ranks:
        hist += p.keys_and_histogram()
    parts = [p.partition(hist) for p in ranks]
    for dst, p in enumerate(ranks):
        views = p.key_recv_views([parts[src][0][dst] for src in range(world)])
        for src in range(world):
            views[src][0].copy_(parts[src][1][dst][0])
            views[src][1].copy_(parts[src][1][dst][1])
    allb = torch.cat([p.build().clone() for p in ranks])
    ghosts = [p.select_ghosts(allb) for p in ranks]
    for dst, p in enumerate(ranks):
        views = p.ghost_recv_views([ghosts[src][0][dst] for src in range(world)])
        for src in range(world):
            views[src].copy_(ghosts[src][1][dst])
    for p in ranks:
        p.collide_local()
    lists = [p.collide_ghosts().clone() for p in ranks]
    merged = torch.cat(lists)
    if merged.numel() > 1:
        ctx.sort_pairs_device(merged.data_ptr(), merged.numel(), id_bits=max(1, int(mesh.ntris - 1).bit_length()))
    torch.cuda.synchronize()
    stats = [{"local_triangles": p.nlocal, "ghosts": p.nghost, "pairs": int(l.numel())} for p, l in zip(ranks, lists)]
    for p in ranks:
        p.bvh.destroy()
    return unpack_pairs(merged), stats


def upload_mesh_sharded(ctx, mesh, xyz_host_ptr, idx_host_ptr, group=None):
    """Host -> all ranks: every rank pushes 1/world of the mesh over its own PCIe link, then the ranks
    all-gather the device buffers in place over NVLink (instead of world copies of the whole mesh
    crossing the host). xyz_host_ptr / idx_host_ptr: the WHOLE mesh in (pinned) host memory of this rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", ctx.device)
    V, N = mesh.nverts, mesh.ntris
    cv, ct = (V + world - 1) // world, (N + world - 1) // world
    v0, t0 = min(cv * rank, V), min(ct * rank, N)
    nv, nt = min(cv, V - v0), min(ct, N - t0)
    mesh.update_slice_from_ptr(xyz_host_ptr + 12 * v0, v0, nv, idx_host_ptr + 12 * t0, t0, nt)
    dv, di = mesh.device_buffers()
    verts = device_view(dv, 4 * cv * world, "<f4", dev)
    idx = device_view(di, 3 * ct * world, "<i4", dev)
    dist.all_gather_into_tensor(verts, verts[4 * cv * rank:4 * cv * (rank + 1)], group=group)
    dist.all_gather_into_tensor(idx, idx[3 * ct * rank:3 * ct * (rank + 1)], group=group)
    return 12 * nv + 12 * nt


class PeerMeshFrames:
    """Double-buffered mesh frames on several GPUs (one process per GPU).

    Every rank holds the same list of mesh objects (normally two). upload_async(k, ...) sends THIS rank's
    1/world slice of a frame host -> own GPU over its own PCIe link and from there into the peers' copies of
    mesh k with the copy engines over NVLink (CUDA-IPC peer memory), on the context's copy stream - no NCCL, no SMs -
    so frame k+1 travels while frame k is built and queried. wait(k) blocks until my slice has landed
    everywhere and then passes a barrier, after which the whole frame is in place on every rank.
    `ok` is False when CUDA IPC is unavailable on any rank (use upload_mesh_sharded then)."""

    def __init__(self, cd, ctx, meshes, group=None):
        self.cd, self.ctx, self.meshes, self.group = cd, ctx, list(meshes), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", ctx.device)
        self._mapped = []
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
        ok, err, mine = 1, "", []
        try:
            mine = [m.ipc_export() for m in self.meshes]
        except Exception as e:  # noqa: BLE001
            ok, err = 0, str(e)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (ok, mine), group=group)
        ok = int(all(e[0] for e in everyone))
        if ok:
            try:
                for k, m in enumerate(self.meshes):
                    peers = [0] * (2 * self.world)
                    for r, (_, exported) in enumerate(everyone):
                        if r == self.rank:
                            continue
                        handles, offsets = exported[k]
                        for i in range(2):
                            base = ctx.ipc_open(handles[64 * i:64 * i + 64])
                            self._mapped.append(base)
                            peers[2 * r + i] = base + offsets[i]
                    m.set_peers(self.world, self.rank, peers)
            except Exception as e:  # noqa: BLE001
                ok, err = 0, str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(int(flag.item()))
        self.error = err
        if not self.ok:
            self.close()

    def slice_of(self, mesh):
        V, N = mesh.nverts, mesh.ntris
        cv, ct = (V + self.world - 1) // self.world, (N + self.world - 1) // self.world
        v0, t0 = min(cv * self.rank, V), min(ct * self.rank, N)
        return v0, min(cv, V - v0), t0, min(ct, N - t0)

    def upload_async(self, k, xyz_host_ptr, idx_host_ptr):
        """my slice of frame k: host -> my GPU -> every peer (returns at once). xyz_host_ptr / idx_host_ptr: the WHOLE
        frame in pinned host memory of this rank. Returns the bytes this rank pushes over PCIe."""
        m = self.meshes[k]
        v0, nv, t0, nt = self.slice_of(m)
        if not idx_host_ptr:  # vertex positions only (the topology of a deforming mesh does not change)
            m.update_slice_async_from_ptr(xyz_host_ptr + 12 * v0, v0, nv, None, 0, 0)
            return 12 * nv
        m.update_slice_async_from_ptr(xyz_host_ptr + 12 * v0, v0, nv, idx_host_ptr + 12 * t0, t0, nt)
        return 12 * nv + 12 * nt

    def wait(self, k):
        """frame k is complete on every rank when this returns (host-blocking on my own copies, then a barrier)"""
        self.meshes[k].wait()
        dist.all_reduce(self._token, group=self.group)
        return self.meshes[k]

    def close(self):
        for base in self._mapped:
            self.ctx.ipc_close(base)
        self._mapped = []
