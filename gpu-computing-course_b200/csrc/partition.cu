// partition.cu — kernels of the PARTITIONED multi-GPU build (one Morton range of triangles per GPU).
//
// The reference is single-GPU and has nothing like this. With the BVH
// replicated, every GPU repeats the whole build and 8 GPUs cannot beat ~2x; here each rank builds
// and queries only its own Morton range and exchanges the thin layer of triangles that can touch
// another rank's range ("ghosts"):
//   key_hist16_kernel   65536-bin histogram of the keys' top bits -> all-reduce -> range splitters
//   (radix_partition, radix_sort.cu: stable bucketing of (key, id) by splitter)
//   chunk_box_kernel    K coarse boxes per rank: AABBs of K equal runs of its sorted leaves
//   ghost_kernel        leaves whose box overlaps any coarse box of a peer -> that peer's ghost list
// The pair SET is unchanged: {a,b} with both triangles on one rank is found by that rank's own
// query; a cross pair is found exactly once, by the higher rank, when the lower rank's triangle
// arrives there as a ghost query (its box overlaps the partner's box, hence the partner's coarse box).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b200cd {

namespace {

__global__ void __launch_bounds__(256) key_hist16_kernel(const uint64_t* __restrict__ keys, uint32_t n, int shift,
                                                        uint32_t* __restrict__ hist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)min(__ldg(keys + i) >> shift, (uint64_t)0xffffull);  // keys beyond 2^(shift+16) share the last bin
        // clustered meshes put whole warps into one bin: aggregate before touching L2
        const uint32_t peers = __match_any_sync(__activemask(), b);
        if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(hist + b, (uint32_t)__popc(peers));
    }
}

// Coarse boxes of a rank = a CUT through its tree: the subtrees with at most T leaves whose parent
// has more than T. They are disjoint, cover every leaf and - being subtrees of a Morton radix tree -
// are octree cells, so their boxes are tight (a fixed run of consecutive leaves can straddle a jump of
// the Morton curve and balloon). One thread per internal node: a node with more than T leaves emits
// the children that have at most T; the child's box sits right there in pairs[s]. If a degenerate
// tree yields more than K cut nodes the surplus is merged into the last box (looser, still covering).
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void cut_box_init_kernel(uint32_t* __restrict__ boxes_ord, uint32_t K) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K * 6) boxes_ord[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
    else if (i == K * 6) boxes_ord[i] = 0u;  // cut-node counter
}

// boxes_ord: K x 6 order-preserving uint images of floats (lo xyz init 0xffffffff, hi xyz init 0)
__global__ void __launch_bounds__(256)
cut_box_kernel(const NodePair* __restrict__ pairs, const float* __restrict__ root_box, uint32_t n, uint32_t T, uint32_t K,
               uint32_t* __restrict__ boxes_ord, uint32_t* __restrict__ counter) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n - 1) return;
    const Node32 l = pairs[s].c[0], r = pairs[s].c[1];
    const uint32_t F = (uint32_t)l.ext, L = (uint32_t)r.ext;
    if (L - F + 1 <= T) return;  // not above the cut
    const uint32_t size[2] = {s - F + 1, L - s};
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        if (size[side] > T) continue;
        const Node32& c = side ? r : l;
        const uint32_t slot = min(atomicAdd(counter, 1u), K - 1);
        uint32_t* b = boxes_ord + 6 * (size_t)slot;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(b + a, f2ord(c.lo[a]));
            atomicMax(b + 3 + a, f2ord(c.hi[a]));
        }
    }
}

__global__ void cut_box_finish_kernel(const uint32_t* __restrict__ boxes_ord, const float* __restrict__ root_box, uint32_t n,
                                      uint32_t T, uint32_t K, float* __restrict__ boxes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * 6) return;
    float v = ord2f(boxes_ord[i]);                 // untouched slots decode to lo = NaN-free +max / hi = -max patterns:
    if (boxes_ord[i] == ((i % 6) < 3 ? 0xffffffffu : 0u)) v = (i % 6) < 3 ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
    if (n <= T && i < 6) v = root_box[i];          // the whole tree is below the cut: one box, the root's
    boxes[i] = v;
}

constexpr int GH_MAXK = 256;

// The same cut, found from the top: ONE block walks the nodes that hold more than T leaves (they form the top of the
// tree: at most n / T ~ K / 2 disjoint ones per level, one dependent 64-byte fetch per level, ~10 levels) instead of
// every node being read to find out whether it is one of them (64 B x (n - 1): 0.32 ms at 2^25 leaves).
// boxes: K x 6 floats; surplus cut nodes of a degenerate tree are merged into the last box.
constexpr int CUT_THREADS = 256;
__global__ void __launch_bounds__(CUT_THREADS)
cut_box_topdown_kernel(const NodePair* __restrict__ pairs, const float* __restrict__ root_box, uint32_t n, uint32_t T, uint32_t K,
                       float* __restrict__ boxes) {
    __shared__ int s_front[2][GH_MAXK];
    __shared__ uint32_t s_nfront[2], s_nout;
    __shared__ uint32_t s_last[6];  // order-preserving images: surplus boxes are merged here
    const uint32_t tid = threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    for (uint32_t i = tid; i < K * 6; i += CUT_THREADS) boxes[i] = (i % 6) < 3 ? inf : -inf;
    if (tid < 6) s_last[tid] = tid < 3 ? 0xffffffffu : 0u;
    if (tid == 0) {
        s_nfront[0] = s_nfront[1] = 0;
        s_nout = 0;
        if (n > 1 && n > T) {
            s_front[0][0] = reinterpret_cast<const int*>(root_box)[6];
            s_nfront[0] = 1;
        }
    }
    __syncthreads();
    if (n <= T || n <= 1) {  // the whole tree is below the cut: one box, the root's
        if (tid < 6 && n) boxes[tid] = root_box[tid];
        return;
    }
    int cur = 0;
    while (s_nfront[cur]) {
        const uint32_t cnt = s_nfront[cur];
        for (uint32_t i = tid; i < cnt; i += CUT_THREADS) {
            const int sidx = s_front[cur][i];
            const Node32 l = pairs[sidx].c[0], r = pairs[sidx].c[1];
            const uint32_t size[2] = {(uint32_t)sidx - (uint32_t)l.ext + 1u, (uint32_t)r.ext - (uint32_t)sidx};
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const Node32& c = side ? r : l;
                if (size[side] > T && c.link >= 0) {  // still above the cut (at most n / T of these are alive at once)
                    const uint32_t at = atomicAdd(&s_nfront[cur ^ 1], 1u);
                    if (at < (uint32_t)GH_MAXK) { s_front[cur ^ 1][at] = c.link; continue; }
                    // (cannot happen for T >= 2n / K; if it does the subtree is emitted whole: looser, still covering)
                }
                const uint32_t slot = atomicAdd(&s_nout, 1u);
                if (slot + 1 < K) {
                    float* b = boxes + 6 * (size_t)slot;
                    b[0] = c.lo[0]; b[1] = c.lo[1]; b[2] = c.lo[2]; b[3] = c.hi[0]; b[4] = c.hi[1]; b[5] = c.hi[2];
                } else {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        atomicMin(&s_last[a], f2ord(c.lo[a]));
                        atomicMax(&s_last[3 + a], f2ord(c.hi[a]));
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_nfront[cur] = 0;
            if (s_nfront[cur ^ 1] > (uint32_t)GH_MAXK) s_nfront[cur ^ 1] = GH_MAXK;
        }
        cur ^= 1;
        __syncthreads();
    }
    if (tid < 6 && s_nout + 1 > K) boxes[6 * (size_t)(K - 1) + tid] = ord2f(s_last[tid]);
}

constexpr int GH_GROUP = 16;  // coarse boxes per super box

// Each thread owns one local leaf; for every selected peer it tests the leaf's box (strict overlap,
// like the traversal) against that peer's coarse boxes - first the peer's overall box, then super
// boxes of 16 consecutive coarse boxes, then the coarse boxes of the super boxes it overlaps - and,
// on the first hit, appends the leaf's 64-byte record to the peer's ghost list (warp-aggregated atomic).
// ---- ghost selection for the C++ multi-GPU step, in two kernels over a list (the tree build left the union box of
// every 256-leaf block):
//   ghost_block_filter_kernel   ONE WARP per block: the block's box against the overall box of every selected peer and
//                               then, 32 at a time, against the peer's K <= 256 coarse boxes; a (block, peer) pair that
//                               overlaps any of them becomes a list item carrying the 256-bit mask of those boxes.
//                               On a Morton-range partition all blocks but those next to a range boundary drop out here.
//   ghost_listed_kernel         one CTA per item: every leaf of the block against the (few) coarse boxes of the mask -
//                               uniform loads, no shared memory, no barriers - and the hits are appended to the peer's
//                               ghost records (one remote atomic per warp + 256-bit stores over NVLink).
// Before: every block started a CTA, loaded 6 KB of peer boxes per overlapping peer into shared memory and mostly found
// nothing (0.2 - 0.3 ms per step on some ranks of 8 on the two sheets, whose neighbouring ranges' overall boxes overlap
// almost entirely); a one-thread-per-block pre-filter walking 256 boxes serially still cost 0.06 - 0.13 ms.
struct GhostItem {
    uint32_t block, peer;
    uint32_t mask[8];  // bit (k & 31) of word (k >> 5): coarse box k of the peer overlaps the block's box
};

__global__ void __launch_bounds__(256)
ghost_block_filter_kernel(const float* __restrict__ block_boxes, uint32_t nblocks, const float* __restrict__ peer_overall,
                          const float* __restrict__ peer_boxes, uint32_t K, uint32_t npeers, uint32_t peer_mask,
                          GhostItem* __restrict__ items, uint32_t* __restrict__ count, uint32_t cap) {
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nblocks) return;  // warp-uniform
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(block_boxes + 8 * (size_t)b));
    const float4 u1 = __ldg(reinterpret_cast<const float4*>(block_boxes + 8 * (size_t)b) + 1);
    for (uint32_t p = 0; p < npeers; ++p) {
        if (!((peer_mask >> p) & 1u)) continue;
        const float* ob = peer_overall + 6 * (size_t)p;
        if (!(u0.x < __ldg(ob + 3) && __ldg(ob) < u0.w && u0.y < __ldg(ob + 4) && __ldg(ob + 1) < u1.x && u0.z < __ldg(ob + 5) &&
              __ldg(ob + 2) < u1.y))
            continue;
        uint32_t m[8], any = 0;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            const uint32_t k = 32 * j + lane;
            bool hit = false;
            if (k < K) {
                const float* c = peer_boxes + ((size_t)p * K + k) * 6;
                hit = u0.x < __ldg(c + 3) && __ldg(c) < u0.w && u0.y < __ldg(c + 4) && __ldg(c + 1) < u1.x && u0.z < __ldg(c + 5) &&
                      __ldg(c + 2) < u1.y;
            }
            m[j] = __ballot_sync(0xffffffffu, hit);
            any |= m[j];
        }
        if (any && lane == 0) {
            const uint32_t slot = atomicAdd(count, 1u);
            if (slot < cap) {  // (cap = blocks x peers: cannot overflow)
                GhostItem it;
                it.block = b;
                it.peer = p;
#pragma unroll
                for (int j = 0; j < 8; ++j) it.mask[j] = m[j];
                items[slot] = it;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
ghost_listed_kernel(const LeafRec* __restrict__ leaves, uint32_t n, const float* __restrict__ peer_boxes, uint32_t K,
                    const GhostItem* __restrict__ items, const uint32_t* __restrict__ count, uint32_t cap,
                    const PeerTable* __restrict__ peers) {
    const uint32_t nitems = min(*count, cap);
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
        const GhostItem it = items[item];  // the same words for the whole CTA
        const uint32_t j = it.block * blockDim.x + threadIdx.x;
        const bool valid = j < n;
        float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        float4 r0, r1, r2, r3;
        if (valid) {
            ld256_nc(leaves + j, r0, r1);
            ld256_nc(reinterpret_cast<const float4*>(leaves + j) + 2, r2, r3);
            lo[0] = fminf(fminf(r0.x, r0.w), r1.z); hi[0] = fmaxf(fmaxf(r0.x, r0.w), r1.z);
            lo[1] = fminf(fminf(r0.y, r1.x), r1.w); hi[1] = fmaxf(fmaxf(r0.y, r1.x), r1.w);
            lo[2] = fminf(fminf(r0.z, r1.y), r2.x); hi[2] = fmaxf(fmaxf(r0.z, r1.y), r2.x);
        }
        bool hit = false;
#pragma unroll
        for (uint32_t w = 0; w < 8; ++w) {
            uint32_t mm = it.mask[w];
            while (mm) {  // CTA-uniform
                const uint32_t k = 32 * w + (uint32_t)(__ffs(mm) - 1);
                mm &= mm - 1;
                const float* c = peer_boxes + ((size_t)it.peer * K + k) * 6;
                hit = hit || (valid && lo[0] < __ldg(c + 3) && __ldg(c) < hi[0] && lo[1] < __ldg(c + 4) && __ldg(c + 1) < hi[1] &&
                              lo[2] < __ldg(c + 5) && __ldg(c + 2) < hi[2]);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            unsigned long long base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(peers->ghost_count[it.peer], (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
            if (hit && base < peers->ghost_cap) {
                LeafRec* dst = peers->ghosts[it.peer] + base;
                st256(dst, r0, r1);
                st256(reinterpret_cast<float4*>(dst) + 2, r2, r3);
            }
        }
    }
}

// Each thread owns one local leaf; for every selected peer it tests the leaf's box (strict overlap,
// like the traversal) against that peer's coarse boxes - first the peer's overall box, then super
// boxes of 16 consecutive coarse boxes, then the coarse boxes of the super boxes it overlaps - and,
// on the first hit, appends the leaf's 64-byte record to the peer's ghost list (warp-aggregated atomic).
template <bool REMOTE>
__global__ void __launch_bounds__(256)
ghost_kernel(const LeafRec* __restrict__ leaves, uint32_t n, const float* __restrict__ peer_boxes, uint32_t npeers,
             uint32_t K, uint32_t peer_mask, LeafRec* __restrict__ ghosts, uint64_t cap_per_peer,
             unsigned long long* __restrict__ counts, const PeerTable* __restrict__ peers,
             const float* __restrict__ peer_overall, const float* __restrict__ block_boxes) {
    __shared__ float s_red[8][6];
    __shared__ float s_union[6];
    const uint32_t blk = blockIdx.x;
    if (block_boxes) {
        if (threadIdx.x < 6) s_union[threadIdx.x] = __ldg(block_boxes + 8 * (size_t)blk + threadIdx.x);
        __syncthreads();
        bool any = false;
        for (uint32_t p = 0; p < npeers; ++p) {
            if (!((peer_mask >> p) & 1u)) continue;
            const float* ob = peer_overall + 6 * (size_t)p;
            any = any || (s_union[0] < __ldg(ob + 3) && __ldg(ob) < s_union[3] && s_union[1] < __ldg(ob + 4) &&
                          __ldg(ob + 1) < s_union[4] && s_union[2] < __ldg(ob + 5) && __ldg(ob + 2) < s_union[5]);
        }
        if (!any) return;  // block-uniform
    }
    __shared__ float s_box[GH_MAXK][6];
    __shared__ float s_sup[GH_MAXK / GH_GROUP + 1][6];  // super boxes; the last used slot + 1 .. : [nsup] = overall box
    const uint32_t j = blk * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const float inf = __int_as_float(0x7f800000);
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    float4 r0, r1, r2, r3;
    const bool valid = j < n;
    if (valid) {
        ld256_nc(leaves + j, r0, r1);
        ld256_nc(reinterpret_cast<const float4*>(leaves + j) + 2, r2, r3);
        lo[0] = fminf(fminf(r0.x, r0.w), r1.z); hi[0] = fmaxf(fmaxf(r0.x, r0.w), r1.z);
        lo[1] = fminf(fminf(r0.y, r1.x), r1.w); hi[1] = fmaxf(fmaxf(r0.y, r1.x), r1.w);
        lo[2] = fminf(fminf(r0.z, r1.y), r2.x); hi[2] = fmaxf(fmaxf(r0.z, r1.y), r2.x);
    }
    const uint32_t nsup = (K + GH_GROUP - 1) / GH_GROUP;
    auto hits = [&](const float* b) {
        return lo[0] < b[3] && b[0] < hi[0] && lo[1] < b[4] && b[1] < hi[1] && lo[2] < b[5] && b[2] < hi[2];
    };
    // union box of the block's 256 consecutive sorted leaves (a compact cell of the Morton curve): a peer whose
    // overall box it does not overlap cannot receive any of them, and the block skips that peer's K boxes
    // altogether - all but the blocks next to a range boundary skip every peer
    if (!block_boxes) {
        float v[6] = {valid ? lo[0] : inf, valid ? lo[1] : inf, valid ? lo[2] : inf,
                      valid ? hi[0] : -inf, valid ? hi[1] : -inf, valid ? hi[2] : -inf};
#pragma unroll
        for (int c = 0; c < 6; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float t = __shfl_xor_sync(0xffffffffu, v[c], o);
                v[c] = c < 3 ? fminf(v[c], t) : fmaxf(v[c], t);
            }
            if (lane == 0) s_red[threadIdx.x >> 5][c] = v[c];
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            const uint32_t c = threadIdx.x;
            float u = s_red[0][c];
            for (int w = 1; w < 8; ++w) u = c < 3 ? fminf(u, s_red[w][c]) : fmaxf(u, s_red[w][c]);
            s_union[c] = u;
        }
        __syncthreads();
    }
    for (uint32_t p = 0; p < npeers; ++p) {
        if (!((peer_mask >> p) & 1u)) continue;  // uniform
        {   // block-uniform: the peer's overall box (peer_overall, 6 floats per peer) against the block's union box
            const float* ob = peer_overall + 6 * (size_t)p;
            const float o0 = __ldg(ob), o1 = __ldg(ob + 1), o2 = __ldg(ob + 2), o3 = __ldg(ob + 3), o4 = __ldg(ob + 4), o5 = __ldg(ob + 5);
            if (!(s_union[0] < o3 && o0 < s_union[3] && s_union[1] < o4 && o1 < s_union[4] && s_union[2] < o5 && o2 < s_union[5]))
                continue;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < K * 6; i += blockDim.x) (&s_box[0][0])[i] = __ldg(peer_boxes + (size_t)p * K * 6 + i);
        __syncthreads();
        if (threadIdx.x < nsup * 6) {  // super box g, component c (empty runs hold lo = +inf, hi = -inf and drop out)
            const uint32_t g = threadIdx.x / 6, c = threadIdx.x % 6;
            float v = c < 3 ? inf : -inf;
            for (uint32_t k = g * GH_GROUP; k < min(K, (g + 1) * GH_GROUP); ++k)
                v = c < 3 ? fminf(v, s_box[k][c]) : fmaxf(v, s_box[k][c]);
            s_sup[g][c] = v;
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            const uint32_t c = threadIdx.x;
            float v = c < 3 ? inf : -inf;
            for (uint32_t g = 0; g < nsup; ++g) v = c < 3 ? fminf(v, s_sup[g][c]) : fmaxf(v, s_sup[g][c]);
            s_sup[nsup][c] = v;
        }
        __syncthreads();
        bool hit = false;
        if (valid && hits(s_sup[nsup])) {
            for (uint32_t g = 0; g < nsup && !hit; ++g) {
                if (!hits(s_sup[g])) continue;
                for (uint32_t k = g * GH_GROUP; k < min(K, (g + 1) * GH_GROUP) && !hit; ++k) hit = hits(s_box[k]);
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            // one atomic per warp reserves the slots: on my own list, or (REMOTE) on peer p's ghost counter
            // over NVLink, followed by 256-bit stores into peer p's ghost records
            unsigned long long* ctr = REMOTE ? peers->ghost_count[p] : counts + p;
            const unsigned long long cap = REMOTE ? peers->ghost_cap : (unsigned long long)cap_per_peer;
            unsigned long long base = 0;
            if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(ctr, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
            if (hit && base < cap) {
                LeafRec* dst = REMOTE ? peers->ghosts[p] + base : ghosts + (size_t)p * cap_per_peer + base;
                st256(dst, r0, r1);
                st256(reinterpret_cast<float4*>(dst) + 2, r2, r3);
            }
        }
    }
}

// Range plan of the partitioned build in ONE launch (one block): from the all-reduced 65536-bin histogram of the
// keys' top bits, the world-1 splitters that cut the keys into equal shares - splitter r-1 is the first bin
// boundary at which the cumulative count reaches r * total / world - and, from this rank's own histogram, how
// many of ITS keys each rank will own (the splitters sit on bin boundaries, so no second pass over the keys).
constexpr int PP_THREADS = 1024, PP_BINS = 65536, PP_PER = PP_BINS / PP_THREADS;
__device__ __forceinline__ uint64_t pp_block_exclusive(uint64_t v, uint64_t* s_w, uint64_t& total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    __syncthreads();  // s_w may still be read from a previous call
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint64_t base = 0, tot = 0;
    for (int w = 0; w < PP_THREADS / 32; ++w) {
        const uint64_t x = s_w[w];
        if (w < (int)warp) base += x;
        tot += x;
    }
    total = tot;
    return base + incl - v;
}
// sums of the 1024 segments of 64 consecutive bins, read with coalesced loads (a warp's 32 consecutive bins lie in
// one segment: warp reduction, then one shared-memory atomic per warp and load)
__device__ __forceinline__ void pp_segment_sums(const uint32_t* __restrict__ hist, uint32_t* s_seg) {
    s_seg[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < (uint32_t)PP_BINS; i += PP_THREADS) {
        uint32_t v = __ldg(hist + i);
        v = __reduce_add_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_seg[i / PP_PER], v);
    }
    __syncthreads();
}
__global__ void __launch_bounds__(PP_THREADS)
partition_plan_kernel(const uint32_t* __restrict__ ghist, const uint32_t* __restrict__ lhist, int shift, int world,
                      uint64_t* __restrict__ splitters, int32_t* __restrict__ counts) {
    __shared__ uint64_t s_w[PP_THREADS / 32];
    __shared__ uint32_t s_seg[PP_THREADS];
    __shared__ uint32_t s_bin[RS_MAX_SPLIT_P1];
    __shared__ uint64_t s_lc[RS_MAX_SPLIT_P1];
    const uint32_t tid = threadIdx.x, b0 = tid * PP_PER;  // thread tid owns segment tid = bins [b0, b0 + 64)
    pp_segment_sums(ghist, s_seg);
    const uint64_t sum = s_seg[tid];
    uint64_t total;
    const uint64_t excl = pp_block_exclusive(sum, s_w, total);
    for (int r = 1; r < world; ++r) {
        const uint64_t target = (uint64_t)r * total / (uint64_t)world;
        // first bin whose cumulative count reaches the target (bin 0 for an empty histogram): only the owner of the
        // segment the target falls into walks its 64 bins
        if (target == 0 ? tid == 0 : (excl < target && target <= excl + sum)) {
            uint32_t bin = b0;
            uint64_t run = excl;
            for (int k = 0; k < PP_PER; ++k) {
                run += ghist[b0 + k];
                if (run >= target) { bin = b0 + k; break; }
            }
            s_bin[r - 1] = min(bin, (uint32_t)PP_BINS - 2u);  // keys beyond the histogram's range share the last bin
        }
    }
    __syncthreads();
    pp_segment_sums(lhist, s_seg);
    const uint64_t lsum = s_seg[tid];
    uint64_t ltotal;
    const uint64_t lexcl = pp_block_exclusive(lsum, s_w, ltotal);
    for (int r = 1; r < world; ++r) {
        const uint32_t bin = s_bin[r - 1];
        if (bin / PP_PER == tid) {  // my own keys in the bins [0, bin]
            uint64_t run = lexcl;
            for (uint32_t k = 0; k <= bin % PP_PER; ++k) run += lhist[b0 + k];
            s_lc[r - 1] = run;
        }
    }
    __syncthreads();
    if (tid < (uint32_t)world) {
        const uint64_t hi = tid + 1 < (uint32_t)world ? s_lc[tid] : ltotal, lo = tid ? s_lc[tid - 1] : 0;
        counts[tid] = (int32_t)(hi - lo);
        if (tid + 1 < (uint32_t)world) splitters[tid] = (uint64_t)(s_bin[tid] + 1u) << shift;  // keys >= splitter r-1 belong to rank >= r
    }
}

// ---- range plan of the C++ multi-GPU step (dist.cu). Every rank holds EVERY rank's 65536-bin histogram (pushed into its
// comm block through peer memory), so each rank derives the same splitters and the whole [source][owner] count matrix
// locally - no second exchange for the counts.
// Stage 1 (64 blocks x 1024 bins): global histogram = sum over the ranks, plus per-rank sums of each 1024-bin block.
__global__ void __launch_bounds__(DIST_HIST_BLOCK_BINS)
dist_hist_reduce_kernel(const uint32_t* hists /* [world][65536], written by the peers */, int world, uint32_t* __restrict__ ghist,
                        uint32_t* __restrict__ part /* [world][DIST_HIST_BLOCKS] */, uint32_t* __restrict__ seg /* [1024]: sums of 64 bins */) {
    __shared__ uint32_t s_part[RS_MAX_SPLIT_P1];
    __shared__ uint32_t s_warp[DIST_HIST_BLOCK_BINS / 32];
    if (threadIdx.x < RS_MAX_SPLIT_P1) s_part[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t bin = blockIdx.x * DIST_HIST_BLOCK_BINS + threadIdx.x;
    uint32_t sum = 0;
    for (int src = 0; src < world; ++src) {
        const uint32_t v = hists[(size_t)src * PP_BINS + bin];
        sum += v;
        const uint32_t w = __reduce_add_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_part[src], w);
    }
    ghist[bin] = sum;
    // the plan kernel (ONE block) wants the sums of the 1024 segments of 64 bins: two warps each, computed here by 64 blocks
    // instead of one block reading all 65536 bins again (37 us of every step's critical path)
    const uint32_t wsum = __reduce_add_sync(0xffffffffu, sum);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x < DIST_HIST_BLOCK_BINS / PP_PER)
        seg[blockIdx.x * (DIST_HIST_BLOCK_BINS / PP_PER) + threadIdx.x] = s_warp[2 * threadIdx.x] + s_warp[2 * threadIdx.x + 1];
    if (threadIdx.x < (uint32_t)world) part[threadIdx.x * DIST_HIST_BLOCKS + blockIdx.x] = s_part[threadIdx.x];
}

// Stage 2 (one block): splitter bins from the global histogram (as partition_plan_kernel), then for every
// (source rank, splitter) the number of the source's keys below the splitter - whole 1024-bin blocks from `part`,
// the remainder from the source's histogram, one warp per task - and from those the count matrix, where my segment
// starts in every owner's receive buffer and how many triangles each rank owns.
__global__ void __launch_bounds__(PP_THREADS)
dist_plan_kernel(const uint32_t* __restrict__ ghist, const uint32_t* hists, const uint32_t* __restrict__ part,
                 const uint32_t* __restrict__ seg, int shift, int world, int rank, DistPlan* __restrict__ out) {
    __shared__ uint64_t s_w[PP_THREADS / 32];
    __shared__ uint32_t s_bin[RS_MAX_SPLIT_P1];
    __shared__ uint32_t s_cum[RS_MAX_SPLIT_P1][RS_MAX_SPLIT_P1];  // [src][j]: src's keys in the bins [0, s_bin[j]]; j = world-1: all
    const uint32_t tid = threadIdx.x, b0 = tid * PP_PER;
    const uint64_t sum = seg[tid];  // bins [b0, b0 + 64), summed by dist_hist_reduce_kernel
    uint64_t total;
    const uint64_t excl = pp_block_exclusive(sum, s_w, total);
    for (int r = 1; r < world; ++r) {
        const uint64_t target = (uint64_t)r * total / (uint64_t)world;
        if (target == 0 ? tid == 0 : (excl < target && target <= excl + sum)) {
            uint32_t bin = b0;
            uint64_t run = excl;
            for (int k = 0; k < PP_PER; ++k) {
                run += ghist[b0 + k];
                if (run >= target) { bin = b0 + k; break; }
            }
            s_bin[r - 1] = min(bin, (uint32_t)PP_BINS - 2u);
        }
    }
    __syncthreads();
    const uint32_t warp = tid >> 5, lane = tid & 31;
    for (int task = (int)warp; task < world * world; task += PP_THREADS / 32) {
        const int src = task / world, j = task % world;
        const uint32_t B = j + 1 < world ? s_bin[j] + 1u : (uint32_t)PP_BINS;  // bins below the splitter
        const uint32_t full = B / DIST_HIST_BLOCK_BINS, rem = B % DIST_HIST_BLOCK_BINS;
        uint32_t acc = 0;
        for (uint32_t b = lane; b < full; b += 32) acc += part[src * DIST_HIST_BLOCKS + b];
        for (uint32_t k = lane; k < rem; k += 32) acc += hists[(size_t)src * PP_BINS + full * DIST_HIST_BLOCK_BINS + k];
        acc = __reduce_add_sync(0xffffffffu, acc);
        if (lane == 0) s_cum[src][j] = acc;
    }
    __syncthreads();
    if (tid < (uint32_t)(world * world)) {
        const int src = tid / world, d = tid % world;
        out->counts[src][d] = s_cum[src][d] - (d ? s_cum[src][d - 1] : 0u);
    }
    if (tid < (uint32_t)world) {
        const int d = tid;
        uint32_t off = 0, tot = 0;
        for (int src = 0; src < world; ++src) {
            const uint32_t c = s_cum[src][d] - (d ? s_cum[src][d - 1] : 0u);
            if (src < rank) off += c;
            tot += c;
        }
        out->recv_off[d] = off;
        out->totals[d] = tot;
        if (d + 1 < world) out->splitters[d] = (uint64_t)(s_bin[d] + 1u) << shift;  // keys >= splitter r-1 belong to rank >= r
    }
}

// overall[p] = union of peer p's K coarse boxes (one warp per peer)
__global__ void __launch_bounds__(32)
peer_overall_kernel(const float* __restrict__ peer_boxes, uint32_t K, float* __restrict__ overall) {
    const uint32_t p = blockIdx.x, lane = threadIdx.x;
    const float inf = __int_as_float(0x7f800000);
    float v[6] = {inf, inf, inf, -inf, -inf, -inf};
    for (uint32_t k = lane; k < K; k += 32) {
        const float* b = peer_boxes + ((size_t)p * K + k) * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = c < 3 ? fminf(v[c], __ldg(b + c)) : fmaxf(v[c], __ldg(b + c));
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float t = __shfl_xor_sync(0xffffffffu, v[c], o);
            v[c] = c < 3 ? fminf(v[c], t) : fmaxf(v[c], t);
        }
        if (lane == 0) overall[6 * (size_t)p + c] = v[c];
    }
}

}  // namespace

void launch_partition_plan(const uint32_t* d_ghist, const uint32_t* d_lhist, int shift, int world, uint64_t* d_splitters,
                           int32_t* d_counts, cudaStream_t s) {
    partition_plan_kernel<<<1, PP_THREADS, 0, s>>>(d_ghist, d_lhist, shift, world, d_splitters, d_counts);
    count_launch();
}

void launch_dist_plan(const uint32_t* d_hists, int world, int rank, int shift, uint32_t* d_ghist, uint32_t* d_part,
                      DistPlan* d_plan, cudaStream_t s) {
    static_assert(DIST_HIST_BLOCK_BINS / PP_PER * DIST_HIST_BLOCKS == PP_THREADS && PP_PER == 64, "one segment per plan thread, two warps each");
    uint32_t* d_seg = d_part + RS_MAX_SPLIT_P1 * DIST_HIST_BLOCKS;  // behind the per-rank block sums
    dist_hist_reduce_kernel<<<DIST_HIST_BLOCKS, DIST_HIST_BLOCK_BINS, 0, s>>>(d_hists, world, d_ghist, d_part, d_seg);
    dist_plan_kernel<<<1, PP_THREADS, 0, s>>>(d_ghist, d_hists, d_part, d_seg, shift, world, rank, d_plan);
    count_launch(2);
    trace_mark("dist_hist_reduce+dist_plan", s);
}

void launch_key_hist16(const uint64_t* d_keys, uint32_t n, int shift, uint32_t* d_hist, int sms, cudaStream_t s) {
    if (!n) return;
    const uint32_t blocks = min((n + 255u) / 256u, (uint32_t)sms * 16u);
    key_hist16_kernel<<<blocks, 256, 0, s>>>(d_keys, n, shift, d_hist);
    count_launch();
    trace_mark("key_hist16", s);
}

// d_scratch: K*6 + 1 words
void launch_chunk_boxes(const NodePair* d_pairs, const float* d_root_box, uint32_t n, uint32_t K, uint32_t* d_scratch,
                        float* d_boxes, cudaStream_t s) {
    if (!K) return;
    const uint32_t T = (uint32_t)(((uint64_t)n + K / 2 - 1) / max(K / 2, 1u));  // at most ~K cut nodes on a balanced tree
    static int variant = -1;  // B200CD_CUT=scan: the round-1 version (every node looks at itself; A/B knob, read once)
    if (variant < 0) {
        const char* e = getenv("B200CD_CUT");
        variant = (e && e[0] == 's') ? 1 : 0;
    }
    if (variant == 0 && K <= (uint32_t)GH_MAXK) {
        cut_box_topdown_kernel<<<1, CUT_THREADS, 0, s>>>(d_pairs, d_root_box, n, T ? T : 1u, K, d_boxes);
        count_launch();
        trace_mark("cut_boxes (top-down)", s);
        return;
    }
    uint32_t* counter = d_scratch + 6 * (size_t)K;
    cut_box_init_kernel<<<(K * 6 + 1 + 255) / 256, 256, 0, s>>>(d_scratch, K);
    count_launch();
    if (n > 1) {
        cut_box_kernel<<<(n - 1 + 255) / 256, 256, 0, s>>>(d_pairs, d_root_box, n, T ? T : 1u, K, d_scratch, counter);
        count_launch();
    }
    cut_box_finish_kernel<<<(K * 6 + 255) / 256, 256, 0, s>>>(d_scratch, d_root_box, n, T ? T : 1u, K, d_boxes);
    count_launch();
    trace_mark("cut_boxes (3 kernels)", s);
}

int ghost_max_k() { return GH_MAXK; }

// d_overall: scratch of 6 * npeers floats (the peers' overall boxes)
void launch_ghosts(const LeafRec* d_leaves, uint32_t n, const float* d_peer_boxes, uint32_t npeers, uint32_t K,
                   uint32_t peer_mask, LeafRec* d_ghosts, uint64_t cap_per_peer, unsigned long long* d_counts,
                   float* d_overall, const float* d_block_boxes, cudaStream_t s) {
    cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * npeers, s);
    if (!n || !npeers || !peer_mask) return;
    peer_overall_kernel<<<npeers, 32, 0, s>>>(d_peer_boxes, K, d_overall);
    ghost_kernel<false><<<(n + 255) / 256, 256, 0, s>>>(d_leaves, n, d_peer_boxes, npeers, K, peer_mask, d_ghosts,
                                                       cap_per_peer, d_counts, nullptr, d_overall, d_block_boxes);
    count_launch(2);
}

void launch_ghosts_to_peers(const LeafRec* d_leaves, uint32_t n, const float* d_peer_boxes, uint32_t npeers, uint32_t K,
                            uint32_t peer_mask, const PeerTable* d_peers, float* d_overall, const float* d_block_boxes,
                            cudaStream_t s, uint32_t* d_list, uint32_t list_cap, int sms) {
    if (!n || !npeers || !peer_mask) return;
    peer_overall_kernel<<<npeers, 32, 0, s>>>(d_peer_boxes, K, d_overall);
    if (d_block_boxes && d_list && K <= 256) {  // only the blocks next to a range boundary start any work
        const uint32_t nblocks = (n + 255) / 256;
        uint32_t* d_count = d_list;  // word 0 of the scratch: the list's length; items behind it (16-byte aligned)
        GhostItem* items = reinterpret_cast<GhostItem*>(d_list + 4);
        cudaMemsetAsync(d_count, 0, sizeof(uint32_t), s);
        ghost_block_filter_kernel<<<(nblocks + 7) / 8, 256, 0, s>>>(d_block_boxes, nblocks, d_overall, d_peer_boxes, K, npeers, peer_mask,
                                                                     items, d_count, list_cap);
        trace_mark("ghost block pre-filter", s);
        ghost_listed_kernel<<<std::min<uint32_t>(nblocks, (uint32_t)std::max(sms, 1) * 8u), 256, 0, s>>>(d_leaves, n, d_peer_boxes, K, items,
                                                                                                     d_count, list_cap, d_peers);
        count_launch(3);
        trace_mark("ghost_kernel (block filter + select + send)", s);
        return;
    }
    ghost_kernel<true><<<(n + 255) / 256, 256, 0, s>>>(d_leaves, n, d_peer_boxes, npeers, K, peer_mask, nullptr, 0, nullptr,
                                                      d_peers, d_overall, d_block_boxes);
    count_launch(2);
    trace_mark("ghost_kernel (select + send)", s);
}

}  // namespace b200cd
