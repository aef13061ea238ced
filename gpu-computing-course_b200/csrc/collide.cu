// collide.cu — K5 (broad phase: BVH traversal) and K6 (narrow phase: fp64 SAT).
//
// Reference semantics (under /root/reference/CollisionDetection/):
//   traversal     collision.cuh:19-71  one query triangle per thread; test both
//                 children of the current internal node (box.cuh:40-43, STRICT
//                 overlap); overlapping leaf -> filters -> narrow phase;
//                 overlapping internal -> push; pop. The root box is never tested.
//   filters       triangle.cuh:18-30 (skip when any vertex INDEX is shared),
//                 tri_contact.cuh:81 (report only when a.ID < b.ID)
//   narrow phase  tri_contact.cuh:19-78, vec3f.cuh:118-125,257-291, mathop.cuh:30-44:
//                 17-axis separating-axis test in double, P = lower-ID triangle,
//                 everything translated by P's first vertex.
//
// The emitted SET is a pure function of the mesh; how it is
// found is ours:
//   - every unordered leaf pair {i, j} is discovered ONCE, by the query with the
//     smaller sorted position: a child whose subtree ends at or before the query's
//     own position is pruned. The AABB predicate is symmetric,
//     so this halves the traversal without changing the set; the narrow phase is
//     still called as (lower ID, higher ID) because the SAT is not symmetric in
//     floating point.
//   - queries do not start at the root: one thread per group of 128 consecutive sorted leaves
//     (B200CD_QUERY_GROUP; one warp of the persistent-lane traversal works through one group)
//     walks the group's ancestors once (entry_kernel) and every query of the group starts at
//     the handful of subtrees that tile "leaves after me" and overlap the group's union box.
//   - on triangle soups the walk reads QUANTISED nodes (common.cuh QNodePair: both children in one 32-byte sector,
//     conservative 15-bit cells) - the traversal is bound by the L1 data pipe, i.e. by sectors per node visit - and
//     emits a superset of the exact candidates; the narrow phase applies the exact strict box test first
//     (broad_kernel_q; meshes keep the exact 64-byte nodes: touching boxes defeat a quantiser).
//   - traversal only emits CANDIDATES (AABB-overlapping leaf pairs) into a compact
//     list: per-warp staging in shared memory filled with ballots (no atomics),
//     flushed with one global atomicAdd per >=32 candidates. The divergent fp64
//     SAT runs afterwards in its own kernel, in two warp-dense stages (face-normal axes first,
//     survivors compacted through a per-warp shared-memory queue, then the other 15 axes).
//   - results are appended the same way (warp-aggregated atomic).
// -fmad=false: every double operation rounds separately, like the host reference.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace b200cd {

namespace {

constexpr int BR_THREADS = B200CD_QUERY_BLOCK;  // one block = 256 CONSECUTIVE sorted leaves
constexpr int BR_WARPS = BR_THREADS / 32;
constexpr int BR_FLUSH = 64;    // flush the per-warp candidate staging when at least this many are waiting (<= 64 arrive per step)
constexpr int BR_ENTRIES = B200CD_MAX_ENTRIES;

constexpr unsigned long long ERR_STACK = 1ull;

// strict overlap, box.cuh:40-43: (a.lo - b.hi) * (b.lo - a.hi) > 0 on every axis. For
// well-formed boxes of fp32-exact values that is exactly a.lo < b.hi && b.lo < a.hi.
__device__ __forceinline__ bool overlap(const float qlo[3], const float qhi[3], float lx, float ly, float lz,
                                        float hx, float hy, float hz) {
    return qlo[0] < hx && lx < qhi[0] && qlo[1] < hy && ly < qhi[1] && qlo[2] < hz && lz < qhi[2];
}

__device__ __forceinline__ float min3_ref(float a, float b, float c) { float t = a; if (b < t) t = b; if (c < t) t = c; return t; }
__device__ __forceinline__ float max3_ref(float a, float b, float c) { float t = a; if (b > t) t = b; if (c > t) t = c; return t; }

// which sorted-leaf position does query slot t of this shard own (block-cyclic chunks)
__device__ __forceinline__ uint64_t query_position(uint32_t t, uint32_t shard, uint32_t nshards, uint32_t chunk) {
    const uint32_t c = t / chunk, w = t - c * chunk;
    return ((uint64_t)c * nshards + shard) * chunk + w;
}

struct Child {
    float4 a, b;  // a = lo.xyz, hi.x ; b = hi.yz, link, ext (Node32)
    __device__ __forceinline__ int link() const { return __float_as_int(b.z); }
    __device__ __forceinline__ int ext() const { return __float_as_int(b.w); }
};
__device__ __forceinline__ void load_children(const NodePair* __restrict__ pairs, int node, Child& l, Child& r) {
    ld256_nc(&pairs[node].c[0], l.a, l.b);
    ld256_nc(&pairs[node].c[1], r.a, r.b);
}

// ---------------------------------------------------------------- K5a: entry lists
// Every query of a block [q0, q1] of consecutive sorted leaves needs the leaves j > q whose box
// overlaps its own. Walking down from the root, almost all of a query's node visits are its own
// ancestors (collision.cuh:19-71 starts every query at the root). The ancestors are the same for
// the whole block, so ONE thread per block walks them once and records the maximal subtrees that
// tile the two target sets:
//     inside  (q0, q1]   : the canonical decomposition of the block's own range
//     beyond  (q1, n-1]  : the right siblings along the root -> q1 path
// (at most two root-to-leaf paths, so O(depth) entries). The traversal kernel then starts each
// query at those entries instead of at the root. Entry = the child's Node32 as stored in its
// parent. The emitted candidate SET is unchanged: the entries cover exactly the leaves > q0, the
// per-query rule (subtree end > q, strict box overlap) is applied to each entry and below it.
__global__ void __launch_bounds__(128)
entry_kernel(const NodePair* __restrict__ pairs, const float* __restrict__ root_box, uint32_t n, uint32_t shard,
             uint32_t nshards, uint32_t chunk, uint32_t gsize, uint32_t nblocks, Node32* __restrict__ entries,
             uint32_t* __restrict__ entry_count /* [nblocks] counts, then [nblocks][8] floats: union box of the group's own leaves */) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint64_t p0 = query_position(b * gsize, shard, nshards, chunk);
    const int root = reinterpret_cast<const int*>(root_box)[6];
    uint32_t cnt = 0;
    bool overflow = false;
    Child* out = reinterpret_cast<Child*>(entries + (size_t)b * BR_ENTRIES);
    const float inf = __int_as_float(0x7f800000);
    float ulo[3] = {inf, inf, inf}, uhi[3] = {-inf, -inf, -inf};  // union box of the group's own leaves [q0, q1]
    auto grow = [&](const Child& c) {
        ulo[0] = fminf(ulo[0], c.a.x); ulo[1] = fminf(ulo[1], c.a.y); ulo[2] = fminf(ulo[2], c.a.z);
        uhi[0] = fmaxf(uhi[0], c.a.w); uhi[1] = fmaxf(uhi[1], c.b.x); uhi[2] = fmaxf(uhi[2], c.b.y);
    };
    auto add = [&](Child c, int last, bool inside) {  // entries carry the LAST leaf of their subtree in ext
        c.b.w = __int_as_float(last);
        if (inside) grow(c);  // subtrees inside the group's range tile it (together with leaf q0)
        if (cnt < (uint32_t)BR_ENTRIES) { st256(out + cnt, c.a, c.b); ++cnt; } else overflow = true;
    };
    if (p0 + 1 < n) {  // the very last leaf has no partner after it
        const int q0 = (int)p0;
        const int q1 = (int)min((uint64_t)n - 1, p0 + gsize - 1);
        int node = root, F = 0, L = (int)n - 1;  // current node = split index; its leaf range [F, L]
        Child l, r;
        bool split_found = false;
        // common part of the two paths. Left child = [F, node], right child = [node+1, L].
        while (true) {
            load_children(pairs, node, l, r);
            const int g = node;
            if (q1 <= g) {            // both ends in the left child: the right child lies entirely beyond q1
                add(r, L, false);
                if (l.link() < 0) { grow(l); break; }  // the group is the single leaf q0 = q1
                node = l.link();
                L = g;
            } else if (q0 > g) {      // both ends in the right child: the left child lies entirely before q0
                if (r.link() < 0) { grow(r); break; }
                node = r.link();
                F = g + 1;
            } else {                  // q0 <= g < q1: paths part here
                split_found = true;
                break;
            }
        }
        if (split_found) {
            const int g = node;
            // walk towards q0 inside the left child [F, g]: right siblings lie inside (q0, q1]
            Child cur = l, a, c;
            int curF = F, curL = g;
            while (true) {
                if (curF >= q0 || cur.link() < 0) {  // whole subtree inside the block's range (leaf q0 itself: harmless)
                    add(cur, curL, true);
                    break;
                }
                const int m = cur.link();
                load_children(pairs, m, a, c);
                if (q0 <= m) { add(c, curL, true); cur = a; curL = m; }
                else { curF = m + 1; cur = c; }
            }
            // walk towards q1 inside the right child [g+1, L]: left siblings lie inside the range, right siblings beyond it
            cur = r;
            curL = L;
            while (true) {
                if (curL <= q1 || cur.link() < 0) {
                    add(cur, curL, true);
                    break;
                }
                const int m = cur.link();
                load_children(pairs, m, a, c);
                if (q1 <= m) { add(c, curL, false); cur = a; curL = m; }
                else { add(a, m, true); cur = c; }
            }
        }
        if (overflow) {  // pathologically deep tree: fall back to "start at the root" (two entries = the root's children)
            load_children(pairs, root, l, r);
            cnt = 0;
            overflow = false;
            add(l, root, true);
            add(r, (int)n - 1, true);
        }
    }
    entry_count[b] = cnt;
    float4* ub = reinterpret_cast<float4*>(entry_count + ((nblocks + 7u) & ~7u)) + 2 * (size_t)b;  // 32-byte aligned rows
    st256(ub, make_float4(ulo[0], ulo[1], ulo[2], uhi[0]), make_float4(uhi[1], uhi[2], 0.f, 0.f));
}

// ---------------------------------------------------------------- K5b (variant 2): persistent lanes
// Persistent lanes (after Aila & Laine 2009). With one query per thread the warp loops as long as
// its busiest lane (measured lane utilisation on the 16 M soup: 0.40). Here a block owns 1024
// consecutive queries (eight 128-leaf groups, each with its entry list) and every warp works
// through its own group: whenever at least BR_REFILL lanes have run out of work, the idle
// lanes take the warp's next queries (their records were prefetched at the previous refill), scan
// the group's entry list and rejoin the traversal loop. Lanes of a warp still walk neighbouring
// queries, so node fetches stay coherent. Shared memory is kept small on purpose: the unified
// L1 is what serves ~70 % of the node fetches (a first version that staged the 1024 query boxes in
// 24 KB of shared memory dropped the L1 hit rate from 68 % to 4 % and ran slower, profiles/r01_*).
constexpr int BR_PER_WARP = 128;                       // queries per warp = one group (one entry list)
constexpr int BR_GROUPS = BR_WARPS;                    // groups per block
constexpr int BR_QB = BR_GROUPS * BR_PER_WARP;         // 1024 queries per block
constexpr int BR_REFILL = 24;                          // refill when this many lanes are idle (16: 2.47 ms, 20: 2.44, 24: 2.42 at 6 CTAs/SM)
constexpr int BR_KEEP = 32;                            // start subtrees kept per group after the union-box filter
constexpr int BR_CQ = 128;                             // candidate staging per warp: < 64 carried + <= 64 new per step

// FILTER (meshes with shared vertices): a staged candidate whose two triangles share a vertex INDEX can never be a
// contact (triangle.cuh:18-30, applied at collision.cuh:38 before the narrow phase) - and on a mesh those are almost
// all the candidates (every triangle's ~12 edge / vertex neighbours overlap its box: 6.2 candidates per triangle on the
// two sheets against 0.002 contacts). They are dropped when the warp's staging buffer is flushed: two 16-byte loads per
// candidate (the index halves of the two leaf records), 32 candidates at a time, off the dependent node-fetch chain -
// instead of 8 bytes written, 8 bytes and 2 x 32 bytes read again by the narrow phase for each of them.
template <int MIN_BLOCKS, bool FILTER>  // resident CTAs per SM the register allocation is bounded for (6 -> 40 registers, 48 warps)
__global__ void __launch_bounds__(BR_THREADS, MIN_BLOCKS)
broad_kernel(const NodePair* __restrict__ pairs, const LeafRec* __restrict__ leaves, uint32_t n, uint32_t shard,
             uint32_t nshards, uint32_t chunk, uint32_t nquery, uint32_t ngroups, int refill, const Node32* __restrict__ entries,
             const uint32_t* __restrict__ entry_count, uint2* __restrict__ cand, uint64_t cand_cap,
             unsigned long long* __restrict__ counters) {
    __shared__ uint2 s_cq[BR_WARPS][BR_CQ];
    __shared__ Child s_entry[BR_GROUPS][BR_KEEP];
    __shared__ uint32_t s_nkeep[BR_GROUPS];
    __shared__ uint32_t s_spill[BR_GROUPS];             // a group kept more than BR_KEEP entries: read the rest from global
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint2* wq = s_cq[warp];
    uint32_t staged = 0;  // warp-uniform
    if (tid < BR_GROUPS) { s_nkeep[tid] = 0; s_spill[tid] = 0; }
    const float inf = __int_as_float(0x7f800000);
    const uint32_t qbase = blockIdx.x * BR_QB;         // first query slot of this block
    __syncthreads();

    // query slot t -> sorted leaf position
    auto slot_position = [&](uint32_t t) -> uint32_t {
        if (t >= nquery) return 0xffffffffu;
        const uint64_t qq = query_position(t, shard, nshards, chunk);
        return qq + 1 < n ? (uint32_t)qq : 0xffffffffu;  // the last leaf has no partner with a larger position
    };

    // ---- entry lists: thread (g, e) keeps entry e of group g if it overlaps the group's union box
    for (uint32_t ge = tid; ge < BR_GROUPS * BR_ENTRIES; ge += BR_THREADS) {
        const uint32_t g = ge / BR_ENTRIES, e = ge % BR_ENTRIES;
        const uint32_t eb = blockIdx.x * BR_GROUPS + g;             // the group's entry list
        if (eb < ngroups && e < __ldg(entry_count + eb)) {
            float4 u0, u1;
            ld256_nc(reinterpret_cast<const float4*>(entry_count + ((ngroups + 7u) & ~7u)) + 2 * (size_t)eb, u0, u1);
            const float ulo[3] = {u0.x, u0.y, u0.z}, uhi[3] = {u0.w, u1.x, u1.y};
            Child c;
            ld256_nc(entries + (size_t)eb * BR_ENTRIES + e, c.a, c.b);
            if (overlap(ulo, uhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y)) {
                const uint32_t at = atomicAdd(&s_nkeep[g], 1u);
                if (at < (uint32_t)BR_KEEP) s_entry[g][at] = c; else s_spill[g] = 1;
            }
        }
    }
    __syncthreads();

    int stack[B200CD_MAX_STACK];
    int sp = 0;
    uint32_t overflow = 0;  // a word, not a bool: nvcc packs bools into byte lanes and re-packs them (PRMT) on every path

    // drop the staged candidates whose triangles share a vertex index (compaction in place, warp-wide)
    auto filter = [&]() {
        uint32_t kept = 0;
        for (uint32_t i0 = 0; i0 < staged; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool keep = false;
            uint2 c = make_uint2(0, 0);
            if (i < staged) {
                c = wq[i];
                // second half of a leaf record: v2.z, vi[0], vi[1], vi[2] (16 bytes at offset 32)
                const float4 a = __ldg(reinterpret_cast<const float4*>(leaves + c.x) + 2);
                const float4 b = __ldg(reinterpret_cast<const float4*>(leaves + c.y) + 2);
                const uint32_t a0 = __float_as_uint(a.y), a1 = __float_as_uint(a.z), a2 = __float_as_uint(a.w);
                const uint32_t b0 = __float_as_uint(b.y), b1 = __float_as_uint(b.z), b2 = __float_as_uint(b.w);
                keep = !(a0 == b0 || a0 == b1 || a0 == b2 || a1 == b0 || a1 == b1 || a1 == b2 || a2 == b0 || a2 == b1 || a2 == b2);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);  // (every lane has read its slot: slots below are free to overwrite)
            if (keep) wq[kept + __popc(m & lt)] = c;
            kept += __popc(m);
            __syncwarp();
        }
        staged = kept;
    };
    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counters + 0, (unsigned long long)staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < staged; i += 32)
            if (base + i < cand_cap) __stcs(cand + base + i, wq[i]);
        staged = 0;
        __syncwarp();
    };
    // stage (q, leaf) candidates of the whole warp: positions by ballot, no atomics
    auto stage2 = [&](uint32_t q, bool candL, int leafL, bool candR, int leafR) {
        const uint32_t bL = __ballot_sync(0xffffffffu, candL), bR = __ballot_sync(0xffffffffu, candR);
        if (bL | bR) {
            const uint32_t nL = __popc(bL);
            if (candL) wq[staged + __popc(bL & lt)] = make_uint2(q, (uint32_t)leafL);
            if (candR) wq[staged + nL + __popc(bR & lt)] = make_uint2(q, (uint32_t)leafR);
            staged += nL + __popc(bR);
            __syncwarp();
            if (staged >= BR_FLUSH) {
                if (FILTER) {
                    filter();
                    if (staged >= 32) flush();  // the few survivors wait for company (< 32 carried + <= 64 new < BR_CQ)
                } else {
                    flush();
                }
            }
        }
    };

    // ---- this warp's 128 queries
    const uint32_t wfirst = warp * BR_PER_WARP;          // block-relative slots [wfirst, wend)
    const uint32_t wend = min(wfirst + BR_PER_WARP, nquery > qbase ? nquery - qbase : 0u);
    const uint32_t g = warp;                             // the warp's group (entry list)
    const uint32_t eb = blockIdx.x * BR_GROUPS + g;
    const bool spilled = s_spill[g] != 0;
    const uint32_t nkeep = spilled ? __ldg(entry_count + eb) : s_nkeep[g];  // spilled: scan the unfiltered global list
    uint32_t next = wfirst;                              // warp-uniform
    int q = 0x7fffffff;                                  // this lane's query position ("none": nothing ends after it)
    float qlo[3] = {inf, inf, inf}, qhi[3] = {-inf, -inf, -inf};
    int node = -1;
    uint32_t visits = 0, iters = 0;  // traversal statistics (b200cd_stats::nodes_visited / warp_steps)
    // the first refill's records
    {
        const uint32_t p0 = slot_position(qbase + wfirst + lane);
        if (wfirst + lane < wend && p0 != 0xffffffffu) asm volatile("prefetch.global.L2 [%0];" ::"l"(leaves + p0));
    }

    while (true) {
        const uint32_t idle = __ballot_sync(0xffffffffu, node < 0);
        if (next < wend && (__popc(idle) >= refill || idle == 0xffffffffu)) {
            // ---- refill: idle lanes take the warp's next queries and collect their start subtrees
            const uint32_t slot = next + __popc(idle & lt);
            bool take = node < 0 && slot < wend;
            next = min(wend, next + (uint32_t)__popc(idle));
            if (take) {
                const uint32_t p = slot_position(qbase + slot);
                take = p != 0xffffffffu;
                if (take) {
                    float4 r0, r1, r2, r3;
                    ld256_nc(leaves + p, r0, r1);
                    ld256_nc(reinterpret_cast<const float4*>(leaves + p) + 2, r2, r3);
                    // v0 = r0.xyz, v1 = (r0.w, r1.x, r1.y), v2 = (r1.z, r1.w, r2.x); box.cuh:13-22
                    qlo[0] = min3_ref(r0.x, r0.w, r1.z); qhi[0] = max3_ref(r0.x, r0.w, r1.z);
                    qlo[1] = min3_ref(r0.y, r1.x, r1.w); qhi[1] = max3_ref(r0.y, r1.x, r1.w);
                    qlo[2] = min3_ref(r0.z, r1.y, r2.x); qhi[2] = max3_ref(r0.z, r1.y, r2.x);
                    q = (int)p;
                }
            }
            {   // records of the NEXT refill (it takes at most 32 slots starting at `next`)
                const uint32_t pn = slot_position(qbase + next + lane);
                if (next + lane < wend && pn != 0xffffffffu) asm volatile("prefetch.global.L2 [%0];" ::"l"(leaves + pn));
            }
            for (uint32_t e = 0; e < nkeep; ++e) {
                Child c;
                if (!spilled) c = s_entry[g][e];
                else ld256_nc(entries + (size_t)eb * BR_ENTRIES + e, c.a, c.b);
                const bool hit = take && c.ext() > q && overlap(qlo, qhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y);
                const int link = c.link();  // the same entry for every lane: the branch below is warp-uniform
                if (link >= 0) {
                    if (hit) {
                        if (sp < B200CD_MAX_STACK) stack[sp++] = link; else overflow = 1;
                    }
                } else {
                    stage2((uint32_t)q, hit, ~link, false, 0);  // a start subtree that is a single leaf (rare)
                }
            }
            if (take) node = sp > 0 ? stack[--sp] : -1;
            continue;
        }
        if (idle == 0xffffffffu) break;  // nothing left to take, nobody working
        ++iters;
        bool candL = false, candR = false;
        int leafL = 0, leafR = 0;
        if (node >= 0) {
            ++visits;
            Child l, r;
            load_children(pairs, node, l, r);
            const int linkL = l.link(), linkR = r.link();
            // left child = leaves [.., node], right child = leaves [node+1, r.ext]: skip what ends at or before q
            const bool hitL = node > q && overlap(qlo, qhi, l.a.x, l.a.y, l.a.z, l.a.w, l.b.x, l.b.y);
            const bool hitR = r.ext() > q && overlap(qlo, qhi, r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y);
            candL = hitL && linkL < 0; leafL = ~linkL;
            candR = hitR && linkR < 0; leafR = ~linkR;
            const bool goL = hitL && linkL >= 0, goR = hitR && linkR >= 0;
            // straight-line (predicated) successor selection: descend left first, park the right child, pop when neither
            if (goL && goR) {
                stack[sp] = linkR;  // sp <= B200CD_MAX_STACK - 1: on overflow the top entry is overwritten and the query reports E_DEPTH
                if (sp < B200CD_MAX_STACK - 1) ++sp; else overflow = 1;
            }
            int nxt = goL ? linkL : linkR;
            if (!(goL || goR)) {
                const bool have = sp > 0;
                sp -= have ? 1 : 0;
                nxt = have ? stack[sp] : -1;
            }
            node = nxt;
        }
        stage2((uint32_t)q, candL, leafL, candR, leafR);
    }
    if (FILTER && staged) filter();
    if (staged) flush();
    if (overflow) atomicOr(counters + 2, ERR_STACK);
    visits = __reduce_add_sync(0xffffffffu, visits);
    if (lane == 0) {
        atomicAdd(counters + 3, (unsigned long long)visits);
        atomicAdd(counters + 4, (unsigned long long)iters);
        atomicAdd(counters + 5, (unsigned long long)nkeep);
    }
}

// ---------------------------------------------------------------- K5b (variant 2q): persistent lanes on quantised nodes
#ifndef BR_STEPS
#define BR_STEPS 2  // walk steps per refill check
#endif
// QUANT: the walk reads the 32-byte quantised nodes (common.cuh QNodePair) - one sector and one LDG.256 per visit instead
// of two, three packed query words per lane instead of six floats, seven integer instructions per child box instead of
// six float compares - and emits a SUPERSET of the exact candidates (15-bit cells are conservative); the narrow phase
// drops the extras with the exact box test before anything else. The start subtrees are still tested exactly.
template <int MIN_BLOCKS, bool FILTER>  // MIN_BLOCKS: resident CTAs per SM the register allocation is bounded for
__global__ void __launch_bounds__(BR_THREADS, MIN_BLOCKS)
broad_kernel_q(const NodePair* __restrict__ pairs, const LeafRec* __restrict__ leaves, uint32_t n, uint32_t shard,
             uint32_t nshards, uint32_t chunk, uint32_t nquery, uint32_t ngroups, int refill, const Node32* __restrict__ entries,
             const uint32_t* __restrict__ entry_count, uint2* __restrict__ cand, uint64_t cand_cap,
             unsigned long long* __restrict__ counters, const QNodePair* __restrict__ qpairs, const float* __restrict__ qframe) {
    constexpr bool QUANT = true;                        // (the exact walk of this kernel is kept for A/B builds)
    __shared__ uint2 s_cq[BR_WARPS][BR_CQ];
    __shared__ Child s_entry[BR_WARPS][BR_KEEP];        // the warp's current group: start subtrees that overlap its union box
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint2* wq = s_cq[warp];
    Child* we = s_entry[warp];
    uint32_t staged = 0;  // warp-uniform: candidates waiting in wq (< 128)
    const float inf = __int_as_float(0x7f800000);

    // query slot t -> sorted leaf position
    auto slot_position = [&](uint32_t t) -> uint32_t {
        if (t >= nquery) return 0xffffffffu;
        const uint64_t qq = query_position(t, shard, nshards, chunk);
        return qq + 1 < n ? (uint32_t)qq : 0xffffffffu;  // the last leaf has no partner with a larger position
    };

    // Per-thread stack with a sentinel: entries live in stack[1 .. MAX], stack[0] = -1 is what an empty stack pops ("no
    // node"), sp = index of the top entry, stack[MAX + 1] is written by an overflowing push only (links are >= 0, so the
    // -1 put there now tells at the end whether one happened) - push and pop are straight-line code, no "is it empty /
    // is it full" branch and no flag register in the walk.
    int stack[B200CD_MAX_STACK + 2];
    int sp = 0;
    stack[0] = -1;
    stack[B200CD_MAX_STACK + 1] = -1;
    auto push = [&](int link) {
        sp = min(sp + 1, B200CD_MAX_STACK + 1);
        stack[sp] = link;
    };
    auto pop = [&]() -> int {
        const int v = stack[sp];
        sp = max(sp - 1, 0);
        return v;
    };

    // drop the staged candidates whose triangles share a vertex index (compaction in place, warp-wide)
    auto filter = [&]() {
        uint32_t kept = 0;
        for (uint32_t i0 = 0; i0 < staged; i0 += 32) {
            const uint32_t i = i0 + lane;
            bool keep = false;
            uint2 c = make_uint2(0, 0);
            if (i < staged) {
                c = wq[i];
                // second half of a leaf record: v2.z, vi[0], vi[1], vi[2] (16 bytes at offset 32)
                const float4 a = __ldg(reinterpret_cast<const float4*>(leaves + c.x) + 2);
                const float4 b = __ldg(reinterpret_cast<const float4*>(leaves + c.y) + 2);
                const uint32_t a0 = __float_as_uint(a.y), a1 = __float_as_uint(a.z), a2 = __float_as_uint(a.w);
                const uint32_t b0 = __float_as_uint(b.y), b1 = __float_as_uint(b.z), b2 = __float_as_uint(b.w);
                keep = !(a0 == b0 || a0 == b1 || a0 == b2 || a1 == b0 || a1 == b1 || a1 == b2 || a2 == b0 || a2 == b1 || a2 == b2);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);  // (every lane has read its slot: slots below are free to overwrite)
            if (keep) wq[kept + __popc(m & lt)] = c;
            kept += __popc(m);
            __syncwarp();
        }
        staged = kept;
    };
    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counters + 0, (unsigned long long)staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < staged; i += 32)
            if (base + i < cand_cap) __stcs(cand + base + i, wq[i]);
        staged = 0;
        __syncwarp();
    };
    // stage (q, leaf) candidates of the whole warp: positions by ballot, no atomics
    auto stage2 = [&](uint32_t q, bool candL, int leafL, bool candR, int leafR) {
        const uint32_t bL = __ballot_sync(0xffffffffu, candL), bR = __ballot_sync(0xffffffffu, candR);
        if (bL | bR) {
            const uint32_t nL = __popc(bL);
            if (candL) wq[staged + __popc(bL & lt)] = make_uint2(q, (uint32_t)leafL);
            if (candR) wq[staged + nL + __popc(bR & lt)] = make_uint2(q, (uint32_t)leafR);
            staged += nL + __popc(bR);
            __syncwarp();
            if (staged >= BR_FLUSH) {
                if (FILTER) {
                    filter();
                    if (staged >= 32) flush();  // the few survivors wait for company (< 32 carried + <= 64 new < BR_CQ)
                } else {
                    flush();
                }
            }
        }
    };

    // ---- the warp's groups: first, first + stride, ... A warp never drains between two of them: lanes that are still
    // walking queries of one group keep going (they only need their own stack) while the idle lanes start on the next.
    // cur = (current group << 8) | queries of it already handed out; "before the first group" = first - stride, all 128 out.
    int cur = (int)(((int)(blockIdx.x * BR_WARPS + warp) - (int)(gridDim.x * BR_WARPS)) * 256 + BR_PER_WARP);
    // nk = start subtrees of the current group to scan (bits 0-6), bit 7: more than BR_KEEP of them overlap the union box -
    // scan the global list; bits 8-31: walk steps of this warp since the group was opened (b200cd_stats::warp_steps).
    // visits = lane-steps with a node (b200cd_stats::nodes_visited): both warp-uniform, both flushed per group.
    uint32_t nk = 0, visits = 0;
    // warp-wide: keep the entries of group g that overlap the group's union box, in list order
    auto open_group = [&](uint32_t g) {
        __syncwarp();                                    // the previous group's entries have been read by every lane
        const uint32_t cnt = __ldg(entry_count + g);
        float4 u0, u1;
        ld256_nc(reinterpret_cast<const float4*>(entry_count + ((ngroups + 7u) & ~7u)) + 2 * (size_t)g, u0, u1);
        const float ulo[3] = {u0.x, u0.y, u0.z}, uhi[3] = {u0.w, u1.x, u1.y};
        uint32_t kept = 0;
#pragma unroll 1
        for (uint32_t e0 = 0; e0 < cnt; e0 += 32) {
            const uint32_t e = e0 + lane;
            bool keep = false;
            Child c;
            if (e < cnt) {
                ld256_nc(entries + (size_t)g * BR_ENTRIES + e, c.a, c.b);
                keep = overlap(ulo, uhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            const uint32_t at = kept + __popc(m & lt);
            if (keep && at < (uint32_t)BR_KEEP) we[at] = c;
            kept += __popc(m);
        }
        const uint32_t nkeep = kept > (uint32_t)BR_KEEP ? (cnt | 0x80u) : kept;  // cnt <= BR_ENTRIES = 64
        if (lane == 0) {
            atomicAdd(counters + 3, (unsigned long long)visits);
            atomicAdd(counters + 4, (unsigned long long)(nk >> 8));
            atomicAdd(counters + 5, (unsigned long long)(nkeep & 0x7fu));  // b200cd_stats::start_entries
        }
        visits = 0;
        nk = nkeep;
        __syncwarp();
        // the first refill's records
        const uint32_t t0 = g * BR_PER_WARP + lane;
        const uint32_t p0 = slot_position(t0);
        if (p0 != 0xffffffffu) asm volatile("prefetch.global.L2 [%0];" ::"l"(leaves + p0));
    };

    int q = 0x7fffffff;                                  // this lane's query position ("none": nothing ends after it)
    float qlo[3] = {inf, inf, inf}, qhi[3] = {-inf, -inf, -inf};  // the query box (!QUANT)
    uint32_t qx = 0, qy = 0, qz = 0;                     // QUANT: the query box as packed grid words (common.cuh)
    int node = -1;

    while (true) {
        const uint32_t idle = __ballot_sync(0xffffffffu, node < 0);
        if (__popc(idle) >= refill || idle == 0xffffffffu) {
            int grp = cur >> 8;
            uint32_t rel = (uint32_t)cur & 255u;
            uint32_t wrel = grp >= 0 ? min((uint32_t)BR_PER_WARP, nquery - (uint32_t)grp * BR_PER_WARP) : 0u;  // queries in the group
            if (rel >= wrel) {                            // this group is handed out: on to the warp's next one
                const uint32_t g2 = (uint32_t)(grp + (int)(gridDim.x * BR_WARPS));
                if (g2 < ngroups) {
                    open_group(g2);
                    grp = (int)g2;
                    rel = 0;
                    wrel = min((uint32_t)BR_PER_WARP, nquery - g2 * BR_PER_WARP);
                }
            }
            if (rel < wrel) {
            // ---- refill: idle lanes take the group's next queries and collect their start subtrees
            const uint32_t mine = rel + __popc(idle & lt);
            bool take = node < 0 && mine < wrel;
            rel = min(wrel, rel + (uint32_t)__popc(idle));
            cur = grp * 256 + (int)rel;
            float blo[3] = {0.f, 0.f, 0.f}, bhi[3] = {0.f, 0.f, 0.f};  // the exact box of the query taken in THIS refill
            if (take) {
                const uint32_t p = slot_position((uint32_t)grp * BR_PER_WARP + mine);
                take = p != 0xffffffffu;
                if (take) {
                    float4 r0, r1, r2, r3;
                    ld256_nc(leaves + p, r0, r1);
                    ld256_nc(reinterpret_cast<const float4*>(leaves + p) + 2, r2, r3);
                    // v0 = r0.xyz, v1 = (r0.w, r1.x, r1.y), v2 = (r1.z, r1.w, r2.x); box.cuh:13-22
                    blo[0] = min3_ref(r0.x, r0.w, r1.z); bhi[0] = max3_ref(r0.x, r0.w, r1.z);
                    blo[1] = min3_ref(r0.y, r1.x, r1.w); bhi[1] = max3_ref(r0.y, r1.x, r1.w);
                    blo[2] = min3_ref(r0.z, r1.y, r2.x); bhi[2] = max3_ref(r0.z, r1.y, r2.x);
                    q = (int)p;
                    if (QUANT) {
                        qx = qquery_word(blo[0], bhi[0], __ldg(qframe + 0), __ldg(qframe + 3));
                        qy = qquery_word(blo[1], bhi[1], __ldg(qframe + 1), __ldg(qframe + 4));
                        qz = qquery_word(blo[2], bhi[2], __ldg(qframe + 2), __ldg(qframe + 5));
                    } else {
#pragma unroll
                        for (int k = 0; k < 3; ++k) { qlo[k] = blo[k]; qhi[k] = bhi[k]; }
                    }
                }
            }
            if (rel + lane < wrel) {   // records of the NEXT refill (it takes at most 32 slots starting at rel)
                const uint32_t pn = slot_position((uint32_t)grp * BR_PER_WARP + rel + lane);
                if (pn != 0xffffffffu) asm volatile("prefetch.global.L2 [%0];" ::"l"(leaves + pn));
            }
            const uint32_t nscan = nk & 0x7fu;
#pragma unroll 1
            for (uint32_t e = 0; e < nscan; ++e) {
                Child c;
                if (!(nk & 0x80u)) c = we[e];
                else ld256_nc(entries + (size_t)grp * BR_ENTRIES + e, c.a, c.b);
                const bool hit = take && c.ext() > q && overlap(blo, bhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y);
                const int link = c.link();  // the same entry for every lane: the branch below is warp-uniform
                if (link >= 0) {
                    if (hit) push(link);
                } else {
                    stage2((uint32_t)q, hit, ~link, false, 0);  // a start subtree that is a single leaf (rare)
                }
            }
            if (take) node = pop();
            continue;
            }
            if (idle == 0xffffffffu) break;  // nothing left to take, nobody working
        }
#pragma unroll
        for (int rep = 0; rep < BR_STEPS; ++rep) {  // walk steps per refill check
            nk += 256u;
            visits += BR_STEPS == 1 ? 32u - (uint32_t)__popc(idle) : (uint32_t)__popc(__ballot_sync(0xffffffffu, node >= 0));
            bool candL = false, candR = false;
            int leafL = 0, leafR = 0;
            if (node >= 0) {
                int linkL, linkR;
                bool hitL, hitR;
                // left child = leaves [.., node], right child = leaves [node+1, last leaf of this node]: skip what ends at
                // or before q. Every visited subtree ends after q (start subtrees: checked in the scan; a right child ends
                // where its parent does; a left child is entered only if node > q), so the right child needs no test.
                if (QUANT) {
                    float4 a, b;
                    ld256_nc(qpairs + node, a, b);
                    linkL = __float_as_int(a.w);
                    linkR = __float_as_int(b.w);
                    hitL = node > q && qoverlap(qx, qy, qz, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z));
                    hitR = qoverlap(qx, qy, qz, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z));
                } else {
                    Child l, r;
                    load_children(pairs, node, l, r);
                    linkL = l.link();
                    linkR = r.link();
                    hitL = node > q && overlap(qlo, qhi, l.a.x, l.a.y, l.a.z, l.a.w, l.b.x, l.b.y);
                    hitR = r.ext() > q && overlap(qlo, qhi, r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y);
                }
                candL = hitL && linkL < 0; leafL = ~linkL;
                candR = hitR && linkR < 0; leafR = ~linkR;
                const bool goL = hitL && linkL >= 0, goR = hitR && linkR >= 0;
                // straight-line (predicated) successor selection: descend left first, park the right child, pop when neither
                int nxt = goL ? linkL : linkR;
                if (goL && goR) push(linkR);
                if (!(goL || goR)) nxt = pop();
                node = nxt;
            }
            stage2((uint32_t)q, candL, leafL, candR, leafR);
        }
    }
    if (FILTER && staged) filter();
    if (staged) flush();
    if (stack[B200CD_MAX_STACK + 1] != -1) atomicOr(counters + 2, ERR_STACK);
    if (lane == 0) {
        atomicAdd(counters + 3, (unsigned long long)visits);
        atomicAdd(counters + 4, (unsigned long long)(nk >> 8));
    }
}

// ---------------------------------------------------------------- K5b (variant 1): one query per thread
// STACKLESS (variant 3, B200CD_TRAVERSAL=3; north_star: "a stackless or shared-memory-stack BVH traversal"): no stack at
// all. The node numbering makes the escape pointer implicit: a subtree that ends at leaf L is followed, in depth-first
// order, by the RIGHT child of split node L (leaves L+1 ...), so "pop" becomes "visit node L again, right child only" -
// one more fetch of a 64-byte line the walk has already been through (it is an ancestor), no per-thread memory, two
// registers of state. Measured against the stack versions in DESIGN.md section 4.
template <bool STACKLESS>
__global__ void __launch_bounds__(BR_THREADS)
broad_kernel_simple(const NodePair* __restrict__ pairs, const LeafRec* __restrict__ leaves, const float* __restrict__ root_box,
             uint32_t n, uint32_t shard, uint32_t nshards, uint32_t chunk, uint32_t nquery, int foreign, uint32_t ghost_base,
             const Node32* __restrict__ entries, const uint32_t* __restrict__ entry_count, uint2* __restrict__ cand,
             uint64_t cand_cap, unsigned long long* __restrict__ counters,
             const unsigned long long* __restrict__ nquery_dev /* optional: the query count lives on the device (ghosts that
                                                                   peers appended during this step); nquery is then the cap */) {
    __shared__ uint2 queue[BR_WARPS][BR_CQ];
    __shared__ Child s_entry[BR_ENTRIES];
    __shared__ float s_red[BR_WARPS][6];
    __shared__ uint32_t s_nentry;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint2* wq = queue[warp];
    uint32_t staged = 0;  // warp-uniform
    if (nquery_dev) nquery = (uint32_t)min((unsigned long long)nquery, *nquery_dev);
    uint32_t visits = 0, iters = 0, nkeep_sum = 0;  // traversal statistics (b200cd_stats::nodes_visited / warp_steps)
    bool overflow = false;
  // one block of 256 queries per iteration; the grid covers every block unless the count is device-side
  for (uint32_t blk = blockIdx.x; (uint64_t)blk * BR_THREADS < nquery; blk += gridDim.x) {
    if (threadIdx.x == 0) s_nentry = 0;

    // this thread's query (sorted leaf position)
    const uint32_t t = blk * BR_THREADS + threadIdx.x;
    // q = where the query's record lives (and what the candidate list reports); qcmp = the position
    // the "only leaves after me" rule compares against (-1 for a ghost query: every local leaf counts)
    uint32_t q = 0xffffffffu;
    int qcmp = 0x7fffffff;
    if (t < nquery) {
        if (foreign) {
            q = ghost_base + t;
            qcmp = -1;
        } else {
            const uint64_t qq = query_position(t, shard, nshards, chunk);
            if (qq + 1 < n) { q = (uint32_t)qq; qcmp = (int)qq; }  // the last leaf has no partner with a larger position
        }
    }
    const float inf = __int_as_float(0x7f800000);
    float qlo[3] = {inf, inf, inf}, qhi[3] = {-inf, -inf, -inf};
    if (q != 0xffffffffu) {
        float4 r0, r1, r2, r3;
        ld256_nc(leaves + q, r0, r1);
        ld256_nc(reinterpret_cast<const float4*>(leaves + q) + 2, r2, r3);
        // v0 = r0.xyz, v1 = (r0.w, r1.x, r1.y), v2 = (r1.z, r1.w, r2.x); box.cuh:13-22
        qlo[0] = min3_ref(r0.x, r0.w, r1.z); qhi[0] = max3_ref(r0.x, r0.w, r1.z);
        qlo[1] = min3_ref(r0.y, r1.x, r1.w); qhi[1] = max3_ref(r0.y, r1.x, r1.w);
        qlo[2] = min3_ref(r0.z, r1.y, r2.x); qhi[2] = max3_ref(r0.z, r1.y, r2.x);
    }
    // union box of the block's queries
    float ulo[3], uhi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ulo[k] = qlo[k]; uhi[k] = qhi[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ulo[k] = fminf(ulo[k], __shfl_xor_sync(0xffffffffu, ulo[k], o));
            uhi[k] = fmaxf(uhi[k], __shfl_xor_sync(0xffffffffu, uhi[k], o));
        }
        if (lane == 0) { s_red[warp][k] = ulo[k]; s_red[warp][3 + k] = uhi[k]; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ulo[k] = s_red[0][k]; uhi[k] = s_red[0][3 + k];
#pragma unroll
        for (int w = 1; w < BR_WARPS; ++w) { ulo[k] = fminf(ulo[k], s_red[w][k]); uhi[k] = fmaxf(uhi[k], s_red[w][3 + k]); }
    }
    // keep the entries whose box overlaps the union box (order is irrelevant)
    const uint32_t nent = foreign ? (n > 1 ? 2u : 1u) : __ldg(entry_count + blk);
    if (threadIdx.x < nent) {
        Child c;
        if (!foreign) {
            ld256_nc(entries + (size_t)blk * BR_ENTRIES + threadIdx.x, c.a, c.b);
        } else if (n > 1) {  // ghost queries start at the root: its two children, ext = last leaf
            const int root = reinterpret_cast<const int*>(root_box)[6];
            ld256_nc(&pairs[root].c[threadIdx.x], c.a, c.b);
            if (threadIdx.x == 0) c.b.w = __int_as_float(root);
        } else {             // a one-leaf tree has no internal node: the leaf itself is the only entry
            c.a = make_float4(root_box[0], root_box[1], root_box[2], root_box[3]);
            c.b = make_float4(root_box[4], root_box[5], __int_as_float(~0), __int_as_float(0));
        }
        if (overlap(ulo, uhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y)) s_entry[atomicAdd(&s_nentry, 1u)] = c;
    }
    __syncthreads();
    const uint32_t nkeep = s_nentry;

    int stack[B200CD_MAX_STACK];
    int sp = 0;

    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counters + 0, (unsigned long long)staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < staged; i += 32)
            if (base + i < cand_cap) __stcs(cand + base + i, wq[i]);
        staged = 0;
        __syncwarp();
    };
    // stage (q, leaf) candidates of the whole warp: positions by ballot, no atomics
    auto stage2 = [&](bool candL, int leafL, bool candR, int leafR) {
        const uint32_t bL = __ballot_sync(0xffffffffu, candL), bR = __ballot_sync(0xffffffffu, candR);
        if (bL | bR) {
            const uint32_t nL = __popc(bL);
            if (candL) wq[staged + __popc(bL & lt)] = make_uint2(q, (uint32_t)leafL);
            if (candR) wq[staged + nL + __popc(bR & lt)] = make_uint2(q, (uint32_t)leafR);
            staged += nL + __popc(bR);
            __syncwarp();
            if (staged >= BR_FLUSH) flush();
        }
    };

  if (STACKLESS) {
    int node = -1;            // split node to visit next (-1: take the next start subtree)
    bool ronly = false;       // visit only its right child: the walk comes back from the left one ("pop")
    int limit = -1;           // last leaf of the start subtree being walked: escapes beyond it end the subtree
    uint32_t ecur = 0;        // this lane's cursor over the kept start subtrees
    nkeep_sum += nkeep;
    const bool valid = q != 0xffffffffu;
    while (__any_sync(0xffffffffu, node >= 0 || (valid && ecur < nkeep))) {
        bool candL = false, candR = false;
        int leafL = 0, leafR = 0;
        ++iters;
        if (node >= 0) {
            ++visits;
            Child l, r;
            load_children(pairs, node, l, r);
            const int linkL = l.link(), linkR = r.link();
            const bool hitL = !ronly && node > qcmp && overlap(qlo, qhi, l.a.x, l.a.y, l.a.z, l.a.w, l.b.x, l.b.y);
            candL = hitL && linkL < 0; leafL = ~linkL;
            if (hitL && linkL >= 0) {          // descend left; the right child is met again on the way back (node, ronly)
                node = linkL;
                ronly = false;
            } else {
                const bool hitR = r.ext() > qcmp && overlap(qlo, qhi, r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y);
                candR = hitR && linkR < 0; leafR = ~linkR;
                if (hitR && linkR >= 0) {
                    node = linkR;
                    ronly = false;
                } else {                       // this subtree is done: it ends at leaf r.ext, go on behind it
                    const int L = r.ext();
                    ronly = true;
                    node = L < limit ? L : -1;
                }
            }
        } else if (valid) {                    // next start subtree that ends after q and overlaps q's box
            while (ecur < nkeep) {
                const Child c = s_entry[ecur++];
                if (!(c.ext() > qcmp && overlap(qlo, qhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y))) continue;
                if (c.link() >= 0) { node = c.link(); ronly = false; limit = c.ext(); }
                else { candL = true; leafL = ~c.link(); }  // a start subtree that is a single leaf
                break;
            }
        }
        stage2(candL, leafL, candR, leafR);
    }
  } else {
    // start points: every kept entry that ends after q and overlaps q's box (warp-uniform loop)
    for (uint32_t e = 0; e < nkeep; ++e) {
        const Child c = s_entry[e];
        const bool hit = q != 0xffffffffu && c.ext() > qcmp &&
                         overlap(qlo, qhi, c.a.x, c.a.y, c.a.z, c.a.w, c.b.x, c.b.y);
        const int link = c.link();
        if (hit && link >= 0) {
            if (sp < B200CD_MAX_STACK) stack[sp++] = link; else overflow = true;
        }
        stage2(hit && link < 0, ~link, false, 0);
    }
    int node = sp > 0 ? stack[--sp] : -1;
    nkeep_sum += nkeep;

    while (__any_sync(0xffffffffu, node >= 0)) {
        bool candL = false, candR = false;
        int leafL = 0, leafR = 0;
        ++iters;
        if (node >= 0) {
            ++visits;
            Child l, r;
            load_children(pairs, node, l, r);
            const int linkL = l.link(), linkR = r.link();
            // left child = leaves [.., node], right child = leaves [node+1, r.ext]: skip what ends at or before q
            const bool hitL = node > qcmp && overlap(qlo, qhi, l.a.x, l.a.y, l.a.z, l.a.w, l.b.x, l.b.y);
            const bool hitR = r.ext() > qcmp && overlap(qlo, qhi, r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y);
            candL = hitL && linkL < 0; leafL = ~linkL;
            candR = hitR && linkR < 0; leafR = ~linkR;
            const bool goL = hitL && linkL >= 0, goR = hitR && linkR >= 0;
            if (goL) {
                node = linkL;
                if (goR) {
                    if (sp < B200CD_MAX_STACK) stack[sp++] = linkR; else overflow = true;
                }
            } else if (goR) {
                node = linkR;
            } else {
                node = sp > 0 ? stack[--sp] : -1;
            }
        }
        stage2(candL, leafL, candR, leafR);
    }
  }
    if (staged) flush();
    __syncthreads();  // s_entry / s_nentry are rewritten by the next iteration
  }
    if (overflow) atomicOr(counters + 2, ERR_STACK);
    visits = __reduce_add_sync(0xffffffffu, visits);
    if (lane == 0 && iters) {
        atomicAdd(counters + 3, (unsigned long long)visits);
        atomicAdd(counters + 4, (unsigned long long)iters);
        atomicAdd(counters + 5, (unsigned long long)nkeep_sum);
    }
}

// ---------------------------------------------------------------- K6
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 operator-(const D3& a, const D3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // vec3f.cuh:97-100
__device__ __forceinline__ D3 neg(const D3& a) { return {-a.x, -a.y, -a.z}; }                                     // vec3f.cuh:88-90
__device__ __forceinline__ D3 cross(const D3& a, const D3& b) {                                                   // vec3f.cuh:118-121
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }     // vec3f.cuh:123-125
__device__ __forceinline__ double dmax3(double a, double b, double c) { double t = a; if (b > t) t = b; if (c > t) t = c; return t; }
__device__ __forceinline__ double dmin3(double a, double b, double c) { double t = a; if (b < t) t = b; if (c < t) t = c; return t; }

// vec3f.cuh:257-270
__device__ __forceinline__ bool project3(const D3& ax, const D3& p1, const D3& p2, const D3& p3) {
    double P1 = dot(ax, p1), P2 = dot(ax, p2), P3 = dot(ax, p3);
    double mx1 = dmax3(P1, P2, P3), mn1 = dmin3(P1, P2, P3);
    if (mn1 > 0) return false;
    if (0 > mx1) return false;
    return true;
}
// vec3f.cuh:272-291
__device__ __forceinline__ bool project6(const D3& ax, const D3& p1, const D3& p2, const D3& p3, const D3& q1,
                                         const D3& q2, const D3& q3) {
    double P1 = dot(ax, p1), P2 = dot(ax, p2), P3 = dot(ax, p3);
    double Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
    double mx1 = dmax3(P1, P2, P3), mn1 = dmin3(P1, P2, P3);
    double mx2 = dmax3(Q1, Q2, Q3), mn2 = dmin3(Q1, Q2, Q3);
    if (mn1 > mx2) return false;
    if (mn2 > mx1) return false;
    return true;
}

// tri_contact.cuh:19-78, split in two. The 17 axis tests are a pure conjunction, so their order
// and any early exit cannot change the result; only each expression's own operation order
// matters, and that is kept literal. Stage A = the two face-normal axes (tri_contact.cuh:58-59),
// stage B = the nine edge-edge axes and the six in-plane edge normals (tri_contact.cuh:61-75).
struct SatInput {
    D3 p2, p3, q1, q2, q3;  // everything translated by P1; p1 is exactly (0,0,0) (tri_contact.cuh:21-26)
};
__device__ __forceinline__ SatInput sat_input(const D3& P1, const D3& P2, const D3& P3, const D3& Q1, const D3& Q2,
                                              const D3& Q3) {
    return {P2 - P1, P3 - P1, Q1 - P1, Q2 - P1, Q3 - P1};
}
__device__ __forceinline__ bool sat_stage_a(const SatInput& t) {
    const D3 p1 = {0.0, 0.0, 0.0};
    const D3 e1 = t.p2 - p1, e2 = t.p3 - t.p2;
    const D3 f1 = t.q2 - t.q1, f2 = t.q3 - t.q2;
    const D3 n1 = cross(e1, e2);
    if (!project3(n1, t.q1, t.q2, t.q3)) return false;
    const D3 m1 = cross(f1, f2);
    return project3(m1, neg(t.q1), t.p2 - t.q1, t.p3 - t.q1);
}
__device__ __forceinline__ bool sat_stage_b(const SatInput& t) {
    const D3 p1 = {0.0, 0.0, 0.0};
    const D3 &p2 = t.p2, &p3 = t.p3, &q1 = t.q1, &q2 = t.q2, &q3 = t.q3;
    const D3 e1 = p2 - p1, e2 = p3 - p2, e3 = p1 - p3;
    const D3 f1 = q2 - q1, f2 = q3 - q2, f3 = q1 - q3;
    if (!project6(cross(e1, f1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e1, f2), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e1, f3), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e2, f1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e2, f2), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e2, f3), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e3, f1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e3, f2), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e3, f3), p1, p2, p3, q1, q2, q3)) return false;
    const D3 n1 = cross(e1, e2), m1 = cross(f1, f2);
    if (!project6(cross(e1, n1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e2, n1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(e3, n1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(f1, m1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(f2, m1), p1, p2, p3, q1, q2, q3)) return false;
    if (!project6(cross(f3, m1), p1, p2, p3, q1, q2, q3)) return false;
    return true;
}

// vertices only (stage B re-reads the two records; they are L1/L2-hot)
__device__ __forceinline__ void load_verts(const LeafRec* __restrict__ leaves, uint32_t pos, D3& v0, D3& v1, D3& v2,
                                           uint32_t& id) {
    float4 r0, r1, r2, r3;
    ld256_nc(leaves + pos, r0, r1);
    ld256_nc(reinterpret_cast<const float4*>(leaves + pos) + 2, r2, r3);
    v0 = {(double)r0.x, (double)r0.y, (double)r0.z};
    v1 = {(double)r0.w, (double)r1.x, (double)r1.y};
    v2 = {(double)r1.z, (double)r1.w, (double)r2.x};
    id = __float_as_uint(r3.x);
}

constexpr int NR_THREADS = 256;
constexpr int NR_WARPS = NR_THREADS / 32;
constexpr int NR_QUEUE = 64;  // per-warp survivors of stage A waiting for stage B (<= 31 carried + 32 new)

// Stage B on up to 32 queued (P position, Q position) entries, one per lane; contacts are appended
// with one atomic per warp.
__device__ __forceinline__ void narrow_stage_b(const LeafRec* __restrict__ leaves, const uint2* wq, uint32_t count,
                                               uint32_t lane, uint2* __restrict__ out, uint64_t out_cap,
                                               unsigned long long* __restrict__ counters) {
    bool hit = false;
    uint2 res = make_uint2(0, 0);
    if (lane < count) {
        const uint2 e = wq[lane];  // x = sorted position of the lower-ID triangle (P), y = of the higher-ID one (Q)
        D3 P1, P2, P3, Q1, Q2, Q3;
        load_verts(leaves, e.x, P1, P2, P3, res.x);
        load_verts(leaves, e.y, Q1, Q2, Q3, res.y);
        hit = sat_stage_b(sat_input(P1, P2, P3, Q1, Q2, Q3));
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (m) {
        unsigned long long o = 0;
        if (lane == 0) o = atomicAdd(counters + 1, (unsigned long long)__popc(m));
        o = __shfl_sync(0xffffffffu, o, 0) + __popc(m & ((1u << lane) - 1u));
        if (hit && o < out_cap) out[o] = res;
    }
}

// Two-stage narrow phase. Most candidates (AABB-overlapping leaf pairs) are rejected by the two
// face-normal axes; running all 17 axes per thread leaves a warp with a handful of live lanes.
// So: stage A (filters + 2 axes) runs dense over the candidate list, its survivors are
// compacted into a per-warp shared-memory queue by ballot, and stage B (15 axes) runs whenever
// 32 survivors are waiting - both stages execute with (nearly) full warps.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(NR_THREADS, MIN_BLOCKS)
narrow_kernel(const LeafRec* __restrict__ leaves, const uint2* __restrict__ cand, uint64_t cand_cap,
              uint2* __restrict__ out, uint64_t out_cap, unsigned long long* __restrict__ counters) {
    __shared__ uint2 queue[NR_WARPS][NR_QUEUE];
    const unsigned long long total = min((unsigned long long)cand_cap, counters[0]);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint2* wq = queue[warp];
    uint32_t queued = 0;  // warp-uniform
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    // warp-uniform trip count so the ballots below are full-warp
    const unsigned long long first = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    for (unsigned long long base = first; base < total; base += stride) {
        const unsigned long long i = base + lane;
        bool alive = false;
        uint2 entry = make_uint2(0, 0);
        if (i < total) {
            const uint2 c = __ldcs(cand + i);
            // the vertex indices and the ID sit in the SECOND 32 bytes of a record: on meshes most candidates are
            // vertex-sharing neighbours, rejected here for one sector per triangle instead of two
            float4 a2, a3, b2, b3;
            ld256_nc(reinterpret_cast<const float4*>(leaves + c.x) + 2, a2, a3);
            ld256_nc(reinterpret_cast<const float4*>(leaves + c.y) + 2, b2, b3);
            const uint32_t ai0 = __float_as_uint(a2.y), ai1 = __float_as_uint(a2.z), ai2 = __float_as_uint(a2.w);
            const uint32_t bi0 = __float_as_uint(b2.y), bi1 = __float_as_uint(b2.z), bi2 = __float_as_uint(b2.w);
            const uint32_t aid = __float_as_uint(a3.x), bid = __float_as_uint(b3.x);
            // triangle.cuh:18-30: neighborCount >= 1 <=> any vertex index shared
            const bool shared = ai0 == bi0 || ai0 == bi1 || ai0 == bi2 || ai1 == bi0 || ai1 == bi1 || ai1 == bi2 ||
                                ai2 == bi0 || ai2 == bi1 || ai2 == bi2;
            if (!shared && aid != bid) {
                float4 a0, a1, b0, b1;
                ld256_nc(leaves + c.x, a0, a1);
                ld256_nc(leaves + c.y, b0, b1);
                // box.cuh:13-22 + 40-43 on the exact vertices: the quantised traversal admits a few box pairs that only
                // touch or just miss (collision.cuh:36 tests the boxes before anything else)
                const float alo[3] = {min3_ref(a0.x, a0.w, a1.z), min3_ref(a0.y, a1.x, a1.w), min3_ref(a0.z, a1.y, a2.x)};
                const float ahi[3] = {max3_ref(a0.x, a0.w, a1.z), max3_ref(a0.y, a1.x, a1.w), max3_ref(a0.z, a1.y, a2.x)};
                const bool boxes = overlap(alo, ahi, min3_ref(b0.x, b0.w, b1.z), min3_ref(b0.y, b1.x, b1.w), min3_ref(b0.z, b1.y, b2.x),
                                           max3_ref(b0.x, b0.w, b1.z), max3_ref(b0.y, b1.x, b1.w), max3_ref(b0.z, b1.y, b2.x));
                if (boxes) {
                const D3 av0 = {(double)a0.x, (double)a0.y, (double)a0.z}, av1 = {(double)a0.w, (double)a1.x, (double)a1.y},
                         av2 = {(double)a1.z, (double)a1.w, (double)a2.x};
                const D3 bv0 = {(double)b0.x, (double)b0.y, (double)b0.z}, bv1 = {(double)b0.w, (double)b1.x, (double)b1.y},
                         bv2 = {(double)b1.z, (double)b1.w, (double)b2.x};
                // tri_contact.cuh:81-86: P is the lower-ID triangle
                if (aid < bid) { alive = sat_stage_a(sat_input(av0, av1, av2, bv0, bv1, bv2)); entry = make_uint2(c.x, c.y); }
                else           { alive = sat_stage_a(sat_input(bv0, bv1, bv2, av0, av1, av2)); entry = make_uint2(c.y, c.x); }
                }
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, alive);
        if (alive) wq[queued + __popc(m & lt)] = entry;
        queued += __popc(m);
        __syncwarp();
        if (queued >= 32) {
            narrow_stage_b(leaves, wq + (queued - 32), 32, lane, out, out_cap, counters);  // newest 32: the rest stays at the front
            queued -= 32;
            __syncwarp();
        }
    }
    if (queued) narrow_stage_b(leaves, wq, queued, lane, out, out_cap, counters);
}

}  // namespace

// B200CD_TRAVERSAL=1: one query per thread, per-thread stack; 3: one query per thread, STACKLESS; 2 (default): persistent
// lanes with a per-thread stack (tuning knob, read once)
static int traversal_variant() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("B200CD_TRAVERSAL");
        v = (e && e[0] == '1') ? 1 : (e && e[0] == '3') ? 3 : 2;
    }
    return v;
}

// Which builds write (and which traversals read) the quantised nodes: triangle soups by default - on a mesh of flat
// sheets the fixed 15-bit grid over the Morton box is too coarse ACROSS the sheets and the extra candidates cost more than
// the halved node fetches save (measured). B200CD_BROAD_QUANT=0 never, =1 always.
bool broad_uses_quantised_nodes(bool shared_vertices) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("B200CD_BROAD_QUANT");
        mode = !e ? 2 : (e[0] == '0' ? 0 : 1);
    }
    if (traversal_variant() != 2) return false;
    return mode == 1 || (mode == 2 && !shared_vertices);
}

void launch_broad(const NodePair* d_pairs, const LeafRec* d_leaves, const float* d_root_box, uint32_t n, uint32_t shard,
                  uint32_t nshards, uint32_t chunk, uint32_t nquery, int foreign, uint32_t ghost_base, Node32* d_entries,
                  uint32_t* d_entry_count, uint2* d_cand, uint64_t cand_cap, unsigned long long* d_counters,
                  cudaStream_t s, const unsigned long long* d_nquery, int sms, bool shared_vertices, const QNodePair* d_qpairs,
                  const float* d_qframe) {
    if (nquery == 0 || n == 0 || (!foreign && n < 2)) return;
    if (foreign && d_nquery) {  // ghost queries whose count only the device knows: a fixed grid strides over the blocks
        const uint32_t blocks = std::min<uint32_t>((nquery + BR_THREADS - 1) / BR_THREADS, (uint32_t)std::max(sms, 1) * 8u);
        broad_kernel_simple<false><<<blocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, d_root_box, n, 0, 1, BR_THREADS, nquery, 1, ghost_base,
                                                          d_entries, d_entry_count, d_cand, cand_cap, d_counters, d_nquery);
        count_launch();
        trace_mark("broad_kernel_simple (ghost queries)", s);
        return;
    }
    const bool persistent = traversal_variant() == 2 && !foreign;  // ghost queries (few, no tree over them) use the simple kernel
    const uint32_t gsize = persistent ? BR_PER_WARP : BR_THREADS;  // consecutive queries per entry list
    const uint32_t groups = (nquery + gsize - 1) / gsize;
    if (!foreign) {
        entry_kernel<<<(groups + 127) / 128, 128, 0, s>>>(d_pairs, d_root_box, n, shard, nshards, chunk, gsize, groups, d_entries,
                                                          d_entry_count);
        count_launch();
        trace_mark("entry_kernel", s);
    }
    if (persistent) {
        const uint32_t blocks = (nquery + BR_QB - 1) / BR_QB;
        static int refill = 0;
        if (!refill) {
            const char* e = getenv("B200CD_REFILL");
            refill = e ? atoi(e) : BR_REFILL;
            if (refill < 1 || refill > 32) refill = BR_REFILL;
        }
        static int occ = 0;  // B200CD_BROAD_OCC=5: the 48-register build (5 CTAs per SM); default 6 CTAs per SM
        if (!occ) {
            const char* e = getenv("B200CD_BROAD_OCC");
            occ = (e && e[0] == '5') ? 5 : 6;
        }
        static int filt = -1;  // B200CD_BROAD_FILTER=0: never drop shared-vertex candidates in the traversal (A/B knob)
        if (filt < 0) {
            const char* e = getenv("B200CD_BROAD_FILTER");
            filt = (e && e[0] == '0') ? 0 : 1;
        }
        const bool filter = filt && shared_vertices;
        if (d_qpairs && d_qframe) {  // the build wrote the quantised nodes (broad_uses_quantised_nodes): walk those
            // B200CD_BROAD_GRID=persist: one resident wave of blocks, every warp works through groups w, w + all warps, ...
            // without draining in between (lane utilisation 0.53 -> 0.61, measured SLOWER: DESIGN.md section 4)
            static int persist = -1;
            if (persist < 0) {
                const char* e = getenv("B200CD_BROAD_GRID");
                persist = (e && e[0] == 'p') ? 1 : 0;
            }
            const uint32_t qblocks = persist ? std::min<uint32_t>(blocks, (uint32_t)(sms > 0 ? sms : 148) * (uint32_t)occ) : blocks;
#define B200CD_BROAD_Q(OCC, F)                                                                                                     \
    broad_kernel_q<OCC, F><<<qblocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, n, shard, nshards, chunk, nquery, groups, refill, d_entries, \
                                                          d_entry_count, d_cand, cand_cap, d_counters, d_qpairs, d_qframe)
            if (occ == 5) { if (filter) B200CD_BROAD_Q(5, true); else B200CD_BROAD_Q(5, false); }
            else          { if (filter) B200CD_BROAD_Q(6, true); else B200CD_BROAD_Q(6, false); }
#undef B200CD_BROAD_Q
        } else if (filter && occ == 5)
            broad_kernel<5, true><<<blocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, n, shard, nshards, chunk, nquery, groups, refill,
                                                                d_entries, d_entry_count, d_cand, cand_cap, d_counters);
        else if (filter)
            broad_kernel<6, true><<<blocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, n, shard, nshards, chunk, nquery, groups, refill,
                                                                d_entries, d_entry_count, d_cand, cand_cap, d_counters);
        else if (occ == 5)
            broad_kernel<5, false><<<blocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, n, shard, nshards, chunk, nquery, groups, refill,
                                                                 d_entries, d_entry_count, d_cand, cand_cap, d_counters);
        else
            broad_kernel<6, false><<<blocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, n, shard, nshards, chunk, nquery, groups, refill,
                                                                 d_entries, d_entry_count, d_cand, cand_cap, d_counters);
    } else {
        if (traversal_variant() == 3 && !foreign)
            broad_kernel_simple<true><<<groups, BR_THREADS, 0, s>>>(d_pairs, d_leaves, d_root_box, n, shard, nshards, chunk, nquery,
                                                                    foreign, ghost_base, d_entries, d_entry_count, d_cand, cand_cap,
                                                                    d_counters, nullptr);
        else
            broad_kernel_simple<false><<<groups, BR_THREADS, 0, s>>>(d_pairs, d_leaves, d_root_box, n, shard, nshards, chunk, nquery,
                                                                     foreign, ghost_base, d_entries, d_entry_count, d_cand, cand_cap,
                                                                     d_counters, nullptr);
    }
    count_launch();
    trace_mark("broad_kernel", s);
}

void launch_narrow(const LeafRec* d_leaves, const uint2* d_cand, uint64_t cand_cap, uint2* d_out, uint64_t out_cap,
                   unsigned long long* d_counters, int sms, cudaStream_t s, bool unshared_vertices) {
    // On a mesh most candidates are vertex-sharing neighbours that leave after one 32-byte read: latency-bound, the
    // 64-register build with 4 CTAs per SM wins (0.77 vs 0.88 ms on the 16 M sheets). On a soup every candidate runs
    // the fp64 face-normal axes and the 64-register build spills (0.85 vs 0.83 ms). B200CD_NARROW_OCC=3/4 overrides.
    static int force = -1;
    if (force < 0) {
        const char* e = getenv("B200CD_NARROW_OCC");
        force = e ? (e[0] == '4' ? 4 : 3) : 0;
    }
    const int occ = force ? force : (unshared_vertices ? 3 : 4);
    if (occ == 4)
        narrow_kernel<4><<<sms * 4 * 4, NR_THREADS, 0, s>>>(d_leaves, d_cand, cand_cap, d_out, out_cap, d_counters);
    else
        narrow_kernel<3><<<sms * 3 * 4, NR_THREADS, 0, s>>>(d_leaves, d_cand, cand_cap, d_out, out_cap, d_counters);
    count_launch();
    trace_mark("narrow_kernel", s);
}

}  // namespace b200cd
