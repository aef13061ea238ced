// lbvh.cu — K3+K4 fused: LBVH topology, leaf records and node boxes in ONE bottom-up pass,
// plus the parity/validation helpers (node export in the reference's numbering, self-checks).
//
// Reference semantics (under /root/reference/CollisionDetection/):
//   hierarchy  bvh.cuh:48 (delta), :100-123 (determineRange), :57-98 (findSplit),
//              :146-199 (generateHierarchyParallel): a binary radix tree over the sorted Morton
//              codes; internal node i covers the key range that contains i and its more-similar
//              neighbour; root = internal 0.
//   refit      bvh.cuh:258-285 (calBoundingBox), box.cuh:13-32 (Box::set / Box::merge),
//              mathop.cuh:17-44 (comparison-based min/max): leaves first, the SECOND thread to
//              reach a node merges its children (atomic visit counter).
//
// What is ours:
//   - The reference finds every internal node's range and split with two binary searches over
//     the keys (top-down information), then climbs again for the boxes. The same tree is the
//     Cartesian tree of the adjacent-key similarities delta(i) = clz(key[i] ^ key[i+1]): a node
//     covering [F, L] hangs under the split between L and L+1 if delta(L) > delta(F-1), else under
//     the split between F-1 and F (Apetrei 2014). So one bottom-up climb builds topology AND
//     boxes: no parent array, no search kernel. Equal keys are ordered by sorted position
//     (delta = 64 + clz(i ^ (i+1)), Karras 2012 section 4); the reference builds a malformed tree for
//     duplicate codes (load_obj.h:110-115 only reports them). With unique keys the tree is
//     node-for-node the reference's; b200cd_bvh_download renumbers nodes to the reference's
//     (Karras) indices for the parity tests.
//   - Storage: internal node s = "the split between sorted leaves s and s+1" lives in pairs[s]
//     (64 B): both children side by side, each a 32-byte Node32 {box, link, ext} with
//     ext = first leaf of the subtree for the left child (its last leaf is s itself) and
//     ext = last leaf for the right child (its first leaf is s+1).
//   - Each block owns 256 consecutive sorted leaves and climbs in SHARED memory while both
//     children of a split lie inside the block (~97 % of all merges): the arrival flag, the
//     sibling's half and the similarities never touch HBM, and the finished 64-byte node is
//     written once. Only the few nodes that straddle block boundaries use the global protocol:
//     publish my half, __threadfence(), atomic arrival flag, read the sibling's half through L2
//     (the reference has neither fence nor cache bypass, bvh.cuh:270-278).
#include "common.cuh"

namespace b200cd {

#ifdef BT_PROFILE  // scratch builds only: cycles per phase of build_kernel (thread 0 of every block) / upper_kernel
__device__ unsigned long long g_bt_prof[16];
extern "C" __attribute__((visibility("default"))) void b200cd_debug_bt_prof(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, g_bt_prof, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_bt_prof, z, sizeof z); }
}
#define BT_MARK(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); atomicAdd(&g_bt_prof[i], (unsigned long long)(t__ - t_prev)); t_prev = t__; } } while (0)
#else
#define BT_MARK(i) do { } while (0)
#endif

namespace {

#ifndef BL_V
#define BL_V 256
#endif
constexpr int BL = BL_V;  // sorted leaves per block (tuning builds: -DBL_V=512)

__device__ __forceinline__ float min3_ref(float a, float b, float c) {  // mathop.cuh:38-44
    float t = a;
    if (b < t) t = b;
    if (c < t) t = c;
    return t;
}
__device__ __forceinline__ float max3_ref(float a, float b, float c) {  // mathop.cuh:30-36
    float t = a;
    if (b > t) t = b;
    if (c > t) t = c;
    return t;
}
__device__ __forceinline__ float min2_ref(float a, float b) { return (a < b) ? a : b; }  // mathop.cuh:21-23
__device__ __forceinline__ float max2_ref(float a, float b) { return (a > b) ? a : b; }  // mathop.cuh:17-19

// similarity of the sorted neighbours i and i+1 (bvh.cuh:48 restricted to adjacent keys, + tie-break)
__device__ __forceinline__ int similarity(uint64_t a, uint64_t b, int i) {
    const uint64_t x = a ^ b;
    return x ? __clzll((long long)x) : 64 + __clz(i ^ (i + 1));
}
__device__ __forceinline__ int similarity_at(const uint64_t* __restrict__ keys, int n, int i) {
    if (i < 0 || i >= n - 1) return -1;
    return similarity(__ldg(keys + i), __ldg(keys + i + 1), i);
}

struct Carry {  // the subtree a climbing thread currently holds
    float lo[3], hi[3];
    int link;  // >= 0: internal node (its split index), < 0: ~leaf position
    int F, L;  // leaf range
};

__device__ __forceinline__ void pack(const Carry& c, int ext, float4& a, float4& b) {
    a = make_float4(c.lo[0], c.lo[1], c.lo[2], c.hi[0]);
    b = make_float4(c.hi[1], c.hi[2], __int_as_float(c.link), __int_as_float(ext));
}

// Box::merge(childA, childB), box.cuh:24-32 - operand order A (left) then B (right)
__device__ __forceinline__ void merge_with(Carry& me, int side, const float4& sa, const float4& sb, int s) {
    const float slo[3] = {sa.x, sa.y, sa.z}, shi[3] = {sa.w, sb.x, sb.y};
    const int sext = __float_as_int(sb.w);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float alo = side ? slo[k] : me.lo[k], blo = side ? me.lo[k] : slo[k];
        const float ahi = side ? shi[k] : me.hi[k], bhi = side ? me.hi[k] : shi[k];
        me.lo[k] = min2_ref(alo, blo);
        me.hi[k] = max2_ref(ahi, bhi);
    }
    if (side) me.F = sext; else me.L = sext;  // the sibling's far end
    me.link = s;
}

__device__ __forceinline__ void write_root(const Carry& c, float* __restrict__ root_box) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { root_box[k] = c.lo[k]; root_box[3 + k] = c.hi[k]; }
    reinterpret_cast<int*>(root_box)[6] = c.link;  // index of the root node (or ~0 when n == 1)
}

__device__ __forceinline__ QFrame load_qframe(const float* __restrict__ qframe) {
    QFrame f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        f.o[k] = qframe ? __ldg(qframe + k) : 0.f;
        f.s[k] = qframe ? __ldg(qframe + 3 + k) : 0.f;
    }
    return f;
}

// The 15-bit grid of the quantised nodes (common.cuh QFrame). A rebuild lays it over the root box of the PREVIOUS build
// (the mesh of a simulation moves little from frame to frame; whatever leaves the box lands in the border cells, still
// conservative) - tight on every axis, where the Morton box of b200cd_params can be far larger than the mesh along one
// of them (two flat sheets in a unit cube: 46 % more candidates on the cube-wide grid). First build: the Morton box.
__global__ void qframe_kernel(const float* __restrict__ root_box, int have_root, const QFrame fallback, float* __restrict__ qframe) {
    const int k = threadIdx.x;
    if (k >= 3) return;
    float o = fallback.o[k], s = fallback.s[k];
    if (have_root) {
        const float lo = root_box[k], hi = root_box[3 + k];
        const float ext = hi - lo;
        if (ext > 0.f && ext < 3.0e38f) {
            const float sc = (float)Q_CELLS / ext;
            if (sc > 0.f && sc < 3.0e38f) { o = lo; s = sc; }
        }
    }
    qframe[k] = o;
    qframe[3 + k] = s;
}

// Global-memory climb for the few subtrees that straddle block boundaries (bvh.cuh:269-283 protocol
// + fences). `have_slot`: the first arrival's split/side are already known (conversion of a deposit
// left in shared memory).
__device__ void climb_global(Carry c, bool have_slot, int s, int side, const uint64_t* __restrict__ keys, int n,
                             uint32_t* __restrict__ flags, NodePair* __restrict__ pairs, float* __restrict__ root_box,
                             QNodePair* __restrict__ qpairs, const QFrame& qf) {  // qf: loaded by the caller iff qpairs
    while (true) {
        if (!have_slot) {
            const int dl = similarity_at(keys, n, c.F - 1), dr = similarity_at(keys, n, c.L);
            if (dl < 0 && dr < 0) {
                write_root(c, root_box);
                return;
            }
            const bool right = dr > dl;
            s = right ? c.L : c.F - 1;
            side = right ? 0 : 1;
        }
        have_slot = false;
        float4 a, b;
        pack(c, side ? c.L : c.F, a, b);
        st256_cg(&pairs[s].c[side], a, b);
        if (qpairs) qpairs[s].c[side] = qnode_half(a, b, qf);  // read by later kernels only: no ordering needed
        // arrival counter with release (my half is visible before the count) and acquire (the sibling's half is
        // read after it) semantics in ONE instruction instead of two device-wide fences around a relaxed atomic
        unsigned int old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(flags + s) : "memory");
        if (old == 0u) return;
        float4 sa, sb;
        ld256_cg(&pairs[s].c[side ^ 1], sa, sb);  // published before the sibling's atomic; read through L2
        merge_with(c, side, sa, sb, s);
    }
}

// A subtree whose parent split lies outside its block's shared-memory range: parked in a global
// list by build_kernel and carried up by upper_kernel, so that the 256-leaf blocks retire as soon
// as their shared-memory climb is done instead of idling behind a few threads doing
// fence + atomic round trips.
struct __align__(16) Pending {
    float4 a, b;      // Node32 image: lo.xyz hi.x | hi.yz link ext
    int F, L;         // leaf range
    int slot, side;   // side < 0: the parent split is still to be decided from the keys
};

__global__ void __launch_bounds__(128)
upper_kernel(const Pending* __restrict__ list, const uint32_t* __restrict__ list_count, uint32_t capacity,
             const uint64_t* __restrict__ keys, int n, uint32_t* __restrict__ flags, NodePair* __restrict__ pairs,
             float* __restrict__ root_box, QNodePair* __restrict__ qpairs, const float* __restrict__ qframe) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(__ldg(list_count), capacity)) return;
    const QFrame qf = load_qframe(qpairs ? qframe : nullptr);
    const Pending p = list[i];
    Carry c;
    c.lo[0] = p.a.x; c.lo[1] = p.a.y; c.lo[2] = p.a.z; c.hi[0] = p.a.w; c.hi[1] = p.b.x; c.hi[2] = p.b.y;
    c.link = __float_as_int(p.b.z);
    c.F = p.F; c.L = p.L;
    climb_global(c, p.side >= 0, p.slot, p.side, keys, n, flags, pairs, root_box, qpairs, qf);
}

template <bool RECS>  // leaf records come from the face-ordered copies K1 wrote (recs) instead of idx / verts
__global__ void __launch_bounds__(BL)
build_kernel(const float4* __restrict__ verts, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ sorted_ids,
             const uint64_t* __restrict__ keys, int n, uint32_t* __restrict__ flags, NodePair* __restrict__ pairs,
             LeafRec* __restrict__ leaves, float* __restrict__ root_box, Pending* __restrict__ list,
             uint32_t* __restrict__ list_count, uint32_t capacity, const LeafRec* __restrict__ recs,
             float* __restrict__ block_boxes /* optional: [blocks][8], union box of each block's 256 leaves */,
             QNodePair* __restrict__ qpairs, const float* __restrict__ qframe) {
    __shared__ float s_bb[BL / 32][6];
    __shared__ int s_sim[BL + 1];          // s_sim[i] = similarity of sorted positions (B0-1+i, B0+i); -1 outside
    __shared__ uint32_t s_flag[BL];        // per split B0+i: bit 0 = left child arrived, bit 1 = right child arrived
    __shared__ float4 s_dep[BL][2][2];     // per split, per side: the child's Node32 (ext as in pairs[])
    __shared__ uint32_t s_npend, s_base;
    const int tid = threadIdx.x;
    const int B0 = blockIdx.x * BL;
    const int j = B0 + tid;
    const int Bend = min(B0 + BL, n) - 1;  // last leaf of this block
    const QFrame qf = load_qframe(qpairs ? qframe : nullptr);
#ifdef BT_PROFILE
    long long t_prev = clock64();
#endif

    // ---- leaf record + leaf box; similarities of the block's neighbours
    Carry c;
    uint64_t kj = 0;
    if (j < n) {
        const uint32_t id = __ldg(sorted_ids + j);
        float4 a, b, d;
        float4* rec = reinterpret_cast<float4*>(leaves + j);
        if (RECS) {
            float4 r0, r1, r2, r3;
            ld256_nc(recs + id, r0, r1);
            ld256_nc(reinterpret_cast<const float4*>(recs + id) + 2, r2, r3);
            st256(rec, r0, r1);
            st256(rec + 2, r2, r3);
            a = make_float4(r0.x, r0.y, r0.z, 0.f); b = make_float4(r0.w, r1.x, r1.y, 0.f); d = make_float4(r1.z, r1.w, r2.x, 0.f);
        } else {
            const uint32_t* f = idx + 3ull * id;
            const uint32_t i0 = __ldg(f), i1 = __ldg(f + 1), i2 = __ldg(f + 2);
            a = __ldg(verts + i0); b = __ldg(verts + i1); d = __ldg(verts + i2);
            st256(rec, make_float4(a.x, a.y, a.z, b.x), make_float4(b.y, b.z, d.x, d.y));
            st256(rec + 2, make_float4(d.z, __uint_as_float(i0), __uint_as_float(i1), __uint_as_float(i2)),
                  make_float4(__uint_as_float(id), 0.f, 0.f, 0.f));
        }
        // box.cuh:13-22
        c.lo[0] = min3_ref(a.x, b.x, d.x); c.lo[1] = min3_ref(a.y, b.y, d.y); c.lo[2] = min3_ref(a.z, b.z, d.z);
        c.hi[0] = max3_ref(a.x, b.x, d.x); c.hi[1] = max3_ref(a.y, b.y, d.y); c.hi[2] = max3_ref(a.z, b.z, d.z);
        c.link = ~j;
        c.F = c.L = j;
        kj = __ldg(keys + j);
        // similarity(j, j+1) goes to s_sim[tid + 1]
        s_sim[tid + 1] = (j + 1 < n) ? similarity(kj, __ldg(keys + j + 1), j) : -1;
        if (tid == 0) s_sim[0] = (j > 0) ? similarity(__ldg(keys + j - 1), kj, j - 1) : -1;
    } else {
        s_sim[tid + 1] = -1;
    }
    s_flag[tid] = 0;
    if (tid == 0) s_npend = 0;
    if (block_boxes) {  // (partitioned builds) the ghost selection tests whole blocks against the peers' boxes first
        const float inf = __int_as_float(0x7f800000);
        float v[6] = {j < n ? c.lo[0] : inf, j < n ? c.lo[1] : inf, j < n ? c.lo[2] : inf,
                      j < n ? c.hi[0] : -inf, j < n ? c.hi[1] : -inf, j < n ? c.hi[2] : -inf};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float t = __shfl_xor_sync(0xffffffffu, v[k], o);
                v[k] = k < 3 ? fminf(v[k], t) : fmaxf(v[k], t);
            }
            if ((tid & 31) == 0) s_bb[tid >> 5][k] = v[k];
        }
    }
    BT_MARK(0);  // gathers + leaf record + similarities (thread 0)
    __syncthreads();
    BT_MARK(1);  // wait for the block
    if (block_boxes && tid < 6) {
        float u = s_bb[0][tid];
#pragma unroll
        for (int w = 1; w < BL / 32; ++w) u = tid < 3 ? fminf(u, s_bb[w][tid]) : fmaxf(u, s_bb[w][tid]);
        block_boxes[8 * (size_t)blockIdx.x + tid] = u;
    }

    // ---- climb inside the block: splits s with both neighbours in the block, B0 <= s < Bend
    bool pending = false;  // holding a subtree whose parent split lies outside the block's shared-memory range
    if (j < n) {
        if (n == 1) {
            write_root(c, root_box);
        } else {
            while (true) {
                const int dl = s_sim[c.F - B0], dr = s_sim[c.L - B0 + 1];  // similarity(F-1), similarity(L)
                if (dl < 0 && dr < 0) {  // covers [0, n-1]
                    write_root(c, root_box);
                    break;
                }
                const bool right = dr > dl;          // parent extends to the right: I am its left child
                const int s = right ? c.L : c.F - 1;
                const int side = right ? 0 : 1;
                if (s < B0 || s >= Bend) {
                    pending = true;
                    break;
                }
                const int ls = s - B0;
                float4 a, b;
                pack(c, side ? c.L : c.F, a, b);
                s_dep[ls][side][0] = a;
                s_dep[ls][side][1] = b;
                __threadfence_block();
                const uint32_t old = atomicOr(&s_flag[ls], 1u << side);
                if (old == 0u) break;  // first arrival: the deposit stays for the sibling
                __threadfence_block();
                const float4 sa = s_dep[ls][side ^ 1][0], sb = s_dep[ls][side ^ 1][1];
                // both halves of node s now sit in s_dep[ls]: the finished nodes of the block are written out
                // together after the climb (global stores inside this loop would make every fence wait for them)
                merge_with(c, side, sa, sb, s);
            }
        }
    }
    BT_MARK(2);  // shared-memory climb (thread 0)
    __syncthreads();
    BT_MARK(3);  // wait for the block's longest climb
    // ---- finished nodes (both children arrived in shared memory): pairs[B0 ...] is one contiguous run,
    // written with full 32-byte sectors, two threads per node
    for (int i = tid; i < 2 * (BL - 1); i += BL) {
        const int ls = i >> 1, half = i & 1;
        if (B0 + ls < Bend && s_flag[ls] == 3u) {
            const float4 a = s_dep[ls][half][0], b = s_dep[ls][half][1];
            st256(&pairs[B0 + ls].c[half], a, b);
            if (qpairs) qpairs[B0 + ls].c[half] = qnode_half(a, b, qf);  // neighbouring threads fill the two halves of a sector
        }
    }
    // ---- leftovers are parked for upper_kernel: (1) my own subtree if its parent split lies outside the
    // block, (2) split B0+tid if only one child arrived (the sibling reaches beyond the block)
    uint32_t f = 0;
    if (tid < BL - 1 && B0 + tid < Bend) f = s_flag[tid];
    const bool conv = (f == 1u || f == 2u);
    const uint32_t mine = (pending ? 1u : 0u) + (conv ? 1u : 0u);
    uint32_t at = 0;
    if (mine) at = atomicAdd(&s_npend, mine);
    __syncthreads();
    if (tid == 0 && s_npend) s_base = atomicAdd(list_count, s_npend);
    __syncthreads();
    BT_MARK(4);  // pending-list reservation
#ifdef BT_PROFILE
    if (tid == 0) atomicAdd(&g_bt_prof[8], (unsigned long long)s_npend);
#endif
    if (!mine) return;
    at += s_base;
    if (pending) {
        if (at < capacity) {
            Pending p;
            pack(c, 0, p.a, p.b);
            p.F = c.F; p.L = c.L; p.slot = 0; p.side = -1;
            list[at] = p;
        } else {
            climb_global(c, false, 0, 0, keys, n, flags, pairs, root_box, qpairs, qf);  // list full: climb here
        }
        ++at;
    }
    if (conv) {
        const int side = (f == 1u) ? 0 : 1;
        const float4 a = s_dep[tid][side][0], b = s_dep[tid][side][1];
        const int ext = __float_as_int(b.w);
        const int s = B0 + tid;
        Pending p;
        p.a = a; p.b = b; p.slot = s; p.side = side;
        if (side == 0) { p.F = ext; p.L = s; } else { p.F = s + 1; p.L = ext; }
        if (at < capacity) {
            list[at] = p;
        } else {
            Carry d;
            d.lo[0] = a.x; d.lo[1] = a.y; d.lo[2] = a.z; d.hi[0] = a.w; d.hi[1] = b.x; d.hi[2] = b.y;
            d.link = __float_as_int(b.z);
            d.F = p.F; d.L = p.L;
            climb_global(d, true, s, side, keys, n, flags, pairs, root_box, qpairs, qf);
        }
    }
}

// ---------------------------------------------------------------- renumbering to the reference's indices
// Karras numbering (bvh.cuh:161-198): an internal node that is a LEFT child has the index of the
// last leaf of its range, a RIGHT child the index of its first leaf, the root is 0.
// unified ids: internal split s -> s, leaf j -> (n-1)+j.
__global__ void __launch_bounds__(256)
parent_kernel(const NodePair* __restrict__ pairs, uint32_t n, uint32_t* __restrict__ parent_side,
              uint32_t* __restrict__ refcount) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n - 1) return;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int link = pairs[s].c[side].link;
        const uint32_t u = link >= 0 ? (uint32_t)link : (n - 1) + (uint32_t)~link;
        if (u < 2 * n - 1) {
            parent_side[u] = (s << 1) | (uint32_t)side;
            if (refcount) atomicAdd(refcount + u, 1u);
        }
    }
}

// nodes_out[0..n-2] internal (Karras index), nodes_out[n-1+j] leaf j; see b200cd_node32.
__global__ void __launch_bounds__(256)
export_kernel(const NodePair* __restrict__ pairs, const float* __restrict__ root_box,
              const uint32_t* __restrict__ parent_side, uint32_t n, b200cd_node32* __restrict__ out) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 1) {
        if (s == 0) {
            for (int k = 0; k < 3; ++k) { out[0].lo[k] = root_box[k]; out[0].hi[k] = root_box[3 + k]; }
            out[0].left = out[0].right = -1;
        }
        return;
    }
    if (s >= n - 1) return;
    const uint32_t root = (uint32_t) reinterpret_cast<const int*>(root_box)[6];
    // range of node t: [pairs[t].c[0].ext (= first leaf), pairs[t].c[1].ext (= last leaf)]
    auto karras = [&](uint32_t t) -> uint32_t {
        if (t == root) return 0u;
        const uint32_t side = parent_side[t] & 1u;
        return side == 0 ? (uint32_t)pairs[t].c[1].ext : (uint32_t)pairs[t].c[0].ext;
    };
    const uint32_t me = karras(s);
    int child_no[2];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const Node32 c = pairs[s].c[side];
        const int node = c.link >= 0 ? (int)karras((uint32_t)c.link) : (int)(n - 1) + ~c.link;
        child_no[side] = node;
        for (int k = 0; k < 3; ++k) { out[node].lo[k] = c.lo[k]; out[node].hi[k] = c.hi[k]; }
        if (c.link < 0) out[node].left = out[node].right = -1;
    }
    out[me].left = child_no[0];
    out[me].right = child_no[1];
    if (s == root)
        for (int k = 0; k < 3; ++k) { out[0].lo[k] = root_box[k]; out[0].hi[k] = root_box[3 + k]; }
}

// ---------------------------------------------------------------- structural self-checks (check.cuh:29-96)
__global__ void __launch_bounds__(256)
validate_kernel(const NodePair* __restrict__ pairs, const LeafRec* __restrict__ leaves, const float* __restrict__ root_box,
                const uint32_t* __restrict__ parent_side, const uint32_t* __restrict__ refcount,
                const uint64_t* __restrict__ keys, uint32_t n, uint32_t nverts, uint32_t* __restrict__ chk) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t root = n > 1 ? (uint32_t) reinterpret_cast<const int*>(root_box)[6] : 0u;
    if (t < n) {  // leaf checks
        if (n > 1) {
            if (refcount[n - 1 + t] != 1u) atomicAdd(chk + 4, 1u);  // null (or duplicate) parent, check.cuh:91
            else {
                const uint32_t pw = parent_side[n - 1 + t];
                const Node32 me = pairs[pw >> 1].c[pw & 1u];
                if (me.link != ~(int)t) atomicAdd(chk + 4, 1u);
                if (!(me.lo[0] <= me.hi[0] && me.lo[1] <= me.hi[1] && me.lo[2] <= me.hi[2])) atomicAdd(chk + 6, 1u);
            }
        }
        const LeafRec r = leaves[t];
        if (r.vi[0] >= nverts || r.vi[1] >= nverts || r.vi[2] >= nverts || r.id >= n) atomicAdd(chk + 5, 1u);
        if (t + 1 < n && keys && !(keys[t] < keys[t + 1])) atomicAdd(chk + 7, 1u);
    }
    if (n > 1 && t < n - 1) {  // internal checks
        const uint32_t rc = refcount[t];
        if (rc == 0u) atomicAdd(chk + 0, 1u);  // no parent: exactly one node (the root) may say so, check.cuh:74
        const Node32 L = pairs[t].c[0], R = pairs[t].c[1];
        auto bad_link = [&](int link) { return link >= 0 ? (uint32_t)link >= n - 1 : (uint32_t)~link >= n; };
        const bool badL = bad_link(L.link), badR = bad_link(R.link);
        if (badL) atomicAdd(chk + 2, 1u);
        if (badR) atomicAdd(chk + 2, 1u);
        // "bounded by exactly two children" (check.cuh:73): the children's leaf ranges must tile [first, last]
        bool tiled = !badL && !badR && L.ext <= (int)t && (int)t < R.ext && rc <= 1u && (rc == 1u || t == root);
        if (tiled) {
            if (L.link >= 0) tiled = pairs[L.link].c[0].ext == L.ext && pairs[L.link].c[1].ext == (int)t;
            else tiled = ~L.link == (int)t && L.ext == (int)t;
        }
        if (tiled) {
            if (R.link >= 0) tiled = pairs[R.link].c[0].ext == (int)t + 1 && pairs[R.link].c[1].ext == R.ext;
            else tiled = ~R.link == (int)t + 1 && R.ext == (int)t + 1;
        }
        if (!tiled) atomicAdd(chk + 1, 1u);
        const bool okL = L.lo[0] <= L.hi[0] && L.lo[1] <= L.hi[1] && L.lo[2] <= L.hi[2];
        const bool okR = R.lo[0] <= R.hi[0] && R.lo[1] <= R.hi[1] && R.lo[2] <= R.hi[2];
        if (!okL || !okR) atomicAdd(chk + 3, 1u);
        // my box (stored in my parent, or the root box) must enclose my children
        float mlo[3], mhi[3];
        bool have = false;
        if (t == root) {
            for (int k = 0; k < 3; ++k) { mlo[k] = root_box[k]; mhi[k] = root_box[3 + k]; }
            have = true;
        } else if (rc == 1u) {
            const uint32_t pw = parent_side[t];
            const Node32 me = pairs[pw >> 1].c[pw & 1u];
            for (int k = 0; k < 3; ++k) { mlo[k] = me.lo[k]; mhi[k] = me.hi[k]; }
            have = me.link == (int)t;
            if (!have) atomicAdd(chk + 8, 1u);
        }
        if (have) {
            bool enc = true;
            for (int k = 0; k < 3; ++k)
                enc = enc && mlo[k] <= L.lo[k] && mlo[k] <= R.lo[k] && mhi[k] >= L.hi[k] && mhi[k] >= R.hi[k];
            if (!enc) atomicAdd(chk + 8, 1u);
        }
    }
}

}  // namespace

uint64_t build_tree_scratch_bytes(uint32_t n) {  // pending list + its counter
    return 16 + sizeof(Pending) * (uint64_t)build_tree_pending_capacity(n);
}
uint32_t build_tree_pending_capacity(uint32_t n) { return n / 8 + 4096; }

void launch_build_tree(const float4* d_verts, const uint32_t* d_idx, const uint32_t* d_sorted_ids, const uint64_t* d_keys,
                       uint32_t n, uint32_t* d_flags, NodePair* d_pairs, LeafRec* d_leaves, float* d_root_box,
                       void* d_scratch, cudaStream_t s, const LeafRec* d_recs, float* d_block_boxes, QNodePair* d_qpairs,
                       float* d_qframe, const QFrame* frame, bool frame_from_root) {
    if (!n) return;
    if (!frame || !d_qframe) d_qpairs = nullptr;
    if (d_qpairs) {  // the grid of the quantised nodes: over the PREVIOUS build's root box when there is one, else over `frame`
        qframe_kernel<<<1, 32, 0, s>>>(d_root_box, frame_from_root ? 1 : 0, *frame, d_qframe);
        count_launch();
    }
    uint32_t* list_count = static_cast<uint32_t*>(d_scratch);
    Pending* list = reinterpret_cast<Pending*>(static_cast<char*>(d_scratch) + 16);
    const uint32_t capacity = build_tree_pending_capacity(n);
    if (n > 1) cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * (size_t)(n - 1), s);
    cudaMemsetAsync(list_count, 0, sizeof(uint32_t), s);
    if (d_recs)
        build_kernel<true><<<(n + BL - 1) / BL, BL, 0, s>>>(d_verts, d_idx, d_sorted_ids, d_keys, (int)n, d_flags, d_pairs, d_leaves,
                                                            d_root_box, list, list_count, capacity, d_recs, d_block_boxes, d_qpairs,
                                                            d_qframe);
    else
        build_kernel<false><<<(n + BL - 1) / BL, BL, 0, s>>>(d_verts, d_idx, d_sorted_ids, d_keys, (int)n, d_flags, d_pairs,
                                                             d_leaves, d_root_box, list, list_count, capacity, nullptr, d_block_boxes,
                                                             d_qpairs, d_qframe);
    count_launch();
    trace_mark("build_kernel", s);
    {
        upper_kernel<<<(capacity + 127) / 128, 128, 0, s>>>(list, list_count, capacity, d_keys, (int)n, d_flags, d_pairs,
                                                            d_root_box, d_qpairs, d_qframe);
        count_launch();
        trace_mark("upper_kernel", s);
    }
}

// d_scratch: 2 * (2n-1) words (parent_side, refcount)
void launch_export_nodes(const NodePair* d_pairs, const float* d_root_box, uint32_t n, uint32_t* d_scratch,
                         b200cd_node32* d_nodes_out, cudaStream_t s) {
    if (!n) return;
    const uint32_t work = n > 1 ? n - 1 : 1;
    if (n > 1) {
        parent_kernel<<<(work + 255) / 256, 256, 0, s>>>(d_pairs, n, d_scratch, nullptr);
        count_launch();
    }
    export_kernel<<<(work + 255) / 256, 256, 0, s>>>(d_pairs, d_root_box, d_scratch, n, d_nodes_out);
    count_launch();
}

void launch_validate(const NodePair* d_pairs, const LeafRec* d_leaves, const float* d_root_box, const uint64_t* d_keys,
                     uint32_t n, uint32_t nverts, uint32_t* d_scratch, uint32_t* d_checks9, cudaStream_t s) {
    cudaMemsetAsync(d_checks9, 0, 9 * sizeof(uint32_t), s);
    if (!n) return;
    uint32_t* parent_side = d_scratch;
    uint32_t* refcount = d_scratch + (2ull * n - 1);
    cudaMemsetAsync(d_scratch, 0, sizeof(uint32_t) * 2 * (2ull * n - 1), s);
    if (n > 1) {
        parent_kernel<<<(n - 1 + 255) / 256, 256, 0, s>>>(d_pairs, n, parent_side, refcount);
        count_launch();
    }
    validate_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_pairs, d_leaves, d_root_box, parent_side, refcount, d_keys, n, nverts,
                                                   d_checks9);
    count_launch();
}

}  // namespace b200cd
