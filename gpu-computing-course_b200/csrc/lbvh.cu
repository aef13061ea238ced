// lbvh.cu — K3 (Karras hierarchy) and K4 (leaf records + bottom-up AABB refit),
// plus the parity/validation helpers (node export, structural self-checks).
//
// Reference semantics (under /root/reference/CollisionDetection/):
//   K3  bvh.cuh:48 (delta), :100-123 (determineRange), :57-98 (findSplit),
//       :146-199 (generateHierarchyParallel): internal node i covers the key range
//       that contains i and its more-similar neighbour; children are
//       leaf/internal by "split == first" / "split+1 == last"; root = internal 0.
//   K4  bvh.cuh:258-285 (calBoundingBox), box.cuh:13-32 (Box::set / Box::merge),
//       mathop.cuh:17-44 (comparison-based min/max).
//
// What is different here, by design:
//   - delta() breaks ties between equal keys with the leaf index (Karras 2012 §4);
//     the reference builds a malformed tree for duplicate codes (load_obj.h:110-115
//     only reports them). With unique keys the tree is node-for-node identical.
//   - no pointer-linked 112-byte Node (bvh.cuh:25-43) and no fillLeafNodes pass
//     (bvh.cuh:125-144): K3 writes only a parent word per node; K4 writes the
//     traversal layout directly — the two children of internal node p live side by
//     side in pairs[p] (64 B), each a 32-byte Node32 {box, link, last}.
//   - the refit publishes a child's box with __threadfence() before the arrival
//     atomic and reads the sibling through L2 (__ldcg); the reference has neither
//     (bvh.cuh:270-278) and relies on luck.
#include "common.cuh"

namespace b200cd {

namespace {

constexpr uint32_t ROOT_PARENT = 0xffffffffu;

// ---------------------------------------------------------------- K3
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, uint64_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t x = ki ^ __ldg(keys + j);
    return x ? __clzll((long long)x) : 64 + __clz(i ^ j);
}

__global__ void __launch_bounds__(256) hierarchy_kernel(const uint64_t* __restrict__ keys, int n,
                                                       uint32_t* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const uint64_t ki = __ldg(keys + i);
    // direction of the range and the prefix length it must beat
    int dn = delta(keys, n, i, ki, i + 1), dp = delta(keys, n, i, ki, i - 1);
    int d = (dn - dp) >= 0 ? 1 : -1;
    int dmin = d > 0 ? dp : dn;
    // exponential search for an upper bound of the range length, then binary search
    int lmax = 2;
    while (delta(keys, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int first = min(i, j), last = max(i, j);
    // split: highest position in [first, last) sharing more than the range's common prefix with `first`
    uint64_t kf = __ldg(keys + first);
    int common = delta(keys, n, first, kf, last);
    int split = first, step = last - first;
    do {
        step = (step + 1) >> 1;
        int cand = split + step;
        if (cand < last && delta(keys, n, first, kf, cand) > common) split = cand;
    } while (step > 1);
    // parent words: internal nodes first (n-1 of them), then leaves
    uint32_t a = (split == first) ? (uint32_t)(n - 1 + split) : (uint32_t)split;
    uint32_t b = (split + 1 == last) ? (uint32_t)(n - 1 + split + 1) : (uint32_t)(split + 1);
    parent[a] = ((uint32_t)i << 1);
    parent[b] = ((uint32_t)i << 1) | 1u;
    if (i == 0) parent[0] = ROOT_PARENT;
}

// ---------------------------------------------------------------- K4
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

__device__ __forceinline__ void store_node(Node32* dst, const float lo[3], const float hi[3], int link, int last) {
    float4* p = reinterpret_cast<float4*>(dst);
    __stcg(p, make_float4(lo[0], lo[1], lo[2], hi[0]));
    __stcg(p + 1, make_float4(hi[1], hi[2], __int_as_float(link), __int_as_float(last)));
}

__global__ void __launch_bounds__(256)
refit_kernel(const float4* __restrict__ verts, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ sorted_ids,
             uint32_t n, const uint32_t* __restrict__ parent, uint32_t* __restrict__ flags, NodePair* __restrict__ pairs,
             LeafRec* __restrict__ leaves, float* __restrict__ root_box) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    // leaf record: the triangle at sorted position j, vertices copied next to their indices and ID
    const uint32_t id = __ldg(sorted_ids + j);
    const uint32_t* f = idx + 3ull * id;
    const uint32_t i0 = __ldg(f), i1 = __ldg(f + 1), i2 = __ldg(f + 2);
    const float4 a = __ldg(verts + i0), b = __ldg(verts + i1), c = __ldg(verts + i2);
    float4* rec = reinterpret_cast<float4*>(leaves + j);
    __stcs(rec, make_float4(a.x, a.y, a.z, b.x));
    __stcs(rec + 1, make_float4(b.y, b.z, c.x, c.y));
    __stcs(rec + 2, make_float4(c.z, __uint_as_float(i0), __uint_as_float(i1), __uint_as_float(i2)));
    __stcs(rec + 3, make_float4(__uint_as_float(id), 0.f, 0.f, 0.f));
    // leaf box, box.cuh:13-22
    float lo[3] = {min3_ref(a.x, b.x, c.x), min3_ref(a.y, b.y, c.y), min3_ref(a.z, b.z, c.z)};
    float hi[3] = {max3_ref(a.x, b.x, c.x), max3_ref(a.y, b.y, c.y), max3_ref(a.z, b.z, c.z)};
    if (n == 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { root_box[k] = lo[k]; root_box[3 + k] = hi[k]; }
        return;
    }
    int link = ~(int)j, last = (int)j;
    uint32_t pw = __ldg(parent + (n - 1) + j);
    // climb: the first thread to reach a node stops, the second merges (bvh.cuh:269-283)
    while (true) {
        const uint32_t p = pw >> 1, side = pw & 1u;
        store_node(&pairs[p].c[side], lo, hi, link, last);
        __threadfence();  // publish my half before announcing arrival
        if (atomicAdd(flags + p, 1u) == 0u) return;
        __threadfence();
        // the sibling's half was published before its atomic; read it through L2
        const float4* sp = reinterpret_cast<const float4*>(&pairs[p].c[side ^ 1u]);
        const float4 s0 = __ldcg(sp), s1 = __ldcg(sp + 1);
        const float slo[3] = {s0.x, s0.y, s0.z}, shi[3] = {s0.w, s1.x, s1.y};
        const int slast = __float_as_int(s1.w);
        // Box::merge(childA, childB), box.cuh:24-32 — operand order A then B
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float alo = side ? slo[k] : lo[k], blo = side ? lo[k] : slo[k];
            float ahi = side ? shi[k] : hi[k], bhi = side ? hi[k] : shi[k];
            lo[k] = min2_ref(alo, blo);
            hi[k] = max2_ref(ahi, bhi);
        }
        last = max(last, slast);
        link = (int)p;
        pw = __ldg(parent + p);
        if (pw == ROOT_PARENT) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { root_box[k] = lo[k]; root_box[3 + k] = hi[k]; }
            return;
        }
    }
}

// ---------------------------------------------------------------- parity export
// nodes_out[0..n-2] internal (Karras index), nodes_out[n-1+j] leaf j; see b200cd_node32.
__global__ void __launch_bounds__(256)
export_kernel(const NodePair* __restrict__ pairs, const float* __restrict__ root_box, uint32_t n,
              b200cd_node32* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 1) {
        if (p == 0) {
            for (int k = 0; k < 3; ++k) { out[0].lo[k] = root_box[k]; out[0].hi[k] = root_box[3 + k]; }
            out[0].left = out[0].right = -1;
        }
        return;
    }
    if (p >= n - 1) return;
    int child_no[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const Node32 c = pairs[p].c[s];
        int node = c.link >= 0 ? c.link : (int)(n - 1) + ~c.link;
        child_no[s] = node;
        for (int k = 0; k < 3; ++k) { out[node].lo[k] = c.lo[k]; out[node].hi[k] = c.hi[k]; }
        if (c.link < 0) out[node].left = out[node].right = -1;
    }
    out[p].left = child_no[0];
    out[p].right = child_no[1];
    if (p == 0)
        for (int k = 0; k < 3; ++k) { out[0].lo[k] = root_box[k]; out[0].hi[k] = root_box[3 + k]; }
}

// ---------------------------------------------------------------- structural self-checks (check.cuh:29-96)
__global__ void __launch_bounds__(256)
validate_kernel(const NodePair* __restrict__ pairs, const LeafRec* __restrict__ leaves,
                const uint32_t* __restrict__ parent, const uint32_t* __restrict__ flags,
                const uint64_t* __restrict__ keys, uint32_t n, uint32_t nverts, uint32_t* __restrict__ chk) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {  // leaf checks
        if (n > 1) {
            uint32_t pw = parent[n - 1 + t];
            if (pw == ROOT_PARENT || (pw >> 1) >= n - 1) atomicAdd(chk + 4, 1u);
            else {
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
        uint32_t pw = parent[t];
        if (pw == ROOT_PARENT) atomicAdd(chk + 0, 1u);
        if (flags[t] != 2u) atomicAdd(chk + 1, 1u);
        const Node32 L = pairs[t].c[0], R = pairs[t].c[1];
        auto bad_link = [&](int link) { return link >= 0 ? (uint32_t)link >= n - 1 : (uint32_t)~link >= n; };
        if (bad_link(L.link)) atomicAdd(chk + 2, 1u);
        if (bad_link(R.link)) atomicAdd(chk + 2, 1u);
        bool okL = L.lo[0] <= L.hi[0] && L.lo[1] <= L.hi[1] && L.lo[2] <= L.hi[2];
        bool okR = R.lo[0] <= R.hi[0] && R.lo[1] <= R.hi[1] && R.lo[2] <= R.hi[2];
        if (!okL || !okR) atomicAdd(chk + 3, 1u);
        if (pw != ROOT_PARENT && (pw >> 1) < n - 1) {  // my box, stored in my parent, must enclose my children
            const Node32 me = pairs[pw >> 1].c[pw & 1u];
            bool enc = true;
            for (int k = 0; k < 3; ++k)
                enc = enc && me.lo[k] <= L.lo[k] && me.lo[k] <= R.lo[k] && me.hi[k] >= L.hi[k] && me.hi[k] >= R.hi[k];
            if (!enc || me.link != (int)t) atomicAdd(chk + 8, 1u);
        }
    }
}

}  // namespace

void launch_hierarchy(const uint64_t* d_keys, uint32_t n, uint32_t* d_parent, cudaStream_t s) {
    if (n < 2) return;
    hierarchy_kernel<<<(n - 1 + 255) / 256, 256, 0, s>>>(d_keys, (int)n, d_parent);
    count_launch();
}

void launch_refit(const float4* d_verts, const uint32_t* d_idx, const uint32_t* d_sorted_ids, uint32_t n,
                  const uint32_t* d_parent, uint32_t* d_flags, NodePair* d_pairs, LeafRec* d_leaves,
                  float* d_root_box, cudaStream_t s) {
    if (!n) return;
    if (n > 1) cudaMemsetAsync(d_flags, 0, sizeof(uint32_t) * (size_t)(n - 1), s);
    refit_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_verts, d_idx, d_sorted_ids, n, d_parent, d_flags, d_pairs,
                                                 d_leaves, d_root_box);
    count_launch();
}

void launch_export_nodes(const NodePair* d_pairs, const LeafRec*, const float* d_root_box, uint32_t n,
                         b200cd_node32* d_nodes_out, cudaStream_t s) {
    if (!n) return;
    uint32_t work = n > 1 ? n - 1 : 1;
    export_kernel<<<(work + 255) / 256, 256, 0, s>>>(d_pairs, d_root_box, n, d_nodes_out);
    count_launch();
}

void launch_validate(const NodePair* d_pairs, const LeafRec* d_leaves, const uint32_t* d_parent,
                     const uint32_t* d_flags, const uint64_t* d_keys, uint32_t n, uint32_t nverts,
                     uint32_t* d_checks9, cudaStream_t s) {
    cudaMemsetAsync(d_checks9, 0, 9 * sizeof(uint32_t), s);
    if (!n) return;
    validate_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_pairs, d_leaves, d_parent, d_flags, d_keys, n, nverts,
                                                   d_checks9);
    count_launch();
}

}  // namespace b200cd
