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
//     own position (Node32::last <= q) is pruned. The AABB predicate is symmetric,
//     so this halves the traversal without changing the set; the narrow phase is
//     still called as (lower ID, higher ID) because the SAT is not symmetric in
//     floating point.
//   - traversal only emits CANDIDATES (AABB-overlapping leaf pairs) into a compact
//     list: per-warp staging in shared memory filled with ballots (no atomics),
//     flushed with one global atomicAdd per >=32 candidates. The divergent fp64
//     SAT runs afterwards in its own kernel, in two warp-dense stages (face-normal axes first,
//     survivors compacted through a per-warp shared-memory queue, then the other 15 axes).
//   - results are appended the same way (warp-aggregated atomic).
// -fmad=false: every double operation rounds separately, like the host reference.
#include "common.cuh"

namespace b200cd {

namespace {

constexpr int BR_THREADS = 128;
constexpr int BR_WARPS = BR_THREADS / 32;
constexpr int BR_QUEUE = 128;   // per-warp candidate staging (uint2 each)
constexpr int BR_FLUSH = 64;    // flush when at least this many are staged (<= 64 arrive per step)

constexpr unsigned long long ERR_STACK = 1ull;

// strict overlap, box.cuh:40-43: (a.lo - b.hi) * (b.lo - a.hi) > 0 on every axis. For
// well-formed boxes of fp32-exact values that is exactly a.lo < b.hi && b.lo < a.hi.
__device__ __forceinline__ bool overlap(const float qlo[3], const float qhi[3], float lx, float ly, float lz,
                                        float hx, float hy, float hz) {
    return qlo[0] < hx && lx < qhi[0] && qlo[1] < hy && ly < qhi[1] && qlo[2] < hz && lz < qhi[2];
}

__device__ __forceinline__ float min3_ref(float a, float b, float c) { float t = a; if (b < t) t = b; if (c < t) t = c; return t; }
__device__ __forceinline__ float max3_ref(float a, float b, float c) { float t = a; if (b > t) t = b; if (c > t) t = c; return t; }

// ---------------------------------------------------------------- K5
__global__ void __launch_bounds__(BR_THREADS)
broad_kernel(const NodePair* __restrict__ pairs, const LeafRec* __restrict__ leaves, uint32_t n, uint32_t shard,
             uint32_t nshards, uint32_t chunk, uint32_t nquery, uint2* __restrict__ cand, uint64_t cand_cap,
             unsigned long long* __restrict__ counters) {
    __shared__ uint2 queue[BR_WARPS][BR_QUEUE];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint2* wq = queue[warp];
    uint32_t staged = 0;  // warp-uniform

    // which query (sorted leaf position) this thread owns: block-cyclic over shards
    const uint32_t t = blockIdx.x * BR_THREADS + threadIdx.x;
    uint32_t q = 0xffffffffu;
    if (t < nquery) {
        uint32_t c = t / chunk, w = t - c * chunk;
        uint64_t qq = ((uint64_t)c * nshards + shard) * chunk + w;
        if (qq < n) q = (uint32_t)qq;
    }
    float qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    int node = -1;
    if (q != 0xffffffffu && q + 1 < n) {  // the last leaf has no partner with a larger position
        const float4* r = reinterpret_cast<const float4*>(leaves + q);
        const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
        // v0 = r0.xyz, v1 = (r0.w, r1.x, r1.y), v2 = (r1.z, r1.w, r2.x); box.cuh:13-22
        qlo[0] = min3_ref(r0.x, r0.w, r1.z); qhi[0] = max3_ref(r0.x, r0.w, r1.z);
        qlo[1] = min3_ref(r0.y, r1.x, r1.w); qhi[1] = max3_ref(r0.y, r1.x, r1.w);
        qlo[2] = min3_ref(r0.z, r1.y, r2.x); qhi[2] = max3_ref(r0.z, r1.y, r2.x);
        node = 0;
    }
    int stack[B200CD_MAX_STACK];
    int sp = 0;
    bool overflow = false;

    while (__any_sync(0xffffffffu, node >= 0)) {
        bool candL = false, candR = false;
        int leafL = 0, leafR = 0;
        if (node >= 0) {
            const float4* p = reinterpret_cast<const float4*>(pairs + node);
            const float4 a0 = __ldg(p), a1 = __ldg(p + 1), b0 = __ldg(p + 2), b1 = __ldg(p + 3);
            const int linkL = __float_as_int(a1.z), lastL = __float_as_int(a1.w);
            const int linkR = __float_as_int(b1.z), lastR = __float_as_int(b1.w);
            const bool hitL = lastL > (int)q && overlap(qlo, qhi, a0.x, a0.y, a0.z, a0.w, a1.x, a1.y);
            const bool hitR = lastR > (int)q && overlap(qlo, qhi, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y);
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
        // stage candidates (q, leaf) — positions by ballot, no atomics
        const uint32_t bL = __ballot_sync(0xffffffffu, candL), bR = __ballot_sync(0xffffffffu, candR);
        if (bL | bR) {
            const uint32_t nL = __popc(bL);
            if (candL) wq[staged + __popc(bL & lt)] = make_uint2(q, (uint32_t)leafL);
            if (candR) wq[staged + nL + __popc(bR & lt)] = make_uint2(q, (uint32_t)leafR);
            staged += nL + __popc(bR);
            __syncwarp();
            if (staged >= BR_FLUSH) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counters + 0, (unsigned long long)staged);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (uint32_t i = lane; i < staged; i += 32)
                    if (base + i < cand_cap) __stcs(cand + base + i, wq[i]);
                staged = 0;
                __syncwarp();
            }
        }
    }
    if (staged) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counters + 0, (unsigned long long)staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < staged; i += 32)
            if (base + i < cand_cap) __stcs(cand + base + i, wq[i]);
    }
    if (overflow) atomicOr(counters + 2, ERR_STACK);
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

struct Tri {
    D3 v0, v1, v2;
    uint32_t i0, i1, i2, id;
};
__device__ __forceinline__ Tri load_tri(const LeafRec* __restrict__ leaves, uint32_t pos) {
    const float4* r = reinterpret_cast<const float4*>(leaves + pos);
    const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
    Tri t;
    t.v0 = {(double)r0.x, (double)r0.y, (double)r0.z};
    t.v1 = {(double)r0.w, (double)r1.x, (double)r1.y};
    t.v2 = {(double)r1.z, (double)r1.w, (double)r2.x};
    t.i0 = __float_as_uint(r2.y); t.i1 = __float_as_uint(r2.z); t.i2 = __float_as_uint(r2.w);
    t.id = __float_as_uint(r3.x);
    return t;
}
// vertices only (stage B re-reads the two records; they are L1/L2-hot)
__device__ __forceinline__ void load_verts(const LeafRec* __restrict__ leaves, uint32_t pos, D3& v0, D3& v1, D3& v2,
                                           uint32_t& id) {
    const float4* r = reinterpret_cast<const float4*>(leaves + pos);
    const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
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
__global__ void __launch_bounds__(NR_THREADS, 2)
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
            const Tri a = load_tri(leaves, c.x), b = load_tri(leaves, c.y);
            // triangle.cuh:18-30: neighborCount >= 1 <=> any vertex index shared
            const bool shared = a.i0 == b.i0 || a.i0 == b.i1 || a.i0 == b.i2 || a.i1 == b.i0 || a.i1 == b.i1 ||
                                a.i1 == b.i2 || a.i2 == b.i0 || a.i2 == b.i1 || a.i2 == b.i2;
            if (!shared && a.id != b.id) {
                // tri_contact.cuh:81-86: P is the lower-ID triangle
                if (a.id < b.id) { alive = sat_stage_a(sat_input(a.v0, a.v1, a.v2, b.v0, b.v1, b.v2)); entry = make_uint2(c.x, c.y); }
                else             { alive = sat_stage_a(sat_input(b.v0, b.v1, b.v2, a.v0, a.v1, a.v2)); entry = make_uint2(c.y, c.x); }
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

void launch_broad(const NodePair* d_pairs, const LeafRec* d_leaves, uint32_t n, uint32_t shard, uint32_t nshards,
                  uint32_t chunk, uint32_t nquery, uint2* d_cand, uint64_t cand_cap, unsigned long long* d_counters,
                  cudaStream_t s) {
    if (n < 2 || nquery == 0) return;
    uint32_t blocks = (nquery + BR_THREADS - 1) / BR_THREADS;
    broad_kernel<<<blocks, BR_THREADS, 0, s>>>(d_pairs, d_leaves, n, shard, nshards, chunk, nquery, d_cand, cand_cap,
                                               d_counters);
    count_launch();
}

void launch_narrow(const LeafRec* d_leaves, const uint2* d_cand, uint64_t cand_cap, uint2* d_out, uint64_t out_cap,
                   unsigned long long* d_counters, int sms, cudaStream_t s) {
    narrow_kernel<<<sms * 2 * 4, NR_THREADS, 0, s>>>(d_leaves, d_cand, cand_cap, d_out, out_cap, d_counters);
    count_launch();
}

}  // namespace b200cd
