// morton.cu — vertex expansion, mesh bounding box and K1: per-triangle centroid + Morton key.
//
// Reference semantics (all under /root/reference/CollisionDetection/):
//   centroid   load_obj.h:89-90   c = (p1 + p2 + p3) / 3 per axis, in double, left to right
//   normalise  morton.h:43-58     (c - origin) / extent
//   scale      morton.h:73-76     * 2^20, implicit double -> u64 truncation (morton.h:80-82)
//   spread     morton.h:7-29      21-bit mask then magic-mask bit spread
//   interleave morton.h:86        x<<2 | y<<1 | z
// In the reference this all runs on the HOST while parsing; here it is one HBM-bound
// kernel: 12 B of indices + three float4 vertex gathers in, one 8-B key out.
// Compiled with -fmad=false, so every fp64 op rounds exactly like the host code.
#include <algorithm>

#include "common.cuh"

namespace b200cd {

// ---- float3 (caller layout) -> float4 (gather-friendly, one 16-B load per vertex)
__global__ void __launch_bounds__(256) expand_verts_kernel(const float* __restrict__ xyz,
                                                          float4* __restrict__ verts, uint32_t nverts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nverts) return;
    const float* p = xyz + 3ull * i;
    verts[i] = make_float4(p[0], p[1], p[2], 0.0f);
}

void launch_expand_verts(const float* d_xyz, float4* d_verts, uint32_t nverts, cudaStream_t s) {
    if (!nverts) return;
    expand_verts_kernel<<<(nverts + 255) / 256, 256, 0, s>>>(d_xyz, d_verts, nverts);
    count_launch();
}

// ---- index sanity: any vertex index >= nverts raises *flag (checked once at upload, on the device)
__global__ void __launch_bounds__(256) check_idx_kernel(const uint32_t* __restrict__ idx, uint64_t count,
                                                       uint32_t nverts, uint32_t* __restrict__ flag) {
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
        bad |= __ldg(idx + i) >= nverts;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

void launch_check_idx(const uint32_t* d_idx, uint32_t ntris, uint32_t nverts, uint32_t* d_flag, int sms, cudaStream_t s) {
    cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), s);
    if (!ntris) return;
    uint64_t count = 3ull * ntris;
    uint32_t blocks = (uint32_t)std::min<uint64_t>((count + 255) / 256, (uint64_t)sms * 16);
    check_idx_kernel<<<blocks, 256, 0, s>>>(d_idx, count, nverts, d_flag);
    count_launch();
}

// same check without clearing the flag first: several slices OR their verdicts into one word
void launch_check_idx_accumulate(const uint32_t* d_idx, uint32_t ntris, uint32_t nverts, uint32_t* d_flag, int sms, cudaStream_t s) {
    if (!ntris) return;
    uint64_t count = 3ull * ntris;
    uint32_t blocks = (uint32_t)std::min<uint64_t>((count + 255) / 256, (uint64_t)sms * 16);
    check_idx_kernel<<<blocks, 256, 0, s>>>(d_idx, count, nverts, d_flag);
    count_launch();
}

// ---- bounding box of all vertices (auto Morton box). Order-preserving uint encoding
// of floats lets plain integer atomicMin/atomicMax do the reduction.
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ verts, uint32_t nverts,
                                                  uint32_t* __restrict__ bbox6) {
    uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nverts; i += gridDim.x * blockDim.x) {
        float4 v = __ldg(verts + i);
        uint32_t o[3] = {f2ord(v.x), f2ord(v.y), f2ord(v.z)};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = min(lo[a], o[a]);
            hi[a] = max(hi[a], o[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(bbox6 + a, lo[a]);
            atomicMax(bbox6 + 3 + a, hi[a]);
        }
    }
}

void launch_bbox(const float4* d_verts, uint32_t nverts, uint32_t* d_bbox6, int sms, cudaStream_t s) {
    cudaMemsetAsync(d_bbox6, 0xff, 12, s);
    cudaMemsetAsync(d_bbox6 + 3, 0x00, 12, s);
    if (!nverts) return;
    uint32_t blocks = min((nverts + 255u) / 256u, (uint32_t)sms * 8u);
    bbox_kernel<<<blocks, 256, 0, s>>>(d_verts, nverts, d_bbox6);
    count_launch();
}

// ---- K1
__device__ __forceinline__ uint64_t spread21(uint64_t v) {  // morton.h:7-29
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}
__device__ __forceinline__ uint32_t spread10(uint32_t v) {  // morton.h:31-40
    v &= 0x3ffu;
    v = (v | v << 16) & 0x30000ffu;
    v = (v | v << 8) & 0x300f00fu;
    v = (v | v << 4) & 0x30c30c3u;
    v = (v | v << 2) & 0x9249249u;
    return v;
}
// double -> u64 like the host conversion for in-range values; non-positive / NaN -> 0
__device__ __forceinline__ uint64_t trunc_u64(double e) {
    if (!(e > 0.0)) return 0ull;
    if (e >= 18446744073709551616.0) return ~0ull;
    return (uint64_t)e;
}

struct MortonBox {
    double o[3], e[3];
};

// HIST: the grid strides over the triangles and every block also counts, in shared memory, the digits the radix sort's
// passes will look at (plan), adding its counts to the global histograms once at the end - the sort then skips its
// own histogram kernel, i.e. one more read of all keys (K1 is bandwidth-bound: the shared-memory atomics ride along).
template <int KEY_BITS, bool HIST>
__global__ void __launch_bounds__(256)
morton_kernel(const float4* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t first, uint32_t n, MortonBox box,
              const uint32_t* __restrict__ bbox6, uint64_t* __restrict__ keys, LeafRec* __restrict__ recs,
              RadixHistPlan plan, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[HIST ? 8 * 256 : 1];
    if (HIST) {
        for (int i = threadIdx.x; i < plan.npass * 256; i += 256) s_hist[i] = 0;
        __syncthreads();
    }
  // triangles first .. first+n-1 (a slice when the build is partitioned over GPUs); keys[] is slice-relative
  for (uint32_t tl = blockIdx.x * blockDim.x + threadIdx.x; tl < n; tl += gridDim.x * blockDim.x) {
    const uint32_t t = first + tl;
    if (bbox6) {  // auto box: origin = bbox.lo, extent = hi - lo (1 if degenerate)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            double lo = (double)ord2f(__ldg(bbox6 + a)), hi = (double)ord2f(__ldg(bbox6 + 3 + a));
            double ext = hi - lo;
            box.o[a] = lo;
            box.e[a] = ext > 0.0 ? ext : 1.0;
        }
    }
    const uint32_t* f = idx + 3ull * t;
    const uint32_t i0 = __ldg(f), i1 = __ldg(f + 1), i2 = __ldg(f + 2);
    float4 a = __ldg(verts + i0), b = __ldg(verts + i1), c = __ldg(verts + i2);
    if (recs) {
        // The triangle's leaf record, written here in FACE order while the indices and vertices stream by: the tree
        // build then moves ONE aligned 64-byte record per leaf into sorted order instead of gathering 12 B of
        // indices + 3 x 16 B of vertices from random places (measured on the 16 M soup: ~290 B of DRAM reads per
        // leaf for 60 B of payload).
        float4* rec = reinterpret_cast<float4*>(recs + tl);
        st256(rec, make_float4(a.x, a.y, a.z, b.x), make_float4(b.y, b.z, c.x, c.y));
        st256(rec + 2, make_float4(c.z, __uint_as_float(i0), __uint_as_float(i1), __uint_as_float(i2)),
              make_float4(__uint_as_float(t), 0.f, 0.f, 0.f));
    }
    double cx = ((double)a.x + (double)b.x + (double)c.x) / 3.0;
    double cy = ((double)a.y + (double)b.y + (double)c.y) / 3.0;
    double cz = ((double)a.z + (double)b.z + (double)c.z) / 3.0;
    double nx = (cx - box.o[0]) / box.e[0];
    double ny = (cy - box.o[1]) / box.e[1];
    double nz = (cz - box.o[2]) / box.e[2];
    uint64_t key;
    if (KEY_BITS == 30) {
        uint32_t xx = spread10((uint32_t)(trunc_u64(nx * 1024.0) & 0x3ffu));
        uint32_t yy = spread10((uint32_t)(trunc_u64(ny * 1024.0) & 0x3ffu));
        uint32_t zz = spread10((uint32_t)(trunc_u64(nz * 1024.0) & 0x3ffu));
        key = (uint64_t)((xx << 2) | (yy << 1) | zz);
    } else {
        uint64_t xx = spread21(trunc_u64(nx * 1048576.0));
        uint64_t yy = spread21(trunc_u64(ny * 1048576.0));
        uint64_t zz = spread21(trunc_u64(nz * 1048576.0));
        key = (xx << 2) | (yy << 1) | zz;
    }
    keys[tl] = key;
    if (HIST) {
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if (p < plan.npass) atomicAdd(&s_hist[p * 256 + (uint32_t)((key >> plan.shift[p]) & plan.mask[p])], 1u);
    }
  }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < plan.npass * 256; i += 256) {
            const uint32_t c = s_hist[i];
            if (c) atomicAdd(&hist[i], c);
        }
    }
}

void launch_morton(const float4* d_verts, const uint32_t* d_idx, uint32_t first, uint32_t n, const b200cd_params& p,
                   const uint32_t* d_bbox6_or_null, uint64_t* d_keys, cudaStream_t s, LeafRec* d_recs,
                   const RadixHistPlan* hist_plan, uint32_t* d_hist, int sms) {
    if (!n) return;
    MortonBox box;
    for (int a = 0; a < 3; ++a) {
        box.o[a] = p.morton_origin[a];
        box.e[a] = p.morton_extent[a];
    }
    uint32_t blocks = (n + 255) / 256;
    RadixHistPlan none{};
    if (hist_plan && d_hist) {  // fused digit histograms: a fixed grid strides over the triangles (one flush per block)
        blocks = std::min<uint32_t>(blocks, (uint32_t)std::max(sms, 1) * 16u);
        if (p.key_bits == 30)
            morton_kernel<30, true><<<blocks, 256, 0, s>>>(d_verts, d_idx, first, n, box, d_bbox6_or_null, d_keys, d_recs, *hist_plan, d_hist);
        else
            morton_kernel<63, true><<<blocks, 256, 0, s>>>(d_verts, d_idx, first, n, box, d_bbox6_or_null, d_keys, d_recs, *hist_plan, d_hist);
    } else if (p.key_bits == 30) {
        morton_kernel<30, false><<<blocks, 256, 0, s>>>(d_verts, d_idx, first, n, box, d_bbox6_or_null, d_keys, d_recs, none, nullptr);
    } else {
        morton_kernel<63, false><<<blocks, 256, 0, s>>>(d_verts, d_idx, first, n, box, d_bbox6_or_null, d_keys, d_recs, none, nullptr);
    }
    count_launch();
    trace_mark("morton_kernel", s);
}

}  // namespace b200cd
