// unique.cu — the second half of the reference's output: the sorted set of the IDs of all triangles that take
// part in a colliding pair. Reference: makeAndPrintSet, main.cu:33-45 (a std::set<unsigned> filled from the pair
// list on the host and printed in ascending order, main.cu:154). Here it stays on the device:
//   1. unique_mark_kernel   one bit per triangle ID: atomicOr of both IDs of every pair into a bitmap
//   2. unique_count_kernel  population count of every 1024-word block of the bitmap; the LAST block to finish
//                           turns the block counts into exclusive offsets (one pass, no second launch)
//   3. unique_emit_kernel   every block scans its words' counts and writes its IDs, ascending, at its offset
// HBM traffic: 8 B per pair + 2 x ceil(N / 8) B of bitmap + 4 B per unique ID.
#include <algorithm>

#include "common.cuh"
#include "internal.cuh"

using namespace b200cd;

#define API extern "C" __attribute__((visibility("default")))

namespace b200cd {

namespace {

constexpr int UQ_THREADS = 256;
constexpr int UQ_WPT = 4;                          // bitmap words per thread
constexpr int UQ_BLOCK_WORDS = UQ_THREADS * UQ_WPT;  // 1024 words = 32768 IDs per block

__global__ void __launch_bounds__(256)
unique_mark_kernel(const uint2* __restrict__ pairs, uint64_t count, uint32_t id_space, uint32_t* __restrict__ bits) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 p = pairs[i];
        if (p.x < id_space) atomicOr(bits + (p.x >> 5), 1u << (p.x & 31));
        if (p.y < id_space) atomicOr(bits + (p.y >> 5), 1u << (p.y & 31));
    }
}

// sums[b] = set bits in block b; the last block to arrive replaces them by exclusive offsets and leaves the
// total in sums[nblocks] and zero in the arrival counter sums[nblocks + 1] (ready for the next call)
__global__ void __launch_bounds__(UQ_THREADS)
unique_count_kernel(const uint32_t* __restrict__ bits, uint64_t words, uint32_t* __restrict__ sums, uint32_t nblocks) {
    __shared__ uint32_t s_w[UQ_THREADS / 32];
    __shared__ bool s_last;
    const uint64_t w0 = (uint64_t)blockIdx.x * UQ_BLOCK_WORDS + (uint64_t)threadIdx.x * UQ_WPT;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < UQ_WPT; ++k)
        if (w0 + k < words) c += __popc(bits[w0 + k]);
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < UQ_THREADS / 32; ++w) t += s_w[w];
        sums[blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(sums + nblocks + 1, 1u) == nblocks - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // exclusive scan of the block counts by this one block (nblocks <= 2^30 / 32768 = 32768)
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += UQ_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? reinterpret_cast<volatile uint32_t*>(sums)[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= (uint32_t)o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t wbase = 0;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_w[w];
        const uint32_t carry = s_carry;
        if (i < nblocks) sums[i] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == UQ_THREADS - 1) s_carry = carry + wbase + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[nblocks] = s_carry;
        sums[nblocks + 1] = 0;
    }
}

__global__ void __launch_bounds__(UQ_THREADS)
unique_emit_kernel(const uint32_t* __restrict__ bits, uint64_t words, const uint32_t* __restrict__ sums,
                   uint32_t* __restrict__ out, uint64_t out_cap) {
    __shared__ uint32_t s_w[UQ_THREADS / 32];
    const uint64_t w0 = (uint64_t)blockIdx.x * UQ_BLOCK_WORDS + (uint64_t)threadIdx.x * UQ_WPT;
    uint32_t word[UQ_WPT];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < UQ_WPT; ++k) {
        word[k] = w0 + k < words ? bits[w0 + k] : 0u;
        c += __popc(word[k]);
    }
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint64_t at = sums[blockIdx.x] + incl - c;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) at += s_w[w];
#pragma unroll
    for (int k = 0; k < UQ_WPT; ++k) {
        uint32_t m = word[k];
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (at < out_cap) out[at] = (uint32_t)((w0 + k) << 5) + (uint32_t)b;
            ++at;
        }
    }
}

}  // namespace

void launch_unique_mark(const uint2* d_pairs, uint64_t count, uint32_t id_space, uint32_t* d_bits, cudaStream_t s) {
    if (!count) return;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((count + 255) / 256, 148ull * 16);
    unique_mark_kernel<<<blocks, 256, 0, s>>>(d_pairs, count, id_space, d_bits);
    count_launch();
}
void launch_unique_count(const uint32_t* d_bits, uint64_t words, uint32_t* d_sums, cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)((words + UQ_BLOCK_WORDS - 1) / UQ_BLOCK_WORDS);
    unique_count_kernel<<<nblocks, UQ_THREADS, 0, s>>>(d_bits, words, d_sums, nblocks);
    count_launch();
}
void launch_unique_emit(const uint32_t* d_bits, uint64_t words, const uint32_t* d_sums, uint32_t* d_out, uint64_t out_cap,
                        cudaStream_t s) {
    const uint32_t nblocks = (uint32_t)((words + UQ_BLOCK_WORDS - 1) / UQ_BLOCK_WORDS);
    unique_emit_kernel<<<nblocks, UQ_THREADS, 0, s>>>(d_bits, words, d_sums, d_out, out_cap);
    count_launch();
}

}  // namespace b200cd

API int b200cd_unique_triangles_device(b200cd_ctx* ctx, const void* d_pairs, uint64_t count, uint32_t id_space,
                                       const void** d_ids_out, uint64_t* count_out) {
    if (!ctx || !count_out || (count && !d_pairs)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *count_out = 0;
    if (d_ids_out) *d_ids_out = nullptr;
    if (id_space == 0 || count == 0) return B200CD_OK;
    if (id_space > (1u << B200CD_MAX_TRIS_LOG2)) return set_error(ctx, B200CD_E_TOOBIG, "id_space too large");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint64_t words = ((uint64_t)id_space + 31) / 32;
    const uint64_t nblocks = (words + 1023) / 1024;
    if (words > ctx->uniq_words) {
        cudaFree(ctx->d_uniq_bits);
        ctx->d_uniq_bits = nullptr;
        ctx->uniq_words = 0;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_uniq_bits), words * sizeof(uint32_t)));
        ctx->uniq_words = words;
    }
    if (nblocks > ctx->uniq_blocks) {
        cudaFree(ctx->d_uniq_sums);
        ctx->d_uniq_sums = nullptr;
        ctx->uniq_blocks = 0;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_uniq_sums), (nblocks + 2) * sizeof(uint32_t)));
        ctx->uniq_blocks = nblocks;
    }
    const uint64_t want = std::min<uint64_t>(2 * count, id_space);  // at most two new IDs per pair
    if (want > ctx->uniq_out_cap) {
        cudaFree(ctx->d_uniq_out);
        ctx->d_uniq_out = nullptr;
        ctx->uniq_out_cap = 0;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_uniq_out), want * sizeof(uint32_t)));
        ctx->uniq_out_cap = want;
    }
    CD_CUDA(ctx, cudaMemsetAsync(ctx->d_uniq_bits, 0, words * sizeof(uint32_t), s));
    CD_CUDA(ctx, cudaMemsetAsync(ctx->d_uniq_sums + nblocks, 0, 2 * sizeof(uint32_t), s));
    launch_unique_mark(static_cast<const uint2*>(d_pairs), count, id_space, ctx->d_uniq_bits, s);
    launch_unique_count(ctx->d_uniq_bits, words, ctx->d_uniq_sums, s);
    launch_unique_emit(ctx->d_uniq_bits, words, ctx->d_uniq_sums, ctx->d_uniq_out, ctx->uniq_out_cap, s);
    CD_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 40, ctx->d_uniq_sums + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));
    CD_CUDA(ctx, cudaGetLastError());
    *count_out = ctx->h_scalars[40];
    if (d_ids_out) *d_ids_out = ctx->d_uniq_out;
    return B200CD_OK;
}

API int b200cd_unique_triangles(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t* ids_out, uint64_t cap, uint64_t* count_out) {
    if (!ctx || !bvh || !count_out || (cap && !ids_out)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bvh->built) return set_error(ctx, B200CD_E_INVALID, "BVH not built");
    const void* d_ids = nullptr;
    int rc = b200cd_unique_triangles_device(ctx, bvh->d_out, bvh->npairs, bvh->id_space ? bvh->id_space : bvh->n, &d_ids, count_out);
    if (rc != B200CD_OK) return rc;
    if (*count_out > cap) return set_error(ctx, B200CD_E_CAPACITY, "id buffer holds " + std::to_string(cap) + ", need " + std::to_string(*count_out));
    DeviceGuard g(ctx->device);
    if (*count_out) CD_CUDA(ctx, cudaMemcpy(ids_out, d_ids, *count_out * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return B200CD_OK;
}
