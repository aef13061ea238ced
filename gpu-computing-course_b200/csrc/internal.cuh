// internal.cuh — host-side helpers of api.cu that dist.cu (the multi-GPU step) and unique.cu reuse.
// Nothing here is part of the C ABI.
#pragma once
#include "common.cuh"

namespace b200cd {

enum Ev { EV_B0, EV_B1, EV_B2, EV_B3, EV_B4, EV_Q0, EV_Q1, EV_Q2, EV_Q3, EV_U0, EV_U1, EV_D0, EV_D1, EV_COUNT };

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
int dev_alloc(b200cd_ctx* ctx, T** p, uint64_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
    return B200CD_OK;
}

float ev_ms(b200cd_ctx* ctx, int a, int b);
int id_bits_for(uint32_t n);
// ghost_cap / max_peers: room for a partitioned build's ghost records, block boxes and peer table;
// ghost_out: also the per-peer outgoing ghost lists of the NCCL send/recv exchange (max_peers x ghost_cap records)
int alloc_bvh(b200cd_ctx* ctx, uint32_t n, uint32_t nverts, bool with_sort, b200cd_bvh** out, uint64_t ghost_cap = 0,
              uint32_t max_peers = 0, bool ghost_out = true);
void free_bvh_buffers(b200cd_bvh* b);
int check_params(b200cd_ctx* ctx, const b200cd_params* p);
int run_build(b200cd_ctx* ctx, b200cd_bvh* b, const b200cd_mesh* m, const b200cd_params* p, bool keys_given = false);
cudaError_t mark_consumed(const b200cd_mesh* cm, cudaStream_t s);
int grow(b200cd_ctx* ctx, uint2** buf, uint64_t* cap, uint64_t want);
int sort_pairs_impl(b200cd_ctx* ctx, uint2** d_pairs, uint2** d_tmp, uint64_t count, int id_bits, uint32_t** d_hist,
                    uint32_t** d_status, uint64_t* status_words, cudaStream_t s);
// entry lists for up to `nquery` query slots (grow-only)
int ensure_entry_lists(b200cd_ctx* ctx, b200cd_bvh* b, uint64_t nquery);

// base address of the allocation a device pointer belongs to (cudaIpcOpenMemHandle maps whole allocations)
int alloc_base_offset(void* ptr, uint64_t* offset_out);

}  // namespace b200cd
