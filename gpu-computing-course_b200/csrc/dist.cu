// dist.cu — the multi-GPU self-collision step in C++ (b200cd_dist_*, include/b200cd.h): one process per GPU,
// each rank owns ONE Morton range of the triangles, builds and queries only that range and exchanges the thin
// layer of triangles that reach into a higher rank's range ("ghosts"). The reference is single-GPU and has no
// communication at all; its main() (main.cu:47-174) is what one step replaces.
//
// Everything the ranks exchange inside a step moves through CUDA-IPC peer memory over NVLink / NVSwitch, written
// by the kernels that produce it; the ranks are ordered by flag barriers over the same mappings
// (st.release.sys / ld.acquire.sys on per-rank epoch words), not by collectives:
//
//   K1 on my slice of the input + 65536-bin histogram of the keys' top bits
//     -> push my histogram into every rank's comm block                                   [barrier 1]
//   every rank now holds all histograms: the same splitters everywhere, the [source][owner] count matrix,
//     receive offsets and range sizes, locally, in two launches (partition.cu dist_plan)
//     -> ONE host read (the size of my range: the build's grids depend on it)
//   fused partition + exchange: (key, id) stored straight into the owner's sort buffers   [barrier 2]
//   sort + tree over my range; coarse boxes (a cut through the tree) pushed to every rank  [barrier 3]
//   fused ghost selection + send (remote atomics + 256-bit stores) ; local query (traverse + narrow)
//                                                                                          [barrier 4]
//   ghost queries against my tree - their count is read ON THE DEVICE (a fixed grid strides over them)
//   my pairs appended to rank 0's gather buffer (remote atomic reservation + stores), my verdict (buffer
//     overflows) written into every rank's status row                                     [barrier 5]
//   -> ONE host read: the verdicts (all ranks retry together if any buffer overflowed) and, on rank 0, the
//      pair count; rank 0 then sorts the gathered list (asynchronously: valid in stream order).
//
// NCCL is optional here (b200cd_dist_nccl_init, dlopen'ed so the library has no link-time dependency): it carries the
// BVH broadcast of the replicated mode (b200cd_dist_broadcast_bvh) - the data path above needs no collective.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <nccl.h>

#include "common.cuh"
#include "internal.cuh"

using namespace b200cd;

#define API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int DIST_MAX = RS_MAX_SPLIT_P1;  // ranks
constexpr int DIST_K = 256;                // coarse boxes per rank (== ghost_max_k(): the ghost kernel strides by K * 6)
constexpr uint32_t DIST_MAGIC = 0xb200cd15u;

// status bits a rank reports to everyone at the end of a step
enum : uint32_t {
    ST_CAND = 1u,        // candidate list overflowed (grow + retry)
    ST_OUT = 2u,         // pair list overflowed (grow + retry)
    ST_STACK = 4u,       // traversal stack exhausted (fatal: B200CD_E_DEPTH)
    ST_GHOST = 8u,       // more ghosts arrived than the ghost buffer holds (fatal: B200CD_E_CAPACITY)
    ST_GATHER = 16u,     // rank 0's gather buffer is too small (fatal: B200CD_E_CAPACITY)
    ST_TIMEOUT = 32u,    // a barrier timed out (fatal: B200CD_E_PEER)
    ST_RANGE = 64u,      // a Morton range holds more triangles than the rank's capacity (fatal: B200CD_E_CAPACITY)
};

// One per rank, zero-initialised, mapped by every other rank. Plain words: every cross-rank access is an
// explicit st.release.sys / ld.acquire.sys / atomic, or a bulk store ordered by a barrier.
struct CommBlock {
    uint32_t flags[DIST_MAX];          // flags[src] = last barrier epoch rank src has reached
    uint32_t status[DIST_MAX];         // status[src] = rank src's verdict of the current step
    unsigned long long ghost_count;    // ghost records the lower ranks appended to MY ghost buffer this step
    unsigned long long gather_count;   // (rank 0) pairs appended to the gather buffer this step
    uint32_t error;                    // a barrier timed out on this rank
    uint32_t pad[27];
    float boxes[DIST_MAX][DIST_K * 6]; // boxes[src] = rank src's coarse boxes (lo xyz, hi xyz)
    uint32_t hist[DIST_MAX][65536];    // hist[src] = rank src's key histogram
};
static_assert(offsetof(CommBlock, boxes) % 32 == 0 && offsetof(CommBlock, hist) % 32 == 0, "alignment");

struct DistPeers {  // passed to kernels by value
    CommBlock* comm[DIST_MAX];
};

struct DistBlob {  // what a rank publishes (b200cd_dist_export); B200CD_DIST_BLOB_BYTES holds it
    uint32_t magic, rank, world, pad;
    uint64_t cap, ghost_cap, gather_cap;
    cudaIpcMemHandle_t h[4];  // key buffer, id buffer, leaf (ghost) buffer, window (comm block + gather buffer)
    uint64_t off[4];
};
static_assert(sizeof(DistBlob) <= B200CD_DIST_BLOB_BYTES, "blob size");

struct HostResult {  // pinned: what the host reads at the two sync points
    uint32_t totals[DIST_MAX];
    uint32_t status[DIST_MAX];
    unsigned long long counters[8];
    unsigned long long ghost_count, gather_count;
    uint32_t error, pad;
};

enum DistEv { DE_START, DE_HIST, DE_PLAN, DE_EXCHANGE, DE_BUILD, DE_GHOST_SEND, DE_LOCAL, DE_GHOST_QUERY, DE_GATHER, DE_SORT, DE_COUNT };

// ---- NCCL through dlopen (optional)
struct NcclApi {
    void* so = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi* nccl_api(std::string* why) {
    static NcclApi api;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.so) break;
        }
        if (!api.so) {
            err = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
        } else {
#define B200CD_NCCL_SYM(f) api.f = reinterpret_cast<decltype(api.f)>(dlsym(api.so, "nccl" #f))
            B200CD_NCCL_SYM(GetUniqueId);
            B200CD_NCCL_SYM(CommInitRank);
            B200CD_NCCL_SYM(CommDestroy);
            B200CD_NCCL_SYM(Broadcast);
            B200CD_NCCL_SYM(AllReduce);
            B200CD_NCCL_SYM(GetErrorString);
#undef B200CD_NCCL_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllReduce) {
                err = "libnccl.so.2 lacks a required symbol";
                api.so = nullptr;
            }
        }
    }
    if (!api.so) {
        if (why) *why = err;
        return nullptr;
    }
    return &api;
}

}  // namespace

struct b200cd_dist {
    b200cd_ctx* ctx = nullptr;
    uint32_t rank = 0, world = 1, ntris_total = 0;
    uint32_t lo = 0, cnt = 0;        // my slice of the INPUT triangles
    uint32_t cap = 0;                // triangles my range may hold
    uint64_t ghost_cap = 0, gather_cap = 0;
    int shift = 44;
    b200cd_bvh* bvh = nullptr;       // partial BVH over my Morton range
    // peer-visible window: [CommBlock][gather buffer]
    char* d_window = nullptr;
    CommBlock* comm = nullptr;
    uint2* d_gather = nullptr;
    DistPeers peers{};               // every rank's comm block as mapped here (own entry = local)
    uint2* gather0 = nullptr;        // rank 0's gather buffer as mapped here
    void* mapped[DIST_MAX][4] = {};  // cudaIpcOpenMemHandle results, to close
    bool connected = false;
    uint32_t epoch = 0;
    // local scratch
    uint64_t* d_keys_slice = nullptr;
    uint32_t* d_lhist = nullptr;     // my histogram
    uint32_t* d_ghist = nullptr;
    uint32_t* d_part = nullptr;
    DistPlan* d_plan = nullptr;
    float* d_boxes = nullptr;        // my coarse boxes
    uint2* d_sorted = nullptr;       // rank 0: the gathered list, sorted
    uint2* d_sorted_tmp = nullptr;
    uint64_t sorted_cap = 0;
    // rank 0's final sort can run on a side stream (b200cd_dist_set_async_sort), off the other ranks' critical path:
    // they would otherwise wait for it at the first barrier of the next step
    bool async_sort = false, sort_pending = false;
    uint64_t sort_deferred = 0;      // pairs copied into d_sorted whose sort has not been ENQUEUED yet (async mode)
    cudaStream_t sort_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_sorted = nullptr;
    uint32_t* d_sort_hist = nullptr;      // radix scratch of that sort (the context's own may be in use on the main stream)
    uint32_t* d_sort_status = nullptr;
    uint64_t sort_status_words = 0;
    HostResult* h_res = nullptr;
    cudaEvent_t ev[DE_COUNT] = {};
    b200cd_dist_stats stats{};
    bool stats_pending = false;
    uint64_t timeout_ns = 30ull * 1000000000ull;
    // optional NCCL communicator
    ncclComm_t nccl = nullptr;
};

namespace {

// ---------------------------------------------------------------- kernels

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Flag barrier across the ranks: everything this rank enqueued before it (including stores into peer memory) is
// visible to a rank that has passed it. Thread p tells rank p "I have reached `epoch`" and waits until rank p says so.
__global__ void __launch_bounds__(32)
dist_barrier_kernel(DistPeers P, CommBlock* mine, uint32_t rank, uint32_t world, uint32_t epoch, unsigned long long timeout_ns) {
    const uint32_t p = threadIdx.x;
    if (p >= world || p == rank) return;
    if (*reinterpret_cast<volatile uint32_t*>(&mine->error)) return;  // a peer is gone: the step fails anyway, do not wait again
    __threadfence_system();
    st_release_sys(&P.comm[p]->flags[rank], epoch);
    const unsigned long long t0 = globaltimer_ns();
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(&mine->flags[p]) - epoch) < 0) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
            atomicOr(&mine->error, 1u);
            break;
        }
        __nanosleep(64);
    }
}

// my `nvec` 16-byte vectors -> the same place (byte offset `dst_off` of the comm block) on every rank
__global__ void __launch_bounds__(256)
dist_push_kernel(const uint4* __restrict__ src, DistPeers P, uint64_t dst_off, uint32_t nvec, uint32_t world) {
    const uint32_t p = blockIdx.y;
    if (p >= world) return;
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.comm[p]) + dst_off);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// End of a step: (1) my verdict goes into every rank's status row, (2) my pairs are appended to rank 0's gather
// buffer: every block reserves room for its chunk with ONE remote atomic and stores the chunk behind it.
constexpr int GA_THREADS = 256, GA_CHUNK = 4096;
__global__ void __launch_bounds__(GA_THREADS)
dist_gather_kernel(const uint2* __restrict__ pairs, const unsigned long long* __restrict__ counters, uint64_t cand_cap,
                   uint64_t out_cap, DistPeers P, uint2* __restrict__ gather0, uint64_t gather_cap, uint64_t ghost_cap,
                   uint32_t rank, uint32_t world) {
    __shared__ unsigned long long s_base;
    const unsigned long long ncand = max(counters[0], counters[6]), npair = counters[1];  // [6]: the local pass ([0] is reused by the ghost pass)
    CommBlock* mine = P.comm[rank];
    if (blockIdx.x == 0 && threadIdx.x < world) {
        uint32_t st = 0;
        if (ncand > cand_cap) st |= ST_CAND;
        if (npair > out_cap) st |= ST_OUT;
        if (counters[2] & 1ull) st |= ST_STACK;
        if (mine->ghost_count > ghost_cap) st |= ST_GHOST;
        if (mine->error) st |= ST_TIMEOUT;
        P.comm[threadIdx.x]->status[rank] = st;
    }
    const unsigned long long have = min(npair, (unsigned long long)out_cap);
    for (unsigned long long c0 = (unsigned long long)blockIdx.x * GA_CHUNK; c0 < have; c0 += (unsigned long long)gridDim.x * GA_CHUNK) {
        const uint32_t len = (uint32_t)min((unsigned long long)GA_CHUNK, have - c0);
        if (threadIdx.x == 0) s_base = atomicAdd(&P.comm[0]->gather_count, (unsigned long long)len);
        __syncthreads();
        const unsigned long long base = s_base;
        for (uint32_t i = threadIdx.x; i < len; i += GA_THREADS)
            if (base + i < gather_cap) gather0[base + i] = pairs[c0 + i];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host helpers

int set_dist_error(b200cd_dist* d, int code, const std::string& msg) { return set_error(d ? d->ctx : nullptr, code, msg); }

// Async mode: the host work of launching the sort (10 launches) is put off until the NEXT step has enqueued its first
// phase (or until somebody asks for the list), so that rank 0 reaches the next step's first barrier with the others.
int launch_deferred_sort(b200cd_dist* d) {
    if (!d->sort_deferred) return B200CD_OK;
    b200cd_ctx* ctx = d->ctx;
    const uint64_t total = d->sort_deferred;
    d->sort_deferred = 0;
    CD_CUDA(ctx, cudaStreamWaitEvent(d->sort_stream, d->ev_fork, 0));
    uint2* const promised = d->d_sorted;  // the address b200cd_dist_step has already handed to the caller
    int rc = sort_pairs_impl(ctx, &d->d_sorted, &d->d_sorted_tmp, total, id_bits_for(d->ntris_total), &d->d_sort_hist,
                             &d->d_sort_status, &d->sort_status_words, d->sort_stream);
    if (rc != B200CD_OK) return rc;
    if (d->d_sorted != promised) {  // (a pair sort runs 2 x ceil(id_bits / 8) passes - an even number - and lands where it started)
        CD_CUDA(ctx, cudaMemcpyAsync(promised, d->d_sorted, total * sizeof(uint2), cudaMemcpyDeviceToDevice, d->sort_stream));
        std::swap(d->d_sorted, d->d_sorted_tmp);
    }
    CD_CUDA(ctx, cudaEventRecord(d->ev_sorted, d->sort_stream));
    d->sort_pending = true;
    return B200CD_OK;
}

void barrier(b200cd_dist* d, cudaStream_t s) {
    if (d->world < 2) return;
    ++d->epoch;
    dist_barrier_kernel<<<1, 32, 0, s>>>(d->peers, d->comm, d->rank, d->world, d->epoch, d->timeout_ns);
    count_launch();
    trace_mark("barrier", s);
}

void push(b200cd_dist* d, const void* src, uint64_t dst_off, uint64_t bytes, cudaStream_t s) {
    const uint32_t nvec = (uint32_t)(bytes / 16);
    const uint32_t bx = std::min<uint32_t>((nvec + 255) / 256, 32u);
    dist_push_kernel<<<dim3(bx, d->world), 256, 0, s>>>(static_cast<const uint4*>(src), d->peers, dst_off, nvec, d->world);
    count_launch();
    trace_mark("push", s);
}

void close_mappings(b200cd_dist* d) {
    for (uint32_t r = 0; r < DIST_MAX; ++r)
        for (int i = 0; i < 4; ++i)
            if (d->mapped[r][i]) {
                cudaIpcCloseMemHandle(d->mapped[r][i]);
                d->mapped[r][i] = nullptr;
            }
}

float dev_ms(b200cd_dist* d, int a, int b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, d->ev[a], d->ev[b]) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

}  // namespace

// ------------------------------------------------------------------ create / export / connect / destroy

API int b200cd_dist_create(b200cd_ctx* ctx, uint32_t rank, uint32_t world, uint32_t ntris_total, double slack,
                           uint64_t pair_capacity, b200cd_dist** out) {
    if (!ctx || !out || world == 0 || world > (uint32_t)DIST_MAX || rank >= world) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    *out = nullptr;
    if (ntris_total > (1u << B200CD_MAX_TRIS_LOG2)) return set_error(ctx, B200CD_E_TOOBIG, "mesh too large");
    if (!(slack >= 1.0)) slack = 1.5;
    DeviceGuard g(ctx->device);
    b200cd_dist* d = new (std::nothrow) b200cd_dist;
    if (!d) return set_error(ctx, B200CD_E_NOMEM, "host allocation failed");
    d->ctx = ctx;
    d->rank = rank;
    d->world = world;
    d->ntris_total = ntris_total;
    d->lo = (uint32_t)((uint64_t)rank * ntris_total / world);
    d->cnt = (uint32_t)((uint64_t)(rank + 1) * ntris_total / world) - d->lo;
    d->cap = (uint32_t)std::min<uint64_t>(ntris_total, (uint64_t)((double)ntris_total / world * slack) + 65536);
    d->ghost_cap = std::max<uint64_t>(ntris_total / world / 2, 65536);
    d->gather_cap = pair_capacity ? pair_capacity : (uint64_t)ntris_total + 65536;
    if (const char* e = getenv("B200CD_BARRIER_TIMEOUT_MS")) {
        const long ms = atol(e);
        if (ms > 0) d->timeout_ns = (uint64_t)ms * 1000000ull;
    }
    int rc = alloc_bvh(ctx, d->cap, 0, true, &d->bvh, d->ghost_cap, world, /*ghost_out*/ false);
    auto A = [&](cudaError_t e) {
        if (rc == B200CD_OK && e != cudaSuccess) {
            cudaGetLastError();
            rc = set_error(ctx, e == cudaErrorMemoryAllocation ? B200CD_E_NOMEM : B200CD_E_CUDA, cudaGetErrorString(e));
        }
    };
    if (rc == B200CD_OK) {
        d->bvh->n = 0;
        const uint64_t window = sizeof(CommBlock) + sizeof(uint2) * d->gather_cap;
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_window), window));
        if (rc == B200CD_OK) A(cudaMemset(d->d_window, 0, sizeof(CommBlock)));
        d->comm = reinterpret_cast<CommBlock*>(d->d_window);
        d->d_gather = reinterpret_cast<uint2*>(d->d_window + sizeof(CommBlock));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_keys_slice), sizeof(uint64_t) * std::max<uint32_t>(d->cnt, 1u)));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_lhist), sizeof(uint32_t) * 65536));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_ghist), sizeof(uint32_t) * 65536));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_part), sizeof(uint32_t) * (DIST_MAX * DIST_HIST_BLOCKS + 1024)));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_plan), sizeof(DistPlan)));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_boxes), sizeof(float) * DIST_K * 6));
        A(cudaMallocHost(reinterpret_cast<void**>(&d->h_res), sizeof(HostResult)));
        A(cudaMalloc(reinterpret_cast<void**>(&d->d_sort_hist), radix_hist_words(8) * sizeof(uint32_t)));
        A(cudaStreamCreateWithFlags(&d->sort_stream, cudaStreamNonBlocking));
        A(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&d->ev_sorted, cudaEventDisableTiming));
        for (int i = 0; i < DE_COUNT && rc == B200CD_OK; ++i) A(cudaEventCreate(&d->ev[i]));
    }
    if (rc == B200CD_OK) {
        memset(d->h_res, 0, sizeof(HostResult));
        for (uint32_t r = 0; r < (uint32_t)DIST_MAX; ++r) d->peers.comm[r] = d->comm;  // until connected: everything is me
        d->gather0 = d->d_gather;
        d->connected = world == 1;
        // the peer table of the partition / ghost kernels (single rank: own buffers)
        PeerTable t;
        memset(&t, 0, sizeof t);
        t.keys[rank] = d->bvh->d_keys[0];
        t.ids[rank] = d->bvh->d_ids[0];
        t.ghosts[rank] = d->bvh->d_leaves + d->bvh->cap;
        t.ghost_count[rank] = &d->comm->ghost_count;
        t.ghost_cap = d->ghost_cap;
        A(cudaMemcpy(d->bvh->d_peers, &t, sizeof t, cudaMemcpyHostToDevice));
        A(cudaDeviceSynchronize());
    }
    if (rc != B200CD_OK) {
        b200cd_dist_destroy(d);
        return rc;
    }
    *out = d;
    return B200CD_OK;
}

API int b200cd_dist_export(b200cd_dist* d, uint8_t* blob_out) {
    if (!d || !blob_out) return set_dist_error(d, B200CD_E_INVALID, "NULL argument");
    b200cd_ctx* ctx = d->ctx;
    DeviceGuard g(ctx->device);
    DistBlob b;
    memset(&b, 0, sizeof b);
    b.magic = DIST_MAGIC;
    b.rank = d->rank;
    b.world = d->world;
    b.cap = d->cap;
    b.ghost_cap = d->ghost_cap;
    b.gather_cap = d->gather_cap;
    void* ptrs[4] = {d->bvh->d_keys[0], d->bvh->d_ids[0], d->bvh->d_leaves, d->d_window};
    for (int i = 0; i < 4; ++i) {
        CD_CUDA(ctx, cudaIpcGetMemHandle(&b.h[i], ptrs[i]));
        if (alloc_base_offset(ptrs[i], &b.off[i]) != 0) return set_error(ctx, B200CD_E_CUDA, "address range query failed");
    }
    b.off[2] += sizeof(LeafRec) * (uint64_t)d->bvh->cap;  // peers address my GHOST records
    memset(blob_out, 0, B200CD_DIST_BLOB_BYTES);
    memcpy(blob_out, &b, sizeof b);
    return B200CD_OK;
}

API int b200cd_dist_connect(b200cd_dist* d, const uint8_t* blobs) {
    if (!d || !blobs) return set_dist_error(d, B200CD_E_INVALID, "NULL argument");
    b200cd_ctx* ctx = d->ctx;
    if (d->connected && d->world > 1) return set_error(ctx, B200CD_E_INVALID, "already connected");
    DeviceGuard g(ctx->device);
    PeerTable t;
    memset(&t, 0, sizeof t);
    t.ghost_cap = d->ghost_cap;
    for (uint32_t r = 0; r < d->world; ++r) {
        DistBlob b;
        memcpy(&b, blobs + (size_t)r * B200CD_DIST_BLOB_BYTES, sizeof b);
        if (b.magic != DIST_MAGIC || b.rank != r || b.world != d->world) {
            close_mappings(d);
            return set_error(ctx, B200CD_E_INVALID, "blob " + std::to_string(r) + " is not rank " + std::to_string(r) + "'s export");
        }
        if (b.cap != d->cap || b.ghost_cap != d->ghost_cap || b.gather_cap != d->gather_cap) {
            close_mappings(d);
            return set_error(ctx, B200CD_E_INVALID, "ranks were created with different sizes");
        }
        char* base[4];
        if (r == d->rank) {
            t.keys[r] = d->bvh->d_keys[0];
            t.ids[r] = d->bvh->d_ids[0];
            t.ghosts[r] = d->bvh->d_leaves + d->bvh->cap;
            d->peers.comm[r] = d->comm;
            if (r == 0) d->gather0 = d->d_gather;
        } else {
            for (int i = 0; i < 4; ++i) {
                void* p = nullptr;
                cudaError_t e = cudaIpcOpenMemHandle(&p, b.h[i], cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    close_mappings(d);
                    return set_error(ctx, B200CD_E_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
                }
                d->mapped[r][i] = p;
                base[i] = static_cast<char*>(p) + b.off[i];
            }
            t.keys[r] = reinterpret_cast<uint64_t*>(base[0]);
            t.ids[r] = reinterpret_cast<uint32_t*>(base[1]);
            t.ghosts[r] = reinterpret_cast<LeafRec*>(base[2]);
            d->peers.comm[r] = reinterpret_cast<CommBlock*>(base[3]);
            if (r == 0) d->gather0 = reinterpret_cast<uint2*>(base[3] + sizeof(CommBlock));
        }
        t.ghost_count[r] = &d->peers.comm[r]->ghost_count;
    }
    CD_CUDA(ctx, cudaMemcpy(d->bvh->d_peers, &t, sizeof t, cudaMemcpyHostToDevice));
    CD_CUDA(ctx, cudaDeviceSynchronize());
    d->connected = true;
    return B200CD_OK;
}

API int b200cd_dist_destroy(b200cd_dist* d) {
    if (!d) return B200CD_OK;
    DeviceGuard g(d->ctx->device);
    launch_deferred_sort(d);
    cudaStreamSynchronize(d->ctx->stream);
    close_mappings(d);
    if (d->nccl) {
        if (NcclApi* api = nccl_api(nullptr)) api->CommDestroy(d->nccl);
    }
    if (d->bvh) b200cd_bvh_destroy(d->bvh);
    cudaFree(d->d_window);
    cudaFree(d->d_keys_slice);
    cudaFree(d->d_lhist);
    cudaFree(d->d_ghist);
    cudaFree(d->d_part);
    cudaFree(d->d_plan);
    cudaFree(d->d_boxes);
    if (d->sort_stream) {
        cudaStreamSynchronize(d->sort_stream);
        cudaStreamDestroy(d->sort_stream);
    }
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    if (d->ev_sorted) cudaEventDestroy(d->ev_sorted);
    cudaFree(d->d_sort_hist);
    cudaFree(d->d_sort_status);
    cudaFree(d->d_sorted);
    cudaFree(d->d_sorted_tmp);
    if (d->h_res) cudaFreeHost(d->h_res);
    for (int i = 0; i < DE_COUNT; ++i)
        if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    delete d;
    return B200CD_OK;
}

// A barrier across the ranks on the context's stream (e.g. "everybody's slice of the next frame has landed").
API int b200cd_dist_barrier(b200cd_dist* d) {
    if (!d) return B200CD_E_INVALID;
    if (!d->connected) return set_dist_error(d, B200CD_E_INVALID, "b200cd_dist_connect has not been called");
    DeviceGuard g(d->ctx->device);
    barrier(d, d->ctx->stream);
    CD_CUDA(d->ctx, cudaGetLastError());
    return B200CD_OK;
}

// Pipelined frames: with on != 0 rank 0's final sort of the gathered list runs on a side stream, so the rank (and with it
// every other rank, at the next step's first barrier) goes straight on to the next frame. The list b200cd_dist_step
// returned is then valid once b200cd_dist_wait_sorted has been called (it makes the context's stream wait for that sort,
// no host blocking) - call it before reading the list; the next b200cd_dist_step does it implicitly.
API int b200cd_dist_set_async_sort(b200cd_dist* d, int on) {
    if (!d) return B200CD_E_INVALID;
    DeviceGuard g(d->ctx->device);
    if (!on) {
        int rc = launch_deferred_sort(d);
        if (rc != B200CD_OK) return rc;
    }
    if (!on && d->sort_pending) {
        CD_CUDA(d->ctx, cudaStreamWaitEvent(d->ctx->stream, d->ev_sorted, 0));
        d->sort_pending = false;
    }
    d->async_sort = on != 0;
    return B200CD_OK;
}

API int b200cd_dist_wait_sorted(b200cd_dist* d) {
    if (!d) return B200CD_E_INVALID;
    DeviceGuard g(d->ctx->device);
    int rc = launch_deferred_sort(d);
    if (rc != B200CD_OK) return rc;
    if (!d->sort_pending) return B200CD_OK;
    CD_CUDA(d->ctx, cudaStreamWaitEvent(d->ctx->stream, d->ev_sorted, 0));
    d->sort_pending = false;
    return B200CD_OK;
}

API int b200cd_dist_bvh(b200cd_dist* d, b200cd_bvh** out) {
    if (!d || !out) return B200CD_E_INVALID;
    *out = d->bvh;
    return B200CD_OK;
}

// ------------------------------------------------------------------ the step

API int b200cd_dist_step(b200cd_dist* d, const b200cd_mesh* mesh, const b200cd_params* params, const void** d_pairs_out,
                         uint64_t* count_out) {
    if (!d || !mesh || !count_out) return set_dist_error(d, B200CD_E_INVALID, "NULL argument");
    b200cd_ctx* ctx = d->ctx;
    *count_out = 0;
    if (d_pairs_out) *d_pairs_out = nullptr;
    if (!d->connected) return set_error(ctx, B200CD_E_INVALID, "b200cd_dist_connect has not been called");
    if (mesh->ntris != d->ntris_total) return set_error(ctx, B200CD_E_INVALID, "mesh size differs from the one given to b200cd_dist_create");
    int rc = check_params(ctx, params);
    if (rc != B200CD_OK) return rc;
    if (params->auto_box) return set_error(ctx, B200CD_E_INVALID, "auto_box is not available for a partitioned build: pass the box");
    if (mesh->pending) return set_error(ctx, B200CD_E_INVALID, "mesh has an asynchronous upload in flight: call b200cd_mesh_wait first");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    b200cd_bvh* b = d->bvh;
    const uint32_t W = d->world, R = d->rank;
    d->shift = params->key_bits == 63 ? 44 : 14;  // top 16 of the 60 / 30 bits in-box keys use
    // buffers the whole step can run in without a host decision (grow-only; a retry enlarges them).
    // B200CD_DIST_TINY_BUFFERS=1 (tests): start far too small, so that the collective retry path runs.
    static const bool tiny = getenv("B200CD_DIST_TINY_BUFFERS") != nullptr;
    rc = grow(ctx, &b->d_cand, &b->cand_cap, std::max<uint64_t>(b->cand_cap, tiny ? 2048 : 4ull * d->cap + 4096));
    if (rc == B200CD_OK) rc = grow(ctx, &b->d_out, &b->out_cap, std::max<uint64_t>(b->out_cap, tiny ? 256 : (uint64_t)d->cap / 2 + 4096));
    if (rc == B200CD_OK) rc = ensure_entry_lists(ctx, b, (uint64_t)d->cap + 2 * B200CD_QUERY_BLOCK * 4);
    if (rc != B200CD_OK) return rc;
    HostResult* H = d->h_res;
    const uint32_t mask_higher = W >= 32 ? 0u : (~0u << (R + 1)) & ((1u << W) - 1u);

    for (int attempt = 0; attempt < 4; ++attempt) {
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_START], s));
        trace_mark("step_begin", s);
        // ---- reset what the peers will append to (ordered before their appends by barrier 1)
        CD_CUDA(ctx, cudaMemsetAsync(&d->comm->ghost_count, 0, 2 * sizeof(unsigned long long), s));  // ghost_count, gather_count
        CD_CUDA(ctx, cudaMemsetAsync(d->d_lhist, 0, sizeof(uint32_t) * 65536, s));
        // ---- K1 on my slice of the input + histogram of the keys' top bits; everyone gets everyone's histogram
        launch_morton(mesh->d_verts, mesh->d_idx, d->lo, d->cnt, *params, nullptr, d->d_keys_slice, s);
        launch_key_hist16(d->d_keys_slice, d->cnt, d->shift, d->d_lhist, ctx->sm_count, s);
        push(d, d->d_lhist, offsetof(CommBlock, hist) + sizeof(uint32_t) * 65536ull * R, sizeof(uint32_t) * 65536ull, s);
        barrier(d, s);
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_HIST], s));
        rc = launch_deferred_sort(d);  // (rank 0, async mode) the previous frame's sort, now that this frame is under way
        if (rc != B200CD_OK) return rc;
        // ---- the plan: splitters, count matrix, receive offsets, range sizes (identical on every rank)
        launch_dist_plan(&d->comm->hist[0][0], (int)W, (int)R, d->shift, d->d_ghist, d->d_part, d->d_plan, s);
        CD_CUDA(ctx, cudaMemcpyAsync(H->totals, d->d_plan->totals, sizeof(uint32_t) * DIST_MAX, cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_PLAN], s));
        CD_CUDA(ctx, cudaStreamSynchronize(s));  // host read 1: the build's launch grids depend on the size of my range
        uint32_t biggest = 0;
        for (uint32_t r = 0; r < W; ++r) biggest = std::max(biggest, H->totals[r]);
        if (biggest > d->cap)  // every rank sees every range's size: all of them return here together
            return set_error(ctx, B200CD_E_CAPACITY, "a Morton range holds " + std::to_string(biggest) + " triangles, capacity " +
                                                         std::to_string(d->cap) + ": raise slack (very uneven mesh)");
        const uint32_t nlocal = H->totals[R];
        // ---- fused partition + exchange: every (key, id) goes straight into its owner's sort buffers
        radix_partition_to_peers(d->d_keys_slice, d->lo, d->cnt, d->d_plan->splitters, (int)W - 1, b->d_peers, d->d_plan->recv_off,
                                 b->d_hist, b->d_tile_status, s);
        barrier(d, s);
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_EXCHANGE], s));
        // ---- sort + tree over my range; my coarse boxes to everyone
        b->n = nlocal;
        b->nverts = mesh->nverts;
        rc = run_build(ctx, b, mesh, params, /*keys_given*/ true);
        if (rc != B200CD_OK) return rc;
        launch_chunk_boxes(b->d_pairs, b->d_root_box, nlocal, DIST_K, b->d_cut_scratch, d->d_boxes, s);
        push(d, d->d_boxes, offsetof(CommBlock, boxes) + sizeof(float) * DIST_K * 6ull * R, sizeof(float) * DIST_K * 6ull, s);
        barrier(d, s);
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_BUILD], s));
        // ---- ghosts for the higher ranks leave while the local query runs
        if (mask_higher && nlocal)
            launch_ghosts_to_peers(b->d_leaves, nlocal, &d->comm->boxes[0][0], W, DIST_K, mask_higher, b->d_peers,
                                   reinterpret_cast<float*>(b->d_cut_scratch), b->d_block_boxes, s, b->d_ghost_list, b->ghost_list_cap, ctx->sm_count);
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_GHOST_SEND], s));
        CD_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, 8 * sizeof(unsigned long long), s));
        if (nlocal >= 2) {
            const uint32_t chunk = (nlocal + B200CD_QUERY_BLOCK - 1) / B200CD_QUERY_BLOCK * B200CD_QUERY_BLOCK;
            CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q0], s));
            launch_broad(b->d_pairs, b->d_leaves, b->d_root_box, nlocal, 0, 1, chunk, chunk, /*foreign*/ 0, 0u, b->d_entries,
                         b->d_entry_count, b->d_cand, b->cand_cap, b->d_counters, s, nullptr, ctx->sm_count, !b->unshared_verts,
                         b->qvalid ? b->d_qpairs : nullptr, b->d_qframe);
            CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q1], s));
            launch_narrow(b->d_leaves, b->d_cand, b->cand_cap, b->d_out, b->out_cap, b->d_counters, ctx->sm_count, s, b->unshared_verts);
            CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q2], s));
        }
        // the local candidate count: kept for the verdict (the ghost pass reuses counter 0)
        CD_CUDA(ctx, cudaMemcpyAsync(b->d_counters + 6, b->d_counters, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_LOCAL], s));
        barrier(d, s);  // every lower rank's ghost records (and their count) have landed
        // ---- ghost queries against my tree; their pairs go behind the local ones (counter 1 runs on)
        if (R > 0 && nlocal) {
            CD_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, sizeof(unsigned long long), s));
            launch_broad(b->d_pairs, b->d_leaves, b->d_root_box, nlocal, 0, 1, B200CD_QUERY_BLOCK, (uint32_t)d->ghost_cap, /*foreign*/ 1,
                         b->cap, b->d_entries, b->d_entry_count, b->d_cand, b->cand_cap, b->d_counters, s, &d->comm->ghost_count,
                         ctx->sm_count);
            launch_narrow(b->d_leaves, b->d_cand, b->cand_cap, b->d_out, b->out_cap, b->d_counters, ctx->sm_count, s, b->unshared_verts);
        }
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_GHOST_QUERY], s));
        // ---- verdicts to everyone, pairs to rank 0
        dist_gather_kernel<<<std::max(ctx->sm_count, 1), GA_THREADS, 0, s>>>(b->d_out, b->d_counters, b->cand_cap, b->out_cap, d->peers,
                                                                          d->gather0, d->gather_cap, d->ghost_cap, R, W);
        count_launch();
        trace_mark("gather (verdicts + pairs to rank 0)", s);
        barrier(d, s);
        CD_CUDA(ctx, cudaMemcpyAsync(H->status, d->comm->status, sizeof(uint32_t) * DIST_MAX, cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaMemcpyAsync(H->counters, b->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaMemcpyAsync(&H->ghost_count, &d->comm->ghost_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaMemcpyAsync(&H->error, &d->comm->error, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_GATHER], s));
        CD_CUDA(ctx, cudaStreamSynchronize(s));  // host read 2: everybody's verdict (+ the pair count on rank 0)
        CD_CUDA(ctx, cudaGetLastError());
        if (H->error) return set_error(ctx, B200CD_E_PEER, "a barrier between the ranks timed out (a peer failed or is far behind)");
        uint32_t any = 0;
        for (uint32_t r = 0; r < W; ++r) any |= H->status[r];
        if (any & ST_TIMEOUT) return set_error(ctx, B200CD_E_PEER, "a barrier timed out on another rank");
        if (any & ST_STACK) return set_error(ctx, B200CD_E_DEPTH, "traversal stack of " + std::to_string(B200CD_MAX_STACK) + " entries exhausted");
        if (any & ST_GHOST) return set_error(ctx, B200CD_E_CAPACITY, "more ghosts arrived on a rank than its ghost buffer holds");
        const uint64_t ncand_ghost = H->counters[0], ncand_local = H->counters[6], npair = H->counters[1];
        if (any & (ST_CAND | ST_OUT)) {  // somebody's list overflowed: that rank grows it, everybody runs the step again
            const uint64_t ncand = std::max(ncand_ghost, ncand_local);
            if (ncand > b->cand_cap) rc = grow(ctx, &b->d_cand, &b->cand_cap, ncand + ncand / 8);
            if (rc == B200CD_OK && npair > b->out_cap) rc = grow(ctx, &b->d_out, &b->out_cap, npair + npair / 8);
            if (rc != B200CD_OK) return rc;
            d->stats.retries++;
            continue;
        }
        // ---- done: statistics, and on rank 0 the sort of the gathered list (asynchronous)
        if (nlocal >= 2) {  // the LOCAL query's stage times (b200cd_get_stats), like b200cd_self_collide's
            ctx->stats.ms_traverse = ev_ms(ctx, EV_Q0, EV_Q1);
            ctx->stats.ms_narrow = ev_ms(ctx, EV_Q1, EV_Q2);
            ctx->stats.ms_pair_sort = 0.f;
            ctx->stats.ms_query = ev_ms(ctx, EV_Q0, EV_Q2);
        }
        ctx->stats.candidates = ncand_local + ncand_ghost;
        ctx->stats.pairs = npair;
        ctx->stats.nodes_visited = H->counters[3];
        ctx->stats.warp_steps = H->counters[4];
        ctx->stats.start_entries = H->counters[5];
        b->npairs = npair;
        d->stats.local_triangles = nlocal;
        d->stats.ghosts = H->ghost_count;
        d->stats.local_pairs = npair;
        d->stats.candidates = ncand_local + ncand_ghost;
        d->stats.total_pairs = 0;
        uint64_t total = 0;
        if (R == 0) {
            total = H->gather_count;
            if (total > d->gather_cap)
                return set_error(ctx, B200CD_E_CAPACITY, "gathered pair list holds " + std::to_string(total) + " pairs, capacity " +
                                                             std::to_string(d->gather_cap) + " (b200cd_dist_create pair_capacity)");
            if (total > d->sorted_cap) {
                cudaFree(d->d_sorted);
                cudaFree(d->d_sorted_tmp);
                d->d_sorted = d->d_sorted_tmp = nullptr;
                d->sorted_cap = 0;
                const uint64_t want = total + total / 4 + 1024;
                CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d->d_sorted), want * sizeof(uint2)));
                CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d->d_sorted_tmp), want * sizeof(uint2)));
                d->sorted_cap = want;
            }
            if (d->sort_pending) {  // the previous step's sort (side stream) still owns d_sorted / d_sorted_tmp
                CD_CUDA(ctx, cudaStreamWaitEvent(s, d->ev_sorted, 0));
                d->sort_pending = false;
            }
            if (total) CD_CUDA(ctx, cudaMemcpyAsync(d->d_sorted, d->d_gather, total * sizeof(uint2), cudaMemcpyDeviceToDevice, s));
            if (d->async_sort) {  // fork: the sort will run beside whatever this rank enqueues next on its main stream
                CD_CUDA(ctx, cudaEventRecord(d->ev_fork, s));
                d->sort_deferred = total;  // enqueued by the next step behind its first phase, or by b200cd_dist_wait_sorted
            } else {
                rc = sort_pairs_impl(ctx, &d->d_sorted, &d->d_sorted_tmp, total, id_bits_for(d->ntris_total), &d->d_sort_hist,
                                     &d->d_sort_status, &d->sort_status_words, s);
                if (rc != B200CD_OK) return rc;
            }
            d->stats.total_pairs = total;
            if (d_pairs_out) *d_pairs_out = d->d_sorted;
        }
        trace_mark("pair sort (rank 0)", s);
        CD_CUDA(ctx, cudaEventRecord(d->ev[DE_SORT], s));
        CD_CUDA(ctx, cudaGetLastError());
        d->stats_pending = true;
        *count_out = total;
        return B200CD_OK;
    }
    return set_error(ctx, B200CD_E_CUDA, "distributed step did not converge after growing its buffers");
}

API int b200cd_dist_get_stats(b200cd_dist* d, b200cd_dist_stats* out) {
    if (!d || !out) return B200CD_E_INVALID;
    DeviceGuard g(d->ctx->device);
    if (d->stats_pending) {
        CD_CUDA(d->ctx, cudaStreamSynchronize(d->ctx->stream));
        d->stats.ms_keys_hist = dev_ms(d, DE_START, DE_HIST);
        d->stats.ms_plan = dev_ms(d, DE_HIST, DE_PLAN);
        d->stats.ms_exchange = dev_ms(d, DE_PLAN, DE_EXCHANGE);
        d->stats.ms_build = dev_ms(d, DE_EXCHANGE, DE_BUILD);
        d->stats.ms_ghost_send = dev_ms(d, DE_BUILD, DE_GHOST_SEND);
        d->stats.ms_local_query = dev_ms(d, DE_GHOST_SEND, DE_LOCAL);
        d->stats.ms_ghost_query = dev_ms(d, DE_LOCAL, DE_GHOST_QUERY);
        d->stats.ms_gather = dev_ms(d, DE_GHOST_QUERY, DE_GATHER);
        d->stats.ms_sort = dev_ms(d, DE_GATHER, DE_SORT);
        d->stats.ms_step = dev_ms(d, DE_START, DE_SORT);
        d->stats_pending = false;
    }
    d->stats.rank = d->rank;
    d->stats.world = d->world;
    *out = d->stats;
    return B200CD_OK;
}

// ------------------------------------------------------------------ NCCL (optional): BVH broadcast of the replicated mode

API int b200cd_nccl_unique_id(uint8_t* id128) {
    if (!id128) return B200CD_E_INVALID;
    NcclApi* api = nccl_api(nullptr);
    if (!api) return B200CD_E_NODEVICE;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return B200CD_E_CUDA;
    memcpy(id128, &id, 128);
    return B200CD_OK;
}

API int b200cd_dist_nccl_init(b200cd_dist* d, const uint8_t* id128) {
    if (!d || !id128) return set_dist_error(d, B200CD_E_INVALID, "NULL argument");
    std::string why;
    NcclApi* api = nccl_api(&why);
    if (!api) return set_error(d->ctx, B200CD_E_NODEVICE, why);
    if (d->nccl) return B200CD_OK;
    DeviceGuard g(d->ctx->device);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    const ncclResult_t r = api->CommInitRank(&d->nccl, (int)d->world, id, (int)d->rank);
    if (r != ncclSuccess) {
        d->nccl = nullptr;
        return set_error(d->ctx, B200CD_E_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "failed"));
    }
    return B200CD_OK;
}

// Replicated mode, "receives the BVH, broadcast via NCCL over NVLink": the three device blobs of a built BVH
// (traversal nodes, leaf records, sorted ids) travel from rank `root` to every other rank, which passes a BVH
// made by b200cd_bvh_alloc_like for the same triangle count. Enqueued on the context's stream.
API int b200cd_dist_broadcast_bvh(b200cd_dist* d, b200cd_bvh* bvh, uint32_t root) {
    if (!d || !bvh || root >= d->world) return set_dist_error(d, B200CD_E_INVALID, "bad argument");
    b200cd_ctx* ctx = d->ctx;
    if (!d->nccl) return set_error(ctx, B200CD_E_INVALID, "b200cd_dist_nccl_init has not been called");
    NcclApi* api = nccl_api(nullptr);
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint32_t n = bvh->n;
    struct { void* p; uint64_t bytes; } blobs[4] = {
        {bvh->d_pairs, n > 1 ? sizeof(NodePair) * (uint64_t)(n - 1) : 0},
        {bvh->d_leaves, sizeof(LeafRec) * (uint64_t)n},
        {bvh->d_ids[bvh->cur], 4ull * n},
        {bvh->d_root_box, 8 * sizeof(float)},
    };
    for (auto& bl : blobs) {
        if (!bl.bytes) continue;
        const ncclResult_t r = api->Broadcast(bl.p, bl.p, bl.bytes, ncclChar, (int)root, d->nccl, s);
        if (r != ncclSuccess) return set_error(ctx, B200CD_E_CUDA, std::string("ncclBroadcast: ") + (api->GetErrorString ? api->GetErrorString(r) : "failed"));
    }
    bvh->built = true;
    bvh->id_space = n;
    if (d->rank != root) bvh->qvalid = false;  // the quantised nodes do not travel: a received BVH is walked on the exact ones
    return B200CD_OK;
}
