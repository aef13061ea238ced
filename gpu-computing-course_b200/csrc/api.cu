// api.cu — the C ABI of libb200cd.so (include/b200cd.h): contexts, meshes, the
// build pipeline K1..K4 and the query pipeline K5..K6, error handling, stage timers.
//
// Host-side mirror of the reference driver main() (reference
// CollisionDetection/main.cu:47-174): load -> alloc/H2D -> build stages -> query ->
// D2H. Differences: handles instead of raw cudaMalloc'ed structs, status codes
// instead of HANDLE_ERROR/exit (common/book.h:21-31), one stream with no
// per-kernel cudaEventSynchronize (main.cu:93,100,109,143), buffers that grow
// instead of the fixed 500-pair list (main.cu:81).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <cuda.h>

#include "common.cuh"
#include "internal.cuh"

using namespace b200cd;

// base address of the allocation a device pointer belongs to. The driver entry point is fetched
// through the runtime, so the library does not link libcuda (it must load on machines without a driver).
static int cuMemGetAddressRange_shim(void** base, size_t* size, void* ptr) {
    typedef CUresult (*fn_t)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return 1;
    CUdeviceptr b = 0;
    size_t s = 0;
    CUresult r = reinterpret_cast<fn_t>(fn)(&b, &s, (CUdeviceptr)ptr);
    *base = (void*)b;
    *size = s;
    return r == CUDA_SUCCESS ? 0 : 1;
}

#define API extern "C" __attribute__((visibility("default")))

namespace b200cd {
int alloc_base_offset(void* ptr, uint64_t* offset_out) {
    void* base = nullptr;
    size_t size = 0;
    if (cuMemGetAddressRange_shim(&base, &size, ptr) != 0) return 1;
    *offset_out = (uint64_t)((char*)ptr - (char*)base);
    return 0;
}
}  // namespace b200cd

namespace b200cd {
unsigned long long g_kernel_launches = 0;

namespace {
struct TraceState {
    int on = -1;  // -1: environment not read yet
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    std::vector<cudaEvent_t> pool;
} g_trace;
}  // namespace

void trace_mark(const char* name, cudaStream_t s) {
    if (g_trace.on == 0) return;
    if (g_trace.on < 0) {
        g_trace.on = getenv("B200CD_TRACE") ? 1 : 0;
        if (!g_trace.on) return;
    }
    if (g_trace.marks.size() >= 200000) return;
    cudaEvent_t e = nullptr;
    if (!g_trace.pool.empty()) {
        e = g_trace.pool.back();
        g_trace.pool.pop_back();
    } else if (cudaEventCreate(&e) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    cudaEventRecord(e, s);
    g_trace.marks.emplace_back(name, e);
}
}  // namespace b200cd

extern "C" __attribute__((visibility("default"))) void b200cd_trace_enable(int on) { g_trace.on = on ? 1 : 0; }

// debug aid (not a reference entry point): write and clear the kernel timeline collected since the last dump
extern "C" __attribute__((visibility("default"))) int b200cd_trace_dump(b200cd_ctx* ctx, const char* path) {
    if (!ctx || !path) return B200CD_E_INVALID;
    cudaDeviceSynchronize();
    FILE* f = fopen(path, "a");
    if (!f) return B200CD_E_IO;
    for (size_t i = 0; i < g_trace.marks.size(); ++i) {
        float ms = 0.f;
        if (i > 0 && cudaEventElapsedTime(&ms, g_trace.marks[i - 1].second, g_trace.marks[i].second) != cudaSuccess) {
            cudaGetLastError();
            ms = -1.f;
        }
        fprintf(f, "%s,%.4f\n", g_trace.marks[i].first, ms);
    }
    fclose(f);
    for (auto& m : g_trace.marks) g_trace.pool.push_back(m.second);
    g_trace.marks.clear();
    return B200CD_OK;
}

namespace b200cd {  // internal helpers shared with dist.cu (declared in internal.cuh)

float ev_ms(b200cd_ctx* ctx, int a, int b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

void free_bvh_buffers(b200cd_bvh* b) {
    for (int i = 0; i < 2; ++i) {
        cudaFree(b->d_keys[i]);
        cudaFree(b->d_ids[i]);
    }
    cudaFree(b->d_hist);
    cudaFree(b->d_tile_status);
    cudaFree(b->d_fix);
    if (b->h_fix) cudaFreeHost(b->h_fix);
    if (b->ev_fix) cudaEventDestroy(b->ev_fix);
    for (int i = 0; i < 4; ++i)
        if (b->ev_sort[i]) cudaEventDestroy(b->ev_sort[i]);
    cudaFree(b->d_flags);
    cudaFree(b->d_build_scratch);
    cudaFree(b->d_pairs);
    cudaFree(b->d_qpairs);
    cudaFree(b->d_qframe);
    cudaFree(b->d_leaves);
    cudaFree(b->d_recs);
    cudaFree(b->d_block_boxes);
    cudaFree(b->d_ghost_out);
    cudaFree(b->d_cut_scratch);
    cudaFree(b->d_peers);
    cudaFree(b->d_ghost_in_count);
    cudaFree(b->d_ghost_list);
    cudaFree(b->d_root_box);
    cudaFree(b->d_cand);
    cudaFree(b->d_entries);
    cudaFree(b->d_entry_count);
    cudaFree(b->d_out);
    cudaFree(b->d_out_tmp);
    cudaFree(b->d_counters);
    if (b->h_counters) cudaFreeHost(b->h_counters);
}

int id_bits_for(uint32_t n) {
    int b = 1;
    while (b < 32 && (n - 1) >> b) ++b;
    return n <= 1 ? 1 : b;
}

int alloc_bvh(b200cd_ctx* ctx, uint32_t n, uint32_t nverts, bool with_sort, b200cd_bvh** out, uint64_t ghost_cap,
              uint32_t max_peers, bool ghost_out) {
    b200cd_bvh* b = new (std::nothrow) b200cd_bvh;
    if (!b) return set_error(ctx, B200CD_E_NOMEM, "host allocation failed");
    b->ctx = ctx;
    b->n = n;
    b->cap = n;
    b->ghost_cap = ghost_cap;
    b->nverts = nverts;
    int rc = B200CD_OK;
    auto A = [&](int r) { if (rc == B200CD_OK) rc = r; };
    if (with_sort) {
        A(dev_alloc(ctx, &b->d_keys[0], n));
        A(dev_alloc(ctx, &b->d_keys[1], n));
        A(dev_alloc(ctx, &b->d_ids[1], n));
        A(dev_alloc(ctx, &b->d_flags, n));
        if (rc == B200CD_OK && cudaMalloc(&b->d_build_scratch, build_tree_scratch_bytes(n)) != cudaSuccess) {
            cudaGetLastError();
            rc = set_error(ctx, B200CD_E_NOMEM, "cudaMalloc failed");
        }
    }
    A(dev_alloc(ctx, &b->d_ids[0], n));
    // radix scratch is sized for the larger of the key sort and an n-pair result sort; grown on demand
    b->tile_status_words = radix_tile_status_words(std::max<uint32_t>(n, 1u), 8);
    A(dev_alloc(ctx, &b->d_hist, radix_hist_words(8)));
    A(dev_alloc(ctx, &b->d_tile_status, b->tile_status_words));
    if (with_sort) {
        A(dev_alloc(ctx, &b->d_fix, 4));
        if (rc == B200CD_OK && (cudaMallocHost(reinterpret_cast<void**>(&b->h_fix), 4 * sizeof(uint32_t)) != cudaSuccess ||
                                cudaEventCreateWithFlags(&b->ev_fix, cudaEventDisableTiming) != cudaSuccess))
            rc = set_error(ctx, B200CD_E_NOMEM, "cudaMallocHost failed");
        for (int i = 0; i < 4 && rc == B200CD_OK; ++i)
            if (cudaEventCreate(&b->ev_sort[i]) != cudaSuccess) rc = set_error(ctx, B200CD_E_NOMEM, "cudaEventCreate failed");
    }
    A(dev_alloc(ctx, &b->d_pairs, n));
    A(dev_alloc(ctx, &b->d_qframe, 8));  // (d_qpairs: on the first build that wants them, run_build)
    A(dev_alloc(ctx, &b->d_leaves, (uint64_t)n + ghost_cap));  // ghost records of a partitioned build live after the local leaves
    if (max_peers) {
        b->ghost_out_cap = ghost_cap;
        b->max_peers = ghost_out ? max_peers : 0;
        if (ghost_out) A(dev_alloc(ctx, &b->d_ghost_out, (uint64_t)max_peers * ghost_cap));  // staging of the NCCL send/recv exchange only
        A(dev_alloc(ctx, &b->d_cut_scratch, (uint64_t)ghost_max_k() * 6 + 1));
        A(dev_alloc(ctx, &b->d_block_boxes, ((uint64_t)n / 256 + 1) * 8));
        A(dev_alloc(ctx, &b->d_ghost_list, ghost_list_words(n, max_peers)));
        b->ghost_list_cap = ghost_list_items(n, max_peers);
        A(dev_alloc(ctx, &b->d_peers, 1));
        A(dev_alloc(ctx, &b->d_ghost_in_count, 2));
    }
    A(dev_alloc(ctx, &b->d_root_box, 8));
    A(dev_alloc(ctx, &b->d_counters, 8));
    if (rc == B200CD_OK && cudaMallocHost(reinterpret_cast<void**>(&b->h_counters), 8 * sizeof(unsigned long long)) != cudaSuccess)
        rc = set_error(ctx, B200CD_E_NOMEM, "cudaMallocHost failed");
    if (rc != B200CD_OK) {
        free_bvh_buffers(b);
        delete b;
        return rc;
    }
    *out = b;
    return B200CD_OK;
}

// The last kernel that reads the mesh's device arrays has been enqueued on `s`: an asynchronous
// upload into the same mesh (b200cd_mesh_update_async, on the copy stream) must not start before it.
cudaError_t mark_consumed(const b200cd_mesh* cm, cudaStream_t s) {
    b200cd_mesh* m = const_cast<b200cd_mesh*>(cm);
    if (!m->ev_consumed) return cudaSuccess;  // never uploaded asynchronously: nothing can race
    m->consumed_valid = true;
    return cudaEventRecord(m->ev_consumed, s);
}

bool leaf_records_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200CD_RECS");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// B200CD_SORT=full: always run every radix pass (tuning / A-B knob, read once)
bool sort_hybrid_disabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200CD_SORT");
        v = (e && e[0] == 'f') ? 1 : 0;
    }
    return v == 1;
}

int check_params(b200cd_ctx* ctx, const b200cd_params* p) {
    if (!p) return set_error(ctx, B200CD_E_INVALID, "params is NULL");
    if (p->key_bits != 63 && p->key_bits != 30) return set_error(ctx, B200CD_E_INVALID, "key_bits must be 63 or 30");
    if (!p->auto_box)
        for (int a = 0; a < 3; ++a)
            if (!(p->morton_extent[a] > 0.0)) return set_error(ctx, B200CD_E_INVALID, "morton_extent must be > 0");
    return B200CD_OK;
}

// keys_given: d_keys[0] / d_ids[0] already hold (key, triangle id) of the n triangles of this tree
// (partitioned multi-GPU build); otherwise K1 computes the keys of the whole mesh and ids are 0..n-1.
int run_build(b200cd_ctx* ctx, b200cd_bvh* b, const b200cd_mesh* m, const b200cd_params* p, bool keys_given) {
    cudaStream_t s = ctx->stream;
    const uint32_t n = b->n;
    if (m->pending) return set_error(ctx, B200CD_E_INVALID, "mesh has an asynchronous upload in flight: call b200cd_mesh_wait first");
    b->params = *p;
    const QFrame qframe = make_qframe(*p);  // the traversal's 15-bit grid: over the previous root box, else over the Morton box
    const bool have_root = b->built && b->root_valid;
    b->built = false;
    b->unshared_verts = 2ull * m->nverts >= 3ull * m->ntris;
    b->qvalid = broad_uses_quantised_nodes(!b->unshared_verts);
    if (b->qvalid && !b->d_qpairs && b->cap > 1 &&  // (sized like d_pairs: a partitioned build's n changes from step to step)
        cudaMalloc(reinterpret_cast<void**>(&b->d_qpairs), sizeof(QNodePair) * (size_t)b->cap) != cudaSuccess) {
        cudaGetLastError();  // no room for the second node array: walk the exact nodes, not an error
        b->d_qpairs = nullptr;
    }
    if (!b->d_qpairs) b->qvalid = false;
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B0], s));
    trace_mark("build_begin", s);
    int npass = 0;
    // ---- the sort's plan first (host logic only): K1 can then count the digits of the passes that will run
    RadixPass passes[8];
    int high = 0;
    if (n) {
        // K2 plan: radix_digit_bits()-wide digits over the significant key bits (60 of 63: morton.h:15 masks 21 bits
        // per axis but the 2^20 scale leaves bit 20 clear for in-box meshes; we still sort all
        // 63 so out-of-box meshes order exactly like the host sort)
        const int total_bits = (p->key_bits == 30) ? 30 : 63;
        const int db = radix_digit_bits();
        for (int sh = 0; sh < total_bits; sh += db) passes[npass++] = {sh, std::min(db, total_bits - sh)};
        // Hybrid sort for 63-bit keys: radix passes over the top b->sort_high digits only, the low bits are put in
        // order by a per-run fix-up (radix_sort.cu). The previous build's run statistics steer sort_high.
        if (npass == 8 && b->d_fix && !sort_hybrid_disabled()) {
            const cudaError_t landed = b->fix_pending ? cudaEventQuery(b->ev_fix) : cudaErrorNotReady;
            if (landed != cudaSuccess) cudaGetLastError();  // "not ready" is an answer, not a failure: do not leave it behind
            if (b->fix_pending && landed == cudaSuccess) {
                b->fix_pending = false;
                const uint32_t overflow = b->h_fix[0] & 1u, longest = b->h_fix[1], in_runs = b->h_fix[2];
                float pass_ms = 0.f, fix_ms = 0.f;  // (the copy of h_fix landed behind them on the same stream)
                if (cudaEventElapsedTime(&pass_ms, b->ev_sort[0], b->ev_sort[1]) != cudaSuccess ||
                    cudaEventElapsedTime(&fix_ms, b->ev_sort[2], b->ev_sort[3]) != cudaSuccess) {
                    cudaGetLastError();
                    pass_ms = fix_ms = 0.f;
                }
                // (bit 1 of h_fix[0]: a key reached above the digit window - the fallback sorted that build; the window
                // only ever moves up, so a mesh whose keys hover around a power of two does not fall back every frame)
                if ((int)b->h_fix[3] > b->sort_top) b->sort_top = (int)b->h_fix[3];
                const bool slow_fix = pass_ms > 0.f && fix_ms > 1.3f * pass_ms;
                b->sort_slow_streak = slow_fix ? b->sort_slow_streak + 1 : 0;
                if (overflow || longest > 24) {           // prefix too short for this mesh: sort more digits, for good
                    b->sort_high = overflow ? 8 : std::min(8, b->sort_high + 1);
                    b->sort_locked = true;
                    b->sort_trial = false;
                } else if (b->sort_trial && pass_ms > 0.f) {
                    // verdict on the extra digit tried last time: keep it only if fix-up + one more pass beat the old fix-up
                    b->sort_trial = false;
                    if (fix_ms + pass_ms >= b->sort_fix_before) {
                        --b->sort_high;                   // no gain (e.g. runs of truly equal keys, which no digit separates)
                        b->sort_time_frozen = true;
                    }
                } else if (!b->sort_time_frozen && b->sort_high < 8 &&
                           (b->sort_slow_streak >= 3 || (pass_ms > 0.f && fix_ms > 3.f * pass_ms))) {
                    // The fix-up orders every run with ONE thread (serial insertion, divergent): cheap while runs are rare and
                    // short, slow once most items sit in runs of 10+ (measured on the 2^25-triangle half of the two sheets:
                    // 1.31 ms behind 4 passes against 0.09 ms behind 5 passes of 0.29 ms each). Its cost depends on the run
                    // length distribution, so it is MEASURED (events around the previous build's fix-up and last pass): when
                    // it exceeds one more radix pass plus the fix-up's floor (two reads of the keys, ~0.3 pass) the next build
                    // sorts one digit more - after three slow builds in a row, or at once when it cost more than three passes
                    // (one mildly slow measurement can be another stream's copy or a time slice) - and the build after that
                    // checks that it paid off.
                    ++b->sort_high;
                    b->sort_locked = true;
                    b->sort_trial = true;
                    b->sort_fix_before = fix_ms;
                    b->sort_slow_streak = 0;
                } else if (!b->sort_locked && b->sort_high > 4 && longest <= 3 && (uint64_t)in_runs * 1024 < n) {
                    --b->sort_high;                       // one digit less multiplies the occupancy of a cell by up to 256
                }
            }
            high = b->sort_high < 8 ? b->sort_high : 0;
        }
    }
    // B200CD_FUSED_HIST=0: the sort reads the keys once more for its histograms (A/B knob, read once)
    static int fused = -1;
    if (fused < 0) {
        const char* e = getenv("B200CD_FUSED_HIST");
        fused = (e && e[0] == '0') ? 0 : 1;
    }
    const bool hist_in_k1 = fused && n && !keys_given;
    RadixHistPlan hplan{};
    if (n && !keys_given) {
        // K1
        uint32_t* d_bbox = nullptr;
        if (p->auto_box) {
            d_bbox = ctx->d_scalars;
            launch_bbox(m->d_verts, m->nverts, d_bbox, ctx->sm_count, s);
        }
        // face-ordered leaf records for the tree build (64 B per triangle, allocated on first use; B200CD_RECS=0 turns
        // them off - tuning / A-B knob). Without them the build gathers indices and vertices itself.
        // Worth it when vertices are (mostly) unshared - a triangle soup, V = 3N: -0.21 ms in the tree build for
        // +0.11 ms in K1 at 16 M. On a mesh (V ~ N/2) the vertex gathers hit L2 anyway and the records only add traffic.
        const bool want_recs = leaf_records_enabled() && b->unshared_verts;
        if (!want_recs && b->d_recs) { cudaFree(b->d_recs); b->d_recs = nullptr; }
        if (!b->d_recs && want_recs && cudaMalloc(reinterpret_cast<void**>(&b->d_recs), sizeof(LeafRec) * (size_t)n) != cudaSuccess) {
            cudaGetLastError();
            b->d_recs = nullptr;  // not enough memory: the gather path needs none
        }
        if (hist_in_k1) {
            radix_hist_plan(passes, npass, /*values*/ true, high, b->d_fix != nullptr, b->sort_top, &hplan);
            CD_CUDA(ctx, cudaMemsetAsync(b->d_hist, 0, sizeof(uint32_t) * radix_hist_words(npass), s));
        }
        launch_morton(m->d_verts, m->d_idx, 0, n, *p, d_bbox, b->d_keys[0], s, b->d_recs, hist_in_k1 ? &hplan : nullptr,
                      hist_in_k1 ? b->d_hist : nullptr, ctx->sm_count);
    }
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B1], s));
    if (n) {
        b->cur = radix_sort(b->d_keys, b->d_ids, n, passes, npass, /*iota*/ !keys_given, b->d_hist, b->d_tile_status,
                            b->tile_status_words, ctx->sm_count, s, high, b->d_fix, b->sort_top, high ? b->ev_sort : nullptr, hist_in_k1);
        if (high) {
            CD_CUDA(ctx, cudaMemcpyAsync(b->h_fix, b->d_fix, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CD_CUDA(ctx, cudaEventRecord(b->ev_fix, s));
            b->fix_pending = true;
            npass = high;
        }
    }
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B2], s));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B3], s));  // (K3 is fused into K4: ms_hierarchy stays ~0)
    launch_build_tree(m->d_verts, m->d_idx, b->d_ids[b->cur], b->d_keys[b->cur], n, b->d_flags, b->d_pairs, b->d_leaves,
                      b->d_root_box, b->d_build_scratch, s, keys_given ? nullptr : b->d_recs, b->d_block_boxes,
                      b->qvalid ? b->d_qpairs : nullptr, b->d_qframe, &qframe, have_root);  // K3+K4
    b->root_valid = n > 0;
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B4], s));
    CD_CUDA(ctx, mark_consumed(m, s));  // nothing after this point reads the mesh
    CD_CUDA(ctx, cudaGetLastError());
    b->id_space = keys_given ? m->ntris : n;
    b->npairs = 0;
    ctx->stats.sort_passes = (uint32_t)npass;
    ctx->stats.ntris = n;
    ctx->stats.nverts = m->nverts;
    ctx->stats.ms_build = -1.f;  // resolved lazily by b200cd_get_stats (build is asynchronous)
    b->built = true;
    return B200CD_OK;
}

}  // namespace b200cd

// ------------------------------------------------------------------ context

API int b200cd_abi_version(void) { return B200CD_ABI_VERSION; }

API const char* b200cd_strerror(int status) {
    switch (status) {
        case B200CD_OK: return "ok";
        case B200CD_E_INVALID: return "invalid argument";
        case B200CD_E_CUDA: return "CUDA runtime error";
        case B200CD_E_NOMEM: return "out of memory";
        case B200CD_E_IO: return "file could not be read";
        case B200CD_E_PARSE: return "OBJ line not in the accepted dialect";
        case B200CD_E_CAPACITY: return "pair buffer too small";
        case B200CD_E_DEPTH: return "BVH traversal stack exhausted";
        case B200CD_E_NODEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
        case B200CD_E_TOOBIG: return "mesh too large";
        case B200CD_E_PEER: return "multi-GPU step: a barrier between the ranks timed out";
        default: return "unknown status";
    }
}

API const char* b200cd_last_error(const b200cd_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

API int b200cd_create(int device, b200cd_ctx** out) {
    if (!out) return B200CD_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return B200CD_E_NODEVICE;
    }
    if (device < 0 || device >= ndev) return B200CD_E_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B200CD_E_CUDA;
    if (prop.major != 10) return B200CD_E_NODEVICE;  // the only code in this library is sm_100a SASS
    b200cd_ctx* ctx = new (std::nothrow) b200cd_ctx;
    if (!ctx) return B200CD_E_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    DeviceGuard g(device);
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < EV_COUNT; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    ok = ok && cudaMalloc(reinterpret_cast<void**>(&ctx->d_scalars), 64 * sizeof(uint32_t)) == cudaSuccess;
    ok = ok && cudaMallocHost(reinterpret_cast<void**>(&ctx->h_scalars), 64 * sizeof(uint32_t)) == cudaSuccess;
    if (!ok) {
        b200cd_destroy(ctx);
        return B200CD_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return B200CD_OK;
}

API int b200cd_destroy(b200cd_ctx* ctx) {
    if (!ctx) return B200CD_OK;
    DeviceGuard g(ctx->device);
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->own_stream);
        cudaStreamDestroy(ctx->own_stream);
    }
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    for (int i = 0; i < 4; ++i) {
        if (ctx->push_stream[i]) {
            cudaStreamSynchronize(ctx->push_stream[i]);
            cudaStreamDestroy(ctx->push_stream[i]);
            cudaEventDestroy(ctx->push_ev[i]);
        }
    }
    for (int i = 0; i < EV_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    cudaFree(ctx->d_scalars);
    cudaFree(ctx->d_sort_tmp);
    cudaFree(ctx->d_sort_hist);
    cudaFree(ctx->d_sort_status);
    cudaFree(ctx->d_uniq_bits);
    cudaFree(ctx->d_uniq_sums);
    cudaFree(ctx->d_uniq_out);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    delete ctx;
    return B200CD_OK;
}

API int b200cd_set_stream(b200cd_ctx* ctx, void* cuda_stream) {
    if (!ctx) return B200CD_E_INVALID;
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cuda_stream == B200CD_PRIVATE_STREAM) ? ctx->own_stream : static_cast<cudaStream_t>(cuda_stream);
    return B200CD_OK;
}

API int b200cd_synchronize(b200cd_ctx* ctx) {
    if (!ctx) return B200CD_E_INVALID;
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200CD_OK;
}

API int b200cd_get_stats(const b200cd_ctx* cctx, b200cd_stats* out) {
    if (!cctx || !out) return B200CD_E_INVALID;
    b200cd_ctx* ctx = const_cast<b200cd_ctx*>(cctx);
    DeviceGuard g(ctx->device);
    if (ctx->stats.ms_build < 0.f) {
        CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stats.ms_morton = ev_ms(ctx, EV_B0, EV_B1);
        ctx->stats.ms_sort = ev_ms(ctx, EV_B1, EV_B2);
        ctx->stats.ms_hierarchy = ev_ms(ctx, EV_B2, EV_B3);
        ctx->stats.ms_refit = ev_ms(ctx, EV_B3, EV_B4);
        ctx->stats.ms_build = ev_ms(ctx, EV_B0, EV_B4);
    }
    ctx->stats.kernel_launches = g_kernel_launches;
    *out = ctx->stats;
    return B200CD_OK;
}

API void b200cd_default_params(b200cd_params* p) {
    if (!p) return;
    // reference morton.h:45,51,57
    p->morton_origin[0] = 0.004501;  p->morton_extent[0] = 3.08;
    p->morton_origin[1] = -0.476622; p->morton_extent[1] = 0.76;
    p->morton_origin[2] = -0.381965; p->morton_extent[2] = 2.36;
    p->key_bits = 63;
    p->auto_box = 0;
    p->pair_capacity_hint = 0;
}

API int b200cd_device_count(int* count_out) {
    if (!count_out) return B200CD_E_INVALID;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count_out = n;
    return n > 0 ? B200CD_OK : B200CD_E_NODEVICE;
}

API int b200cd_copy_to_host(b200cd_ctx* ctx, void* dst, const void* d_src, uint64_t bytes) {
    if (!ctx || (bytes && (!dst || !d_src))) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bytes) return B200CD_OK;
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200CD_OK;
}

API int b200cd_host_alloc(void** out, uint64_t bytes) {
    if (!out) return B200CD_E_INVALID;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? B200CD_OK : B200CD_E_NOMEM;
}
API int b200cd_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? B200CD_OK : B200CD_E_CUDA; }

// ------------------------------------------------------------------ mesh

namespace {
int new_mesh(b200cd_ctx* ctx, uint32_t nverts, uint32_t ntris, b200cd_mesh** out) {
    if (nverts > (1u << 30) || ntris > (1u << 30)) return set_error(ctx, B200CD_E_TOOBIG, "more than 2^30 vertices or triangles");
    b200cd_mesh* m = new (std::nothrow) b200cd_mesh;
    if (!m) return set_error(ctx, B200CD_E_NOMEM, "host allocation failed");
    m->ctx = ctx;
    m->nverts = nverts;
    m->ntris = ntris;
    // + 16 elements of padding: a multi-GPU caller all-gathers equal chunks of ceil(n / ranks) in place
    int rc = dev_alloc(ctx, &m->d_verts, (uint64_t)nverts + 16);
    if (rc == B200CD_OK) rc = dev_alloc(ctx, &m->d_idx, 3ull * ((uint64_t)ntris + 16));
    if (rc != B200CD_OK) {
        cudaFree(m->d_verts);
        cudaFree(m->d_idx);
        delete m;
        return rc;
    }
    *out = m;
    return B200CD_OK;
}

// keep_stage: the float3 landing buffer stays allocated for the next frame (b200cd_mesh_update);
// one-shot uploads (mesh_from_arrays / load_obj) give it back.
int upload(b200cd_ctx* ctx, b200cd_mesh* m, const float* xyz, const uint32_t* idx, bool on_device, bool keep_stage) {
    cudaStream_t s = ctx->stream;
    if (m->pending) return set_error(ctx, B200CD_E_INVALID, "mesh has an asynchronous upload in flight: call b200cd_mesh_wait first");
    if (xyz && m->nverts && !on_device && !m->d_stage)
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&m->d_stage), 12ull * m->nverts));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_U0], s));
    if (xyz && m->nverts) {
        const float* src = xyz;
        if (!on_device) {
            CD_CUDA(ctx, cudaMemcpyAsync(m->d_stage, xyz, 12ull * m->nverts, cudaMemcpyHostToDevice, s));
            src = m->d_stage;
        }
        launch_expand_verts(src, m->d_verts, m->nverts, s);
    }
    if (idx && m->ntris)
        CD_CUDA(ctx, cudaMemcpyAsync(m->d_idx, idx, 12ull * m->ntris, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    if (idx) {  // every index must name an existing vertex (the reference only prints a warning, load_obj.h:77-79)
        launch_check_idx(m->d_idx, m->ntris, m->nverts, ctx->d_scalars + 32, ctx->sm_count, s);
        CD_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 32, ctx->d_scalars + 32, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_U1], s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));  // the caller may free / reuse its buffers on return
    CD_CUDA(ctx, cudaGetLastError());
    ctx->stats.ms_upload = ev_ms(ctx, EV_U0, EV_U1);
    if (!keep_stage && m->d_stage) {
        cudaFree(m->d_stage);
        m->d_stage = nullptr;
    }
    if (idx && ctx->h_scalars[32]) return set_error(ctx, B200CD_E_INVALID, "triangle references a vertex index >= nverts");
    return B200CD_OK;
}
}  // namespace

API int b200cd_mesh_from_arrays(b200cd_ctx* ctx, const float* xyz, uint32_t nverts, const uint32_t* tri_idx,
                                uint32_t ntris, b200cd_mesh** out) {
    if (!ctx || !out || (nverts && !xyz) || (ntris && !tri_idx)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    b200cd_mesh* m = nullptr;
    int rc = new_mesh(ctx, nverts, ntris, &m);
    if (rc != B200CD_OK) return rc;
    rc = upload(ctx, m, xyz, tri_idx, false, false);
    if (rc != B200CD_OK) {
        b200cd_mesh_destroy(m);
        return rc;
    }
    *out = m;
    return B200CD_OK;
}

API int b200cd_mesh_from_device(b200cd_ctx* ctx, const void* d_xyz, uint32_t nverts, const void* d_tri_idx,
                                uint32_t ntris, b200cd_mesh** out) {
    if (!ctx || !out || (nverts && !d_xyz) || (ntris && !d_tri_idx)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    b200cd_mesh* m = nullptr;
    int rc = new_mesh(ctx, nverts, ntris, &m);
    if (rc != B200CD_OK) return rc;
    rc = upload(ctx, m, static_cast<const float*>(d_xyz), static_cast<const uint32_t*>(d_tri_idx), true, false);
    if (rc != B200CD_OK) {
        b200cd_mesh_destroy(m);
        return rc;
    }
    *out = m;
    return B200CD_OK;
}

API int b200cd_mesh_update(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, const uint32_t* tri_idx, int on_device) {
    if (!ctx || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    return upload(ctx, mesh, xyz, tri_idx, on_device != 0, true);
}

// Double-buffered frames: the upload of frame k+1 (H2D, float3 -> float4 expansion, index check) runs on the
// context's copy stream while the work stream builds and queries frame k from ANOTHER mesh object.
// Ordering: the copy waits for the last kernel that read this mesh (ev_consumed); users of the mesh call
// b200cd_mesh_wait, which blocks the host until the upload has landed and reports the index check.
API int b200cd_mesh_update_async(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, const uint32_t* tri_idx) {
    if (!ctx || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (mesh->pending) return set_error(ctx, B200CD_E_INVALID, "mesh already has an asynchronous upload in flight");
    DeviceGuard g(ctx->device);
    if (!ctx->copy_stream) CD_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!mesh->ev_ready) {
        CD_CUDA(ctx, cudaEventCreateWithFlags(&mesh->ev_ready, cudaEventDisableTiming));
        CD_CUDA(ctx, cudaEventCreateWithFlags(&mesh->ev_consumed, cudaEventDisableTiming));
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&mesh->d_async_flag), sizeof(uint32_t)));
        CD_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&mesh->h_async_flag), sizeof(uint32_t)));
        // work enqueued on the work stream before this call may still be reading the mesh
        CD_CUDA(ctx, mark_consumed(mesh, ctx->stream));
    }
    if (xyz && mesh->nverts && !mesh->d_stage)
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&mesh->d_stage), 12ull * mesh->nverts));
    cudaStream_t c = ctx->copy_stream;
    if (mesh->consumed_valid) CD_CUDA(ctx, cudaStreamWaitEvent(c, mesh->ev_consumed, 0));
    if (xyz && mesh->nverts) {
        CD_CUDA(ctx, cudaMemcpyAsync(mesh->d_stage, xyz, 12ull * mesh->nverts, cudaMemcpyHostToDevice, c));
        launch_expand_verts(mesh->d_stage, mesh->d_verts, mesh->nverts, c);
    }
    *mesh->h_async_flag = 0;
    if (tri_idx && mesh->ntris) {
        CD_CUDA(ctx, cudaMemcpyAsync(mesh->d_idx, tri_idx, 12ull * mesh->ntris, cudaMemcpyHostToDevice, c));
        launch_check_idx(mesh->d_idx, mesh->ntris, mesh->nverts, mesh->d_async_flag, ctx->sm_count, c);
        CD_CUDA(ctx, cudaMemcpyAsync(mesh->h_async_flag, mesh->d_async_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c));
    }
    CD_CUDA(ctx, cudaEventRecord(mesh->ev_ready, c));
    CD_CUDA(ctx, cudaGetLastError());
    mesh->pending = true;
    return B200CD_OK;
}

API int b200cd_mesh_wait(b200cd_ctx* ctx, b200cd_mesh* mesh) {
    if (!ctx || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!mesh->pending) return B200CD_OK;
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaEventSynchronize(mesh->ev_ready));  // the caller's host buffers are free again from here on
    mesh->pending = false;
    if (*mesh->h_async_flag) return set_error(ctx, B200CD_E_INVALID, "triangle references a vertex index >= nverts");
    return B200CD_OK;
}

// Partial upload for multi-GPU callers: every rank sends 1/ranks of the mesh over PCIe and the ranks
// all-gather the device buffers over NVLink (b200cd_mesh_device_buffers), instead of every rank
// pulling the whole mesh through the host.
API int b200cd_mesh_update_slice(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, uint32_t first_vert, uint32_t nverts,
                                 const uint32_t* tri_idx, uint32_t first_tri, uint32_t ntris) {
    if (!ctx || !mesh || (nverts && !xyz) || (ntris && !tri_idx)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if ((uint64_t)first_vert + nverts > mesh->nverts || (uint64_t)first_tri + ntris > mesh->ntris)
        return set_error(ctx, B200CD_E_INVALID, "slice out of range");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    if (nverts && !mesh->d_stage) CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&mesh->d_stage), 12ull * mesh->nverts));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_U0], s));
    if (nverts) {
        CD_CUDA(ctx, cudaMemcpyAsync(mesh->d_stage, xyz, 12ull * nverts, cudaMemcpyHostToDevice, s));
        launch_expand_verts(mesh->d_stage, mesh->d_verts + first_vert, nverts, s);
    }
    if (ntris) {
        CD_CUDA(ctx, cudaMemcpyAsync(mesh->d_idx + 3ull * first_tri, tri_idx, 12ull * ntris, cudaMemcpyHostToDevice, s));
        launch_check_idx(mesh->d_idx + 3ull * first_tri, ntris, mesh->nverts, ctx->d_scalars + 32, ctx->sm_count, s);
        CD_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars + 32, ctx->d_scalars + 32, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_U1], s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));
    CD_CUDA(ctx, cudaGetLastError());
    ctx->stats.ms_upload = ev_ms(ctx, EV_U0, EV_U1);
    if (ntris && ctx->h_scalars[32]) return set_error(ctx, B200CD_E_INVALID, "triangle references a vertex index >= nverts");
    return B200CD_OK;
}

// device views of the mesh: float4[nverts + 16] vertices (xyz, w = 0) and uint32[3 * (ntris + 16)] indices
API int b200cd_mesh_device_buffers(b200cd_ctx* ctx, b200cd_mesh* mesh, void** d_verts4, void** d_idx) {
    if (!ctx || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (d_verts4) *d_verts4 = mesh->d_verts;
    if (d_idx) *d_idx = mesh->d_idx;
    return B200CD_OK;
}

API int b200cd_mesh_info(const b200cd_mesh* mesh, uint32_t* nverts, uint32_t* ntris) {
    if (!mesh) return B200CD_E_INVALID;
    if (nverts) *nverts = mesh->nverts;
    if (ntris) *ntris = mesh->ntris;
    return B200CD_OK;
}

API int b200cd_mesh_download(b200cd_ctx* ctx, const b200cd_mesh* mesh, float* xyz, uint32_t* tri_idx) {
    if (!ctx || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (xyz && mesh->nverts) {
        std::vector<float4> tmp(mesh->nverts);
        CD_CUDA(ctx, cudaMemcpy(tmp.data(), mesh->d_verts, sizeof(float4) * mesh->nverts, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < mesh->nverts; ++i) {
            xyz[3 * (size_t)i] = tmp[i].x; xyz[3 * (size_t)i + 1] = tmp[i].y; xyz[3 * (size_t)i + 2] = tmp[i].z;
        }
    }
    if (tri_idx && mesh->ntris)
        CD_CUDA(ctx, cudaMemcpy(tri_idx, mesh->d_idx, 12ull * mesh->ntris, cudaMemcpyDeviceToHost));
    return B200CD_OK;
}

API int b200cd_mesh_destroy(b200cd_mesh* mesh) {
    if (!mesh) return B200CD_OK;
    DeviceGuard g(mesh->ctx->device);
    cudaStreamSynchronize(mesh->ctx->stream);
    if (mesh->ev_ready) {
        cudaEventSynchronize(mesh->ev_ready);
        cudaEventDestroy(mesh->ev_ready);
        cudaEventDestroy(mesh->ev_consumed);
        cudaFree(mesh->d_async_flag);
        cudaFreeHost(mesh->h_async_flag);
    }
    cudaFree(mesh->d_verts);
    cudaFree(mesh->d_idx);
    cudaFree(mesh->d_stage);
    delete mesh;
    return B200CD_OK;
}

// OBJ ingest with the reference parser's dialect (load_obj.h:24-103); the parser itself is obj_parse.cu.
API int b200cd_mesh_load_obj(b200cd_ctx* ctx, const char* path, b200cd_mesh** out) {
    if (!ctx || !path || !out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *out = nullptr;
    std::vector<float> xyz;
    std::vector<uint32_t> idx;
    std::string err;
    const int rc = obj_parse(path, xyz, idx, err);
    if (rc != B200CD_OK) return set_error(ctx, rc, err);
    return b200cd_mesh_from_arrays(ctx, xyz.data(), (uint32_t)(xyz.size() / 3), idx.data(), (uint32_t)(idx.size() / 3), out);
}

// The same parser without a GPU: host arrays out (malloc'ed; release with b200cd_host_array_free).
API int b200cd_obj_parse_host(const char* path, float** xyz_out, uint32_t* nverts_out, uint32_t** idx_out, uint32_t* ntris_out,
                              char* err_out, uint64_t err_len) {
    if (!path || !xyz_out || !nverts_out || !idx_out || !ntris_out) return B200CD_E_INVALID;
    *xyz_out = nullptr;
    *idx_out = nullptr;
    *nverts_out = *ntris_out = 0;
    std::vector<float> xyz;
    std::vector<uint32_t> idx;
    std::string err;
    const int rc = obj_parse(path, xyz, idx, err);
    if (err_out && err_len) {
        const size_t k = std::min<size_t>(err.size(), (size_t)err_len - 1);
        memcpy(err_out, err.data(), k);
        err_out[k] = '\0';
    }
    if (rc != B200CD_OK) return rc;
    float* x = static_cast<float*>(malloc(std::max<size_t>(xyz.size(), 1) * sizeof(float)));
    uint32_t* t = static_cast<uint32_t*>(malloc(std::max<size_t>(idx.size(), 1) * sizeof(uint32_t)));
    if (!x || !t) {
        free(x);
        free(t);
        return B200CD_E_NOMEM;
    }
    if (!xyz.empty()) memcpy(x, xyz.data(), xyz.size() * sizeof(float));
    if (!idx.empty()) memcpy(t, idx.data(), idx.size() * sizeof(uint32_t));
    *xyz_out = x;
    *idx_out = t;
    *nverts_out = (uint32_t)(xyz.size() / 3);
    *ntris_out = (uint32_t)(idx.size() / 3);
    return B200CD_OK;
}

API void b200cd_host_array_free(void* p) { free(p); }

// ------------------------------------------------------------------ build

API int b200cd_bvh_build(b200cd_ctx* ctx, const b200cd_mesh* mesh, const b200cd_params* params, b200cd_bvh** out) {
    if (!ctx || !mesh || !out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *out = nullptr;
    int rc = check_params(ctx, params);
    if (rc != B200CD_OK) return rc;
    DeviceGuard g(ctx->device);
    b200cd_bvh* b = nullptr;
    rc = alloc_bvh(ctx, mesh->ntris, mesh->nverts, true, &b);
    if (rc != B200CD_OK) return rc;
    rc = run_build(ctx, b, mesh, params);
    if (rc != B200CD_OK) {
        b200cd_bvh_destroy(b);
        return rc;
    }
    *out = b;
    return B200CD_OK;
}

API int b200cd_bvh_rebuild(b200cd_ctx* ctx, b200cd_bvh* bvh, const b200cd_mesh* mesh, const b200cd_params* params) {
    if (!ctx || !bvh || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (bvh->n != mesh->ntris || !bvh->d_keys[0]) return set_error(ctx, B200CD_E_INVALID, "BVH was not built for a mesh of this size");
    int rc = check_params(ctx, params);
    if (rc != B200CD_OK) return rc;
    DeviceGuard g(ctx->device);
    bvh->nverts = mesh->nverts;
    return run_build(ctx, bvh, mesh, params);
}

API int b200cd_bvh_refit(b200cd_ctx* ctx, b200cd_bvh* bvh, const b200cd_mesh* mesh) {
    if (!ctx || !bvh || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bvh->built || bvh->n != mesh->ntris || !bvh->d_keys[0]) return set_error(ctx, B200CD_E_INVALID, "BVH not built for this mesh");
    if (mesh->pending) return set_error(ctx, B200CD_E_INVALID, "mesh has an asynchronous upload in flight: call b200cd_mesh_wait first");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B0], s));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B1], s));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B2], s));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B3], s));
    const QFrame qframe = make_qframe(bvh->params);
    // the sorted keys are kept, so the same climb reproduces the same topology around the new boxes
    launch_build_tree(mesh->d_verts, mesh->d_idx, bvh->d_ids[bvh->cur], bvh->d_keys[bvh->cur], bvh->n, bvh->d_flags,
                      bvh->d_pairs, bvh->d_leaves, bvh->d_root_box, bvh->d_build_scratch, s, nullptr, bvh->d_block_boxes,
                      bvh->qvalid ? bvh->d_qpairs : nullptr, bvh->d_qframe, &qframe, bvh->root_valid);
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_B4], s));
    CD_CUDA(ctx, mark_consumed(mesh, s));
    CD_CUDA(ctx, cudaGetLastError());
    ctx->stats.ms_build = -1.f;
    ctx->stats.sort_passes = 0;
    return B200CD_OK;
}

API int b200cd_bvh_download(b200cd_ctx* ctx, const b200cd_bvh* bvh, b200cd_node32* nodes, uint64_t* sorted_keys,
                            uint32_t* sorted_ids) {
    if (!ctx || !bvh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bvh->built) return set_error(ctx, B200CD_E_INVALID, "BVH not built");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint32_t n = bvh->n;
    if (!n) return B200CD_OK;
    if (nodes) {
        b200cd_node32* d_nodes = nullptr;
        const uint64_t cnt = 2ull * n - 1;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d_nodes), cnt * sizeof(b200cd_node32)));
        cudaMemsetAsync(d_nodes, 0xff, cnt * sizeof(b200cd_node32), s);
        uint32_t* d_scratch = nullptr;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_scratch), 2 * cnt * sizeof(uint32_t));
        if (e == cudaSuccess) {
            launch_export_nodes(bvh->d_pairs, bvh->d_root_box, n, d_scratch, d_nodes, s);
            e = cudaMemcpyAsync(nodes, d_nodes, cnt * sizeof(b200cd_node32), cudaMemcpyDeviceToHost, s);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_nodes);
        cudaFree(d_scratch);
        CD_CUDA(ctx, e);
    }
    if (sorted_keys) {
        if (!bvh->d_keys[0]) return set_error(ctx, B200CD_E_INVALID, "this BVH was received, not built: no keys");
        CD_CUDA(ctx, cudaMemcpyAsync(sorted_keys, bvh->d_keys[bvh->cur], 8ull * n, cudaMemcpyDeviceToHost, s));
    }
    if (sorted_ids) CD_CUDA(ctx, cudaMemcpyAsync(sorted_ids, bvh->d_ids[bvh->cur], 4ull * n, cudaMemcpyDeviceToHost, s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));
    return B200CD_OK;
}

API int b200cd_bvh_validate(b200cd_ctx* ctx, const b200cd_bvh* bvh, const b200cd_mesh* mesh, b200cd_checks* out) {
    if (!ctx || !bvh || !out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bvh->built) return set_error(ctx, B200CD_E_INVALID, "BVH not built");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    uint32_t* d_chk = ctx->d_scalars + 16;
    uint32_t* d_scratch = nullptr;
    CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&d_scratch), 2 * (2ull * std::max<uint32_t>(bvh->n, 1u)) * sizeof(uint32_t)));
    launch_validate(bvh->d_pairs, bvh->d_leaves, bvh->d_root_box, bvh->d_keys[0] ? bvh->d_keys[bvh->cur] : nullptr, bvh->n,
                    mesh ? mesh->nverts : bvh->nverts, d_scratch, d_chk, s);
    cudaError_t e = cudaMemcpyAsync(ctx->h_scalars + 16, d_chk, 9 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_scratch);
    CD_CUDA(ctx, e);
    memcpy(out, ctx->h_scalars + 16, 9 * sizeof(uint32_t));
    return B200CD_OK;
}

API int b200cd_bvh_destroy(b200cd_bvh* bvh) {
    if (!bvh) return B200CD_OK;
    DeviceGuard g(bvh->ctx->device);
    cudaStreamSynchronize(bvh->ctx->stream);
    free_bvh_buffers(bvh);
    delete bvh;
    return B200CD_OK;
}

API int b200cd_bvh_alloc_like(b200cd_ctx* ctx, uint32_t ntris, b200cd_bvh** out) {
    if (!ctx || !out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *out = nullptr;
    DeviceGuard g(ctx->device);
    b200cd_bvh* b = nullptr;
    int rc = alloc_bvh(ctx, ntris, 0, false, &b);
    if (rc != B200CD_OK) return rc;
    b->built = true;  // contents arrive through the device views
    b->cur = 0;
    *out = b;
    return B200CD_OK;
}

// ------------------------------------------------------------------ partitioned (multi-GPU) build

API int b200cd_morton_keys_device(b200cd_ctx* ctx, const b200cd_mesh* mesh, const b200cd_params* params, uint32_t first,
                                  uint32_t count, void* d_keys_out) {
    if (!ctx || !mesh || (count && !d_keys_out)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    int rc = check_params(ctx, params);
    if (rc != B200CD_OK) return rc;
    if (params->auto_box) return set_error(ctx, B200CD_E_INVALID, "auto_box is not available for a partitioned build: pass the box");
    if ((uint64_t)first + count > mesh->ntris) return set_error(ctx, B200CD_E_INVALID, "triangle slice out of range");
    DeviceGuard g(ctx->device);
    if (mesh->pending) return set_error(ctx, B200CD_E_INVALID, "mesh has an asynchronous upload in flight: call b200cd_mesh_wait first");
    launch_morton(mesh->d_verts, mesh->d_idx, first, count, *params, nullptr, static_cast<uint64_t*>(d_keys_out), ctx->stream);
    CD_CUDA(ctx, mark_consumed(mesh, ctx->stream));
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

API int b200cd_key_histogram_device(b200cd_ctx* ctx, const void* d_keys, uint32_t count, int32_t shift, void* d_hist65536) {
    if (!ctx || (count && !d_keys) || !d_hist65536 || shift < 0 || shift > 63) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    DeviceGuard g(ctx->device);
    launch_key_hist16(static_cast<const uint64_t*>(d_keys), count, shift, static_cast<uint32_t*>(d_hist65536), ctx->sm_count,
                      ctx->stream);
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

API int b200cd_partition_plan_device(b200cd_ctx* ctx, const void* d_global_hist65536, const void* d_local_hist65536, int32_t shift,
                                     uint32_t world, void* d_splitters_out, void* d_counts_out) {
    if (!ctx || !d_global_hist65536 || !d_local_hist65536 || !d_counts_out || (world > 1 && !d_splitters_out) || world == 0 ||
        world > RS_MAX_SPLIT_P1 || shift < 0 || shift > 47)
        return set_error(ctx, B200CD_E_INVALID, "bad argument");
    DeviceGuard g(ctx->device);
    launch_partition_plan(static_cast<const uint32_t*>(d_global_hist65536), static_cast<const uint32_t*>(d_local_hist65536), shift,
                          (int)world, static_cast<uint64_t*>(d_splitters_out), static_cast<int32_t*>(d_counts_out), ctx->stream);
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

API int b200cd_bvh_alloc_partial(b200cd_ctx* ctx, uint32_t capacity, uint64_t ghost_capacity, uint32_t max_peers,
                                 b200cd_bvh** out) {
    if (!ctx || !out || max_peers > 32) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    *out = nullptr;
    if (capacity > (1u << B200CD_MAX_TRIS_LOG2) || capacity + ghost_capacity > (1ull << 31))
        return set_error(ctx, B200CD_E_TOOBIG, "partial BVH too large");
    DeviceGuard g(ctx->device);
    b200cd_bvh* b = nullptr;
    int rc = alloc_bvh(ctx, capacity, 0, true, &b, ghost_capacity, max_peers);
    if (rc != B200CD_OK) return rc;
    b->n = 0;
    *out = b;
    return B200CD_OK;
}

API int b200cd_bvh_key_buffers(b200cd_ctx* ctx, b200cd_bvh* bvh, void** d_keys, void** d_ids, uint32_t* capacity) {
    if (!ctx || !bvh || !bvh->d_keys[0]) return set_error(ctx, B200CD_E_INVALID, "BVH has no key buffers");
    if (d_keys) *d_keys = bvh->d_keys[0];
    if (d_ids) *d_ids = bvh->d_ids[0];
    if (capacity) *capacity = bvh->cap;
    return B200CD_OK;
}

API int b200cd_partition_keys_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_keys, uint32_t first_id, uint32_t count,
                                     const void* d_splitters, uint32_t nsplit, void* d_keys_out, void* d_ids_out,
                                     uint64_t* counts_out) {
    if (!ctx || !bvh || !counts_out || (count && (!d_keys || !d_keys_out || !d_ids_out)) || (nsplit && !d_splitters) || nsplit > 15)
        return set_error(ctx, B200CD_E_INVALID, "bad argument");
    if (radix_tile_status_words(count, 1) > bvh->tile_status_words) return set_error(ctx, B200CD_E_INVALID, "slice larger than the BVH's scratch");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    radix_partition(static_cast<const uint64_t*>(d_keys), nullptr, first_id, static_cast<uint64_t*>(d_keys_out),
                    static_cast<uint32_t*>(d_ids_out), count, static_cast<const uint64_t*>(d_splitters), (int)nsplit, bvh->d_hist,
                    bvh->d_tile_status, ctx->sm_count, s);
    CD_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, bvh->d_hist + (1 << radix_digit_bits()), sizeof(uint32_t) * (nsplit + 1), cudaMemcpyDeviceToHost, s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));
    CD_CUDA(ctx, cudaGetLastError());
    for (uint32_t d = 0; d <= nsplit; ++d) counts_out[d] = ctx->h_scalars[d];
    return B200CD_OK;
}

API int b200cd_bvh_build_partial(b200cd_ctx* ctx, b200cd_bvh* bvh, const b200cd_mesh* mesh, const b200cd_params* params,
                                 uint32_t count) {
    if (!ctx || !bvh || !mesh) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bvh->d_keys[0] || count > bvh->cap) return set_error(ctx, B200CD_E_CAPACITY, "more triangles than the partial BVH was allocated for");
    int rc = check_params(ctx, params);
    if (rc != B200CD_OK) return rc;
    DeviceGuard g(ctx->device);
    bvh->n = count;
    bvh->nverts = mesh->nverts;
    return run_build(ctx, bvh, mesh, params, /*keys_given*/ true);
}

API int b200cd_bvh_chunk_boxes_device(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t K, void* d_boxes_out) {
    if (!ctx || !bvh || !d_boxes_out || K == 0 || K > (uint32_t)ghost_max_k()) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    if (!bvh->built) return set_error(ctx, B200CD_E_INVALID, "BVH not built");
    DeviceGuard g(ctx->device);
    if (!bvh->d_cut_scratch) return set_error(ctx, B200CD_E_INVALID, "BVH was not allocated for a partitioned build");
    launch_chunk_boxes(bvh->d_pairs, bvh->d_root_box, bvh->n, K, bvh->d_cut_scratch, static_cast<float*>(d_boxes_out), ctx->stream);
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

API int b200cd_select_ghosts_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_peer_boxes, uint32_t npeers, uint32_t K,
                                    uint32_t peer_mask, const void** d_ghosts_out, uint64_t* stride_out, uint64_t* counts_out) {
    if (!ctx || !bvh || !d_peer_boxes || !counts_out || npeers == 0 || npeers > 32 || K == 0 || K > (uint32_t)ghost_max_k())
        return set_error(ctx, B200CD_E_INVALID, "bad argument");
    if (!bvh->built || !bvh->d_ghost_out) return set_error(ctx, B200CD_E_INVALID, "BVH was not allocated for a partitioned build");
    // one outgoing list per peer was allocated: a larger npeers (or mask bits at or above it) would write past them
    if (npeers > bvh->max_peers) return set_error(ctx, B200CD_E_INVALID, "npeers exceeds the max_peers given to b200cd_bvh_alloc_partial");
    if (npeers < 32 && (peer_mask >> npeers)) return set_error(ctx, B200CD_E_INVALID, "peer_mask has bits at or above npeers");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(ctx->d_scalars);  // 32 x u64 = 64 words
    launch_ghosts(bvh->d_leaves, bvh->n, static_cast<const float*>(d_peer_boxes), npeers, K, peer_mask, bvh->d_ghost_out,
                  bvh->ghost_out_cap, d_counts, reinterpret_cast<float*>(bvh->d_cut_scratch), bvh->d_block_boxes, s);  // (cut scratch: free by now)
    CD_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, d_counts, sizeof(unsigned long long) * npeers, cudaMemcpyDeviceToHost, s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));
    CD_CUDA(ctx, cudaGetLastError());
    bool over = false;
    for (uint32_t p = 0; p < npeers; ++p) {
        counts_out[p] = reinterpret_cast<unsigned long long*>(ctx->h_scalars)[p];
        over = over || counts_out[p] > bvh->ghost_out_cap;
    }
    if (d_ghosts_out) *d_ghosts_out = bvh->d_ghost_out;
    if (stride_out) *stride_out = bvh->ghost_out_cap;
    if (over) return set_error(ctx, B200CD_E_CAPACITY, "ghost list larger than ghost_capacity");
    return B200CD_OK;
}

API int b200cd_bvh_ghost_buffer(b200cd_ctx* ctx, b200cd_bvh* bvh, void** d_ptr, uint64_t* capacity) {
    if (!ctx || !bvh || !d_ptr) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    *d_ptr = bvh->d_leaves + bvh->cap;  // fixed place: after the room for local leaves
    if (capacity) *capacity = bvh->ghost_cap;
    return B200CD_OK;
}

// ---- peer memory (NVLink): ranks write (key, id) and ghost records straight into each other's buffers

API int b200cd_ipc_export(b200cd_ctx* ctx, b200cd_bvh* bvh, uint8_t* handles_out /* 4 x 64 bytes */, uint64_t* offsets_out /* 4 */) {
    if (!ctx || !bvh || !handles_out || !offsets_out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (!bvh->d_peers) return set_error(ctx, B200CD_E_INVALID, "BVH was not allocated for a partitioned build");
    DeviceGuard g(ctx->device);
    void* ptrs[4] = {bvh->d_keys[0], bvh->d_ids[0], bvh->d_leaves, bvh->d_ghost_in_count};
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    for (int i = 0; i < 4; ++i) {
        cudaIpcMemHandle_t h;
        CD_CUDA(ctx, cudaIpcGetMemHandle(&h, ptrs[i]));
        memcpy(handles_out + 64 * i, &h, 64);
        // cudaIpcOpenMemHandle maps the whole allocation: remember where inside it the buffer starts
        void* base = nullptr;
        size_t size = 0;
        if (cuMemGetAddressRange_shim(&base, &size, ptrs[i]) != 0) return set_error(ctx, B200CD_E_CUDA, "address range query failed");
        offsets_out[i] = (uint64_t)((char*)ptrs[i] - (char*)base);
    }
    offsets_out[2] += sizeof(LeafRec) * (uint64_t)bvh->cap;  // peers address my GHOST records
    return B200CD_OK;
}

API int b200cd_ipc_open(b200cd_ctx* ctx, const uint8_t* handle64, void** d_ptr_out) {
    if (!ctx || !handle64 || !d_ptr_out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CD_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return B200CD_OK;
}

API int b200cd_ipc_close(b200cd_ctx* ctx, void* d_ptr) {
    if (!ctx) return B200CD_E_INVALID;
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return B200CD_OK;
}

// ---- the mesh itself over peer memory: every rank uploads 1/ranks of a frame over its own PCIe link and copies
// that slice into the other ranks' mesh buffers with the copy engines (NVLink), all on the copy stream, so the
// next frame travels while the current one is being built and queried (b200cd_mesh_update_slice_async).

API int b200cd_mesh_ipc_export(b200cd_ctx* ctx, b200cd_mesh* mesh, uint8_t* handles_out /* 2 x 64 bytes */, uint64_t* offsets_out /* 2 */) {
    if (!ctx || !mesh || !handles_out || !offsets_out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    void* ptrs[2] = {mesh->d_verts, mesh->d_idx};
    for (int i = 0; i < 2; ++i) {
        cudaIpcMemHandle_t h;
        CD_CUDA(ctx, cudaIpcGetMemHandle(&h, ptrs[i]));
        memcpy(handles_out + 64 * i, &h, 64);
        void* base = nullptr;
        size_t size = 0;
        if (cuMemGetAddressRange_shim(&base, &size, ptrs[i]) != 0) return set_error(ctx, B200CD_E_CUDA, "address range query failed");
        offsets_out[i] = (uint64_t)((char*)ptrs[i] - (char*)base);
    }
    return B200CD_OK;
}

/* peers[2 * r + {0,1}] = rank r's vertex (float4) and index buffers as mapped on THIS GPU (b200cd_ipc_open + offset);
 * the entries of my_rank are ignored. */
API int b200cd_mesh_set_peers(b200cd_ctx* ctx, b200cd_mesh* mesh, uint32_t nranks, uint32_t my_rank, void* const* peers) {
    if (!ctx || !mesh || !peers || nranks == 0 || nranks > 16 || my_rank >= nranks) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    mesh->npeers = nranks;
    mesh->my_rank = my_rank;
    for (uint32_t r = 0; r < nranks; ++r) {
        mesh->peer_verts[r] = r == my_rank ? mesh->d_verts : static_cast<float4*>(peers[2 * r]);
        mesh->peer_idx[r] = r == my_rank ? mesh->d_idx : static_cast<uint32_t*>(peers[2 * r + 1]);
        if (!mesh->peer_verts[r] || !mesh->peer_idx[r]) return set_error(ctx, B200CD_E_INVALID, "NULL peer buffer");
    }
    return B200CD_OK;
}

/* Asynchronous twin of b200cd_mesh_update_slice for double-buffered frames on several GPUs: the slice goes
 * host -> my mesh (H2D, expansion, index check) and from there into every peer's copy of the mesh (D2D over
 * NVLink, copy engines), all on the copy stream. b200cd_mesh_wait(mesh) returns when MY slice has landed
 * everywhere; a barrier across the ranks after it (any collective) then guarantees the whole frame is in place.
 * The caller must not start this before every rank has finished the last build that read this mesh object
 * (true at the start of step k for the mesh of step k-1 when the steps end with a collective). */
API int b200cd_mesh_update_slice_async(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, uint32_t first_vert, uint32_t nverts,
                                       const uint32_t* tri_idx, uint32_t first_tri, uint32_t ntris) {
    if (!ctx || !mesh || (nverts && !xyz) || (ntris && !tri_idx)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if ((uint64_t)first_vert + nverts > mesh->nverts || (uint64_t)first_tri + ntris > mesh->ntris)
        return set_error(ctx, B200CD_E_INVALID, "slice out of range");
    if (mesh->pending) return set_error(ctx, B200CD_E_INVALID, "mesh already has an asynchronous upload in flight");
    DeviceGuard g(ctx->device);
    if (!ctx->copy_stream) CD_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!mesh->ev_ready) {
        CD_CUDA(ctx, cudaEventCreateWithFlags(&mesh->ev_ready, cudaEventDisableTiming));
        CD_CUDA(ctx, cudaEventCreateWithFlags(&mesh->ev_consumed, cudaEventDisableTiming));
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&mesh->d_async_flag), sizeof(uint32_t)));
        CD_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&mesh->h_async_flag), sizeof(uint32_t)));
        CD_CUDA(ctx, mark_consumed(mesh, ctx->stream));
    }
    if (nverts && !mesh->d_stage) CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&mesh->d_stage), 12ull * mesh->nverts));
    cudaStream_t c = ctx->copy_stream;
    constexpr int NPUSH = 4;
    for (int i = 0; i < NPUSH; ++i) {
        if (!ctx->push_stream[i]) {
            CD_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->push_stream[i], cudaStreamNonBlocking));
            CD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->push_ev[i], cudaEventDisableTiming));
        }
    }
    if (mesh->consumed_valid) CD_CUDA(ctx, cudaStreamWaitEvent(c, mesh->ev_consumed, 0));
    *mesh->h_async_flag = 0;
    // The slice moves in NPUSH chunks: chunk i's H2D (+ expansion) runs on the copy stream while push stream i
    // copies the chunks that have landed into the peers' meshes - the PCIe transfer and the NVLink pushes overlap,
    // and the pushes of different chunks use different copy engines.
    if (ntris) CD_CUDA(ctx, cudaMemsetAsync(mesh->d_async_flag, 0, sizeof(uint32_t), c));
    for (int i = 0; i < NPUSH; ++i) {
        const uint32_t v0 = (uint32_t)((uint64_t)nverts * i / NPUSH), v1 = (uint32_t)((uint64_t)nverts * (i + 1) / NPUSH);
        const uint32_t t0 = (uint32_t)((uint64_t)ntris * i / NPUSH), t1 = (uint32_t)((uint64_t)ntris * (i + 1) / NPUSH);
        if (v1 > v0) {
            CD_CUDA(ctx, cudaMemcpyAsync(mesh->d_stage + 3ull * v0, xyz + 3ull * v0, 12ull * (v1 - v0), cudaMemcpyHostToDevice, c));
            launch_expand_verts(mesh->d_stage + 3ull * v0, mesh->d_verts + first_vert + v0, v1 - v0, c);
        }
        if (t1 > t0) {
            CD_CUDA(ctx, cudaMemcpyAsync(mesh->d_idx + 3ull * (first_tri + t0), tri_idx + 3ull * t0, 12ull * (t1 - t0), cudaMemcpyHostToDevice, c));
            launch_check_idx_accumulate(mesh->d_idx + 3ull * (first_tri + t0), t1 - t0, mesh->nverts, mesh->d_async_flag, ctx->sm_count, c);
        }
        CD_CUDA(ctx, cudaEventRecord(ctx->push_ev[i], c));
        cudaStream_t ps = ctx->push_stream[i];
        CD_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->push_ev[i], 0));
        for (uint32_t d = 1; d < mesh->npeers; ++d) {  // staggered: rank r starts with r+1, so no two ranks hit the same GPU at once
            const uint32_t r = (mesh->my_rank + d) % mesh->npeers;
            if (v1 > v0)
                CD_CUDA(ctx, cudaMemcpyAsync(mesh->peer_verts[r] + first_vert + v0, mesh->d_verts + first_vert + v0,
                                             sizeof(float4) * (size_t)(v1 - v0), cudaMemcpyDefault, ps));
            if (t1 > t0)
                CD_CUDA(ctx, cudaMemcpyAsync(mesh->peer_idx[r] + 3ull * (first_tri + t0), mesh->d_idx + 3ull * (first_tri + t0),
                                             12ull * (t1 - t0), cudaMemcpyDefault, ps));
        }
    }
    if (ntris) CD_CUDA(ctx, cudaMemcpyAsync(mesh->h_async_flag, mesh->d_async_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c));
    for (int i = 0; i < NPUSH; ++i) {  // the copy stream (and ev_ready on it) waits for every push stream
        CD_CUDA(ctx, cudaEventRecord(ctx->push_ev[i], ctx->push_stream[i]));
        CD_CUDA(ctx, cudaStreamWaitEvent(c, ctx->push_ev[i], 0));
    }
    CD_CUDA(ctx, cudaEventRecord(mesh->ev_ready, c));
    CD_CUDA(ctx, cudaGetLastError());
    mesh->pending = true;
    return B200CD_OK;
}

/* peers[4 * r + {0,1,2,3}] = rank r's key buffer, id buffer, ghost records, ghost counter as seen from THIS GPU
 * (the rank's own entries = its local buffers; pass NULLs for r == my_rank to have them filled in). */
API int b200cd_bvh_set_peers(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t nranks, uint32_t my_rank, void* const* peers) {
    if (!ctx || !bvh || !peers || nranks == 0 || nranks > RS_MAX_SPLIT_P1 || my_rank >= nranks)
        return set_error(ctx, B200CD_E_INVALID, "bad argument");
    if (!bvh->d_peers) return set_error(ctx, B200CD_E_INVALID, "BVH was not allocated for a partitioned build");
    DeviceGuard g(ctx->device);
    PeerTable t;
    memset(&t, 0, sizeof t);
    for (uint32_t r = 0; r < nranks; ++r) {
        if (r == my_rank) {
            t.keys[r] = bvh->d_keys[0];
            t.ids[r] = bvh->d_ids[0];
            t.ghosts[r] = bvh->d_leaves + bvh->cap;
            t.ghost_count[r] = bvh->d_ghost_in_count;
        } else {
            t.keys[r] = static_cast<uint64_t*>(peers[4 * r + 0]);
            t.ids[r] = static_cast<uint32_t*>(peers[4 * r + 1]);
            t.ghosts[r] = static_cast<LeafRec*>(peers[4 * r + 2]);
            t.ghost_count[r] = static_cast<unsigned long long*>(peers[4 * r + 3]);
        }
    }
    t.ghost_cap = bvh->ghost_cap;
    CD_CUDA(ctx, cudaMemcpyAsync(bvh->d_peers, &t, sizeof t, cudaMemcpyHostToDevice, ctx->stream));
    CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bvh->nranks = nranks;
    return B200CD_OK;
}

API int b200cd_partition_counts_device(b200cd_ctx* ctx, const void* d_keys, uint32_t count, const void* d_splitters,
                                       uint32_t nsplit, void* d_counts_out) {
    if (!ctx || !d_counts_out || (count && !d_keys) || (nsplit && !d_splitters) || nsplit > 15) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    DeviceGuard g(ctx->device);
    radix_partition_counts(static_cast<const uint64_t*>(d_keys), count, static_cast<const uint64_t*>(d_splitters), (int)nsplit,
                           static_cast<uint32_t*>(d_counts_out), ctx->sm_count, ctx->stream);
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

API int b200cd_partition_to_peers_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_keys, uint32_t first_id, uint32_t count,
                                         const void* d_splitters, uint32_t nsplit, const void* d_recv_offsets) {
    if (!ctx || !bvh || !d_recv_offsets || (count && !d_keys) || (nsplit && !d_splitters) || nsplit > 15) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    if (!bvh->d_peers || bvh->nranks == 0) return set_error(ctx, B200CD_E_INVALID, "b200cd_bvh_set_peers has not been called");
    if (nsplit + 1 != bvh->nranks) return set_error(ctx, B200CD_E_INVALID, "nsplit + 1 differs from the nranks given to b200cd_bvh_set_peers");
    if (radix_tile_status_words(count, 1) > bvh->tile_status_words) return set_error(ctx, B200CD_E_INVALID, "slice larger than the BVH's scratch");
    DeviceGuard g(ctx->device);
    radix_partition_to_peers(static_cast<const uint64_t*>(d_keys), first_id, count, static_cast<const uint64_t*>(d_splitters),
                             (int)nsplit, bvh->d_peers, static_cast<const uint32_t*>(d_recv_offsets), bvh->d_hist,
                             bvh->d_tile_status, ctx->stream);
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

API int b200cd_ghost_counter_reset(b200cd_ctx* ctx, b200cd_bvh* bvh) {
    if (!ctx || !bvh || !bvh->d_ghost_in_count) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaMemsetAsync(bvh->d_ghost_in_count, 0, sizeof(unsigned long long), ctx->stream));
    return B200CD_OK;
}

API int b200cd_ghost_counter_read(b200cd_ctx* ctx, b200cd_bvh* bvh, uint64_t* count_out) {
    if (!ctx || !bvh || !count_out || !bvh->d_ghost_in_count) return set_error(ctx, B200CD_E_INVALID, "bad argument");
    DeviceGuard g(ctx->device);
    CD_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, bvh->d_ghost_in_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count_out = *reinterpret_cast<unsigned long long*>(ctx->h_scalars);
    if (*count_out > bvh->ghost_cap) return set_error(ctx, B200CD_E_CAPACITY, "more ghosts arrived than ghost_capacity");
    return B200CD_OK;
}

API int b200cd_send_ghosts_to_peers_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_peer_boxes, uint32_t npeers, uint32_t K,
                                           uint32_t peer_mask) {
    if (!ctx || !bvh || !d_peer_boxes || npeers == 0 || npeers > RS_MAX_SPLIT_P1 || K == 0 || K > (uint32_t)ghost_max_k())
        return set_error(ctx, B200CD_E_INVALID, "bad argument");
    if (!bvh->built || !bvh->d_peers) return set_error(ctx, B200CD_E_INVALID, "BVH not built / peers not set");
    // the peer table holds entries for the ranks given to b200cd_bvh_set_peers only: anything else would index NULLs
    if (npeers != bvh->nranks) return set_error(ctx, B200CD_E_INVALID, "npeers differs from the nranks given to b200cd_bvh_set_peers");
    if (npeers < 32 && (peer_mask >> npeers)) return set_error(ctx, B200CD_E_INVALID, "peer_mask has bits at or above npeers");
    DeviceGuard g(ctx->device);
    if (!bvh->d_cut_scratch) return set_error(ctx, B200CD_E_INVALID, "BVH was not allocated for a partitioned build");
    launch_ghosts_to_peers(bvh->d_leaves, bvh->n, static_cast<const float*>(d_peer_boxes), npeers, K, peer_mask, bvh->d_peers,
                           reinterpret_cast<float*>(bvh->d_cut_scratch), bvh->d_block_boxes, ctx->stream, bvh->d_ghost_list, bvh->ghost_list_cap, ctx->sm_count);
    CD_CUDA(ctx, cudaGetLastError());
    return B200CD_OK;
}

// ------------------------------------------------------------------ query

namespace b200cd {  // query internals shared with dist.cu (declared in internal.cuh)

int grow(b200cd_ctx* ctx, uint2** buf, uint64_t* cap, uint64_t want) {
    if (*cap >= want && *buf) return B200CD_OK;
    cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(buf), std::max<uint64_t>(want, 1) * sizeof(uint2)));
    *cap = want;
    return B200CD_OK;
}

int sort_pairs_impl(b200cd_ctx* ctx, uint2** d_pairs, uint2** d_tmp, uint64_t count, int id_bits, uint32_t** d_hist,
                    uint32_t** d_status, uint64_t* status_words, cudaStream_t s) {
    if (count < 2) return B200CD_OK;
    if (small_pair_sort(*d_pairs, count, s)) return B200CD_OK;  // a few thousand pairs: one launch instead of ten
    if (count >= (1ull << 30)) return set_error(ctx, B200CD_E_TOOBIG, "pair list too long to sort");
    // memory word of a pair {lo_id, hi_id} read as u64 = hi_id << 32 | lo_id: LSD order = hi_id digits, then lo_id digits
    RadixPass passes[8];
    int np = 0;
    const int db = radix_digit_bits();
    for (int sh = 0; sh < id_bits; sh += db) passes[np++] = {32 + sh, std::min(db, id_bits - sh)};
    for (int sh = 0; sh < id_bits; sh += db) passes[np++] = {sh, std::min(db, id_bits - sh)};
    uint64_t need = radix_tile_status_words((uint32_t)count, np);
    if (need > *status_words) {
        cudaFree(*d_status);
        *d_status = nullptr;
        *status_words = 0;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(d_status), need * sizeof(uint32_t)));
        *status_words = need;
    }
    uint64_t* keys[2] = {reinterpret_cast<uint64_t*>(*d_pairs), reinterpret_cast<uint64_t*>(*d_tmp)};
    int cur = radix_sort(keys, nullptr, (uint32_t)count, passes, np, false, *d_hist, *d_status, *status_words,
                         ctx->sm_count, s);
    if (cur == 1) std::swap(*d_pairs, *d_tmp);
    return B200CD_OK;
}

int ensure_entry_lists(b200cd_ctx* ctx, b200cd_bvh* b, uint64_t nquery) {
    const uint64_t qblocks = (nquery + B200CD_QUERY_GROUP - 1) / B200CD_QUERY_GROUP;
    if (qblocks > b->entry_blocks) {
        cudaFree(b->d_entries);
        cudaFree(b->d_entry_count);
        b->d_entries = nullptr;
        b->d_entry_count = nullptr;
        b->entry_blocks = 0;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->d_entries), qblocks * B200CD_MAX_ENTRIES * sizeof(Node32)));
        // counts, then (32-byte aligned) one 8-float union box per group
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&b->d_entry_count), (((qblocks + 7) & ~7ull) + 8 * qblocks) * sizeof(uint32_t)));
        b->entry_blocks = qblocks;
    }
    return B200CD_OK;
}

int run_query(b200cd_ctx* ctx, b200cd_bvh* b, uint32_t shard, uint32_t nshards, uint32_t chunk, int sorted,
              uint64_t* count_out) {
    if (!b->built) return set_error(ctx, B200CD_E_INVALID, "BVH not built");
    if (nshards == 0 || shard >= nshards) return set_error(ctx, B200CD_E_INVALID, "shard must be < nshards");
    cudaStream_t s = ctx->stream;
    const uint32_t n = b->n;
    ctx->stats.query_retries = 0;
    ctx->stats.candidates = ctx->stats.pairs = 0;
    *count_out = 0;
    if (n < 2) {
        CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q0], s));
        for (int e : {EV_Q1, EV_Q2, EV_Q3}) CD_CUDA(ctx, cudaEventRecord(ctx->ev[e], s));
        CD_CUDA(ctx, cudaStreamSynchronize(s));
        ctx->stats.ms_traverse = ctx->stats.ms_narrow = ctx->stats.ms_pair_sort = ctx->stats.ms_query = 0.f;
        return B200CD_OK;
    }
    // query threads of this shard: whole chunks c = shard, shard + nshards, ... A chunk is a whole number of
    // traversal blocks (B200CD_QUERY_BLOCK consecutive sorted leaves), so it is rounded up to a multiple of that.
    if (chunk == 0) chunk = (n + nshards - 1) / nshards;
    if (chunk > 0xffffff00u) return set_error(ctx, B200CD_E_INVALID, "chunk too large");
    chunk = (chunk + B200CD_QUERY_BLOCK - 1) / B200CD_QUERY_BLOCK * B200CD_QUERY_BLOCK;
    const uint64_t nchunks = ((uint64_t)n + chunk - 1) / chunk;
    const uint64_t my_chunks = nchunks > shard ? (nchunks - shard + nshards - 1) / nshards : 0;
    const uint64_t nquery64 = my_chunks * chunk;
    if (nquery64 > 0xffffffffull) return set_error(ctx, B200CD_E_INVALID, "chunk too large");
    const uint32_t nquery = (uint32_t)nquery64;

    const uint64_t hint = b->params.pair_capacity_hint;
    int rc = grow(ctx, &b->d_cand, &b->cand_cap, std::max<uint64_t>(b->cand_cap, std::max<uint64_t>(4ull * nquery + 4096, 4 * hint)));
    if (rc == B200CD_OK) rc = grow(ctx, &b->d_out, &b->out_cap, std::max<uint64_t>(b->out_cap, std::max<uint64_t>(nquery / 2 + 4096, hint)));
    if (rc == B200CD_OK && sorted) rc = grow(ctx, &b->d_out_tmp, &b->out_tmp_cap, b->out_cap);
    if (rc != B200CD_OK) return rc;
    rc = ensure_entry_lists(ctx, b, nquery);
    if (rc != B200CD_OK) return rc;

    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q0], s));
    trace_mark("query_begin", s);
    bool need_broad = true;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (need_broad) {
            CD_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, 8 * sizeof(unsigned long long), s));
            launch_broad(b->d_pairs, b->d_leaves, b->d_root_box, n, shard, nshards, chunk, nquery, /*foreign*/ 0, 0u, b->d_entries,
                         b->d_entry_count, b->d_cand, b->cand_cap, b->d_counters, s, nullptr, ctx->sm_count, !b->unshared_verts,
                         b->qvalid ? b->d_qpairs : nullptr, b->d_qframe);
            CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q1], s));
        } else {
            CD_CUDA(ctx, cudaMemsetAsync(b->d_counters + 1, 0, sizeof(unsigned long long), s));
        }
        launch_narrow(b->d_leaves, b->d_cand, b->cand_cap, b->d_out, b->out_cap, b->d_counters, ctx->sm_count, s, b->unshared_verts);
        CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q2], s));
        CD_CUDA(ctx, cudaMemcpyAsync(b->h_counters, b->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaStreamSynchronize(s));
        CD_CUDA(ctx, cudaGetLastError());
        const uint64_t ncand = b->h_counters[0], npair = b->h_counters[1];
        if (b->h_counters[2] & 1ull)
            return set_error(ctx, B200CD_E_DEPTH, "traversal stack of " + std::to_string(B200CD_MAX_STACK) + " entries exhausted");
        if (ncand > b->cand_cap) {  // candidate list overflowed: grow to the exact need and redo the traversal
            rc = grow(ctx, &b->d_cand, &b->cand_cap, ncand + ncand / 16);
            if (rc != B200CD_OK) return rc;
            need_broad = true;
            ctx->stats.query_retries++;
            continue;
        }
        if (npair > b->out_cap) {  // result list overflowed: candidates are intact, redo the narrow phase only
            rc = grow(ctx, &b->d_out, &b->out_cap, npair + npair / 16);
            if (rc != B200CD_OK) return rc;
            need_broad = false;
            ctx->stats.query_retries++;
            continue;
        }
        ctx->stats.candidates = ncand;
        ctx->stats.pairs = npair;
        ctx->stats.nodes_visited = b->h_counters[3];
        ctx->stats.warp_steps = b->h_counters[4];
        ctx->stats.start_entries = b->h_counters[5];
        b->npairs = npair;
        *count_out = npair;
        if (sorted && npair > 1) {
            rc = grow(ctx, &b->d_out_tmp, &b->out_tmp_cap, b->out_cap);
            if (rc != B200CD_OK) return rc;
            rc = sort_pairs_impl(ctx, &b->d_out, &b->d_out_tmp, npair, id_bits_for(b->id_space ? b->id_space : n), &b->d_hist, &b->d_tile_status,
                                 &b->tile_status_words, s);
            if (rc != B200CD_OK) return rc;
        }
        trace_mark("pair sort", s);
        CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_Q3], s));
        CD_CUDA(ctx, cudaStreamSynchronize(s));
        CD_CUDA(ctx, cudaGetLastError());
        ctx->stats.ms_traverse = ev_ms(ctx, EV_Q0, EV_Q1);
        ctx->stats.ms_narrow = ev_ms(ctx, EV_Q1, EV_Q2);
        ctx->stats.ms_pair_sort = ev_ms(ctx, EV_Q2, EV_Q3);
        ctx->stats.ms_query = ev_ms(ctx, EV_Q0, EV_Q3);
        return B200CD_OK;
    }
    return set_error(ctx, B200CD_E_CUDA, "query did not converge after growing its buffers");
}

// Ghost queries of a partitioned build: the nghost records stored after the local leaves are
// tested against the whole local tree. keep != 0 appends to the pair list of the last local query.
int run_ghost_query(b200cd_ctx* ctx, b200cd_bvh* b, uint64_t nghost, int keep, uint64_t* count_out) {
    if (!b->built) return set_error(ctx, B200CD_E_INVALID, "BVH not built");
    cudaStream_t s = ctx->stream;
    const uint32_t n = b->n;
    const uint64_t base = keep ? b->npairs : 0;
    *count_out = base;
    if (nghost > b->ghost_cap) return set_error(ctx, B200CD_E_CAPACITY, "more ghosts than the BVH has room for");
    if (nghost == 0 || n == 0) return B200CD_OK;
    if (nghost > 0xffffff00ull) return set_error(ctx, B200CD_E_TOOBIG, "too many ghost queries");
    const uint32_t nquery = (uint32_t)nghost;
    int rc = grow(ctx, &b->d_cand, &b->cand_cap, std::max<uint64_t>(b->cand_cap, 4ull * nquery + 4096));
    if (rc != B200CD_OK) return rc;
    auto grow_out_keep = [&](uint64_t want) -> int {  // enlarge d_out without losing the first `base` pairs
        if (b->out_cap >= want && b->d_out) return B200CD_OK;
        uint2* bigger = nullptr;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&bigger), want * sizeof(uint2)));
        if (base && b->d_out) {
            cudaError_t e = cudaMemcpyAsync(bigger, b->d_out, base * sizeof(uint2), cudaMemcpyDeviceToDevice, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { cudaFree(bigger); CD_CUDA(ctx, e); }
        }
        cudaFree(b->d_out);
        b->d_out = bigger;
        b->out_cap = want;
        return B200CD_OK;
    };
    rc = grow_out_keep(base + nquery / 2 + 4096);
    if (rc != B200CD_OK) return rc;
    bool need_broad = true;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (need_broad) CD_CUDA(ctx, cudaMemsetAsync(b->d_counters, 0, 8 * sizeof(unsigned long long), s));
        b->h_counters[7] = base;
        CD_CUDA(ctx, cudaMemcpyAsync(b->d_counters + 1, b->h_counters + 7, sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        if (need_broad)
            launch_broad(b->d_pairs, b->d_leaves, b->d_root_box, n, 0, 1, B200CD_QUERY_BLOCK, nquery, /*foreign*/ 1, b->cap, b->d_entries,
                         b->d_entry_count, b->d_cand, b->cand_cap, b->d_counters, s);
        launch_narrow(b->d_leaves, b->d_cand, b->cand_cap, b->d_out, b->out_cap, b->d_counters, ctx->sm_count, s, b->unshared_verts);
        CD_CUDA(ctx, cudaMemcpyAsync(b->h_counters, b->d_counters, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CD_CUDA(ctx, cudaStreamSynchronize(s));
        CD_CUDA(ctx, cudaGetLastError());
        const uint64_t ncand = b->h_counters[0], npair = b->h_counters[1];
        if (b->h_counters[2] & 1ull)
            return set_error(ctx, B200CD_E_DEPTH, "traversal stack of " + std::to_string(B200CD_MAX_STACK) + " entries exhausted");
        if (ncand > b->cand_cap) {
            rc = grow(ctx, &b->d_cand, &b->cand_cap, ncand + ncand / 16);
            if (rc != B200CD_OK) return rc;
            need_broad = true;
            continue;
        }
        if (npair > b->out_cap) {
            rc = grow_out_keep(npair + npair / 16);
            if (rc != B200CD_OK) return rc;
            need_broad = false;
            continue;
        }
        ctx->stats.candidates += ncand;
        ctx->stats.pairs = npair;
        b->npairs = npair;
        *count_out = npair;
        return B200CD_OK;
    }
    return set_error(ctx, B200CD_E_CUDA, "ghost query did not converge after growing its buffers");
}

}  // namespace b200cd

API int b200cd_collide_ghosts_device(b200cd_ctx* ctx, b200cd_bvh* bvh, uint64_t nghost, int keep_pairs,
                                     const void** d_pairs_out, uint64_t* count_out) {
    if (!ctx || !bvh || !count_out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    int rc = run_ghost_query(ctx, bvh, nghost, keep_pairs, count_out);
    if (d_pairs_out) *d_pairs_out = (rc == B200CD_OK) ? bvh->d_out : nullptr;
    return rc;
}

API int b200cd_self_collide_device(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t shard, uint32_t nshards, uint32_t chunk,
                                   int sorted, const void** d_pairs_out, uint64_t* count_out) {
    if (!ctx || !bvh || !count_out) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    int rc = run_query(ctx, bvh, shard, nshards, chunk, sorted, count_out);
    if (d_pairs_out) *d_pairs_out = (rc == B200CD_OK) ? bvh->d_out : nullptr;
    return rc;
}

API int b200cd_self_collide_shard(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t shard, uint32_t nshards, uint32_t chunk,
                                  uint32_t* pairs_out, uint64_t cap, uint64_t* count_out, int sorted) {
    if (!ctx || !bvh || !count_out || (cap && !pairs_out)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    DeviceGuard g(ctx->device);
    int rc = run_query(ctx, bvh, shard, nshards, chunk, sorted, count_out);
    if (rc != B200CD_OK) return rc;
    if (*count_out > cap) return set_error(ctx, B200CD_E_CAPACITY, "pair buffer holds " + std::to_string(cap) + ", need " + std::to_string(*count_out));
    cudaStream_t s = ctx->stream;
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_D0], s));
    if (*count_out) CD_CUDA(ctx, cudaMemcpyAsync(pairs_out, bvh->d_out, *count_out * sizeof(uint2), cudaMemcpyDeviceToHost, s));
    CD_CUDA(ctx, cudaEventRecord(ctx->ev[EV_D1], s));
    CD_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->stats.ms_download = ev_ms(ctx, EV_D0, EV_D1);
    return B200CD_OK;
}

API int b200cd_self_collide(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t* pairs_out, uint64_t cap, uint64_t* count_out,
                            int sorted) {
    return b200cd_self_collide_shard(ctx, bvh, 0, 1, 0, pairs_out, cap, count_out, sorted);
}

API int b200cd_sort_pairs_device(b200cd_ctx* ctx, void* d_pairs, uint64_t count, uint32_t id_bits) {
    if (!ctx || (count && !d_pairs)) return set_error(ctx, B200CD_E_INVALID, "NULL argument");
    if (id_bits == 0 || id_bits > 32) id_bits = 32;
    if (count < 2) return B200CD_OK;
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    // scratch lives in the context and only grows: this call sits inside every multi-GPU step.
    // Asynchronous: the sorted list is valid in stream order.
    if (ctx->sort_tmp_cap < count) {
        cudaFree(ctx->d_sort_tmp);
        ctx->d_sort_tmp = nullptr;
        ctx->sort_tmp_cap = 0;
        CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_sort_tmp), (count + count / 4) * sizeof(uint2)));
        ctx->sort_tmp_cap = count + count / 4;
    }
    if (!ctx->d_sort_hist) CD_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_sort_hist), radix_hist_words(8) * sizeof(uint32_t)));
    uint2* p = static_cast<uint2*>(d_pairs);
    uint2* q = ctx->d_sort_tmp;
    int rc = sort_pairs_impl(ctx, &p, &q, count, (int)id_bits, &ctx->d_sort_hist, &ctx->d_sort_status, &ctx->sort_status_words, s);
    if (rc == B200CD_OK && p != d_pairs)  // result landed in the temporary: copy back in place
        if (cudaMemcpyAsync(d_pairs, p, count * sizeof(uint2), cudaMemcpyDeviceToDevice, s) != cudaSuccess) rc = B200CD_E_CUDA;
    return rc;
}
