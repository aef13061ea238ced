// common.cuh — internal types shared by the kernels and the C-ABI host code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "b200cd.h"

// traversal stack entries per query: a query first parks every start subtree of its group that it overlaps (at most
// B200CD_MAX_ENTRIES = 64) and then descends (at most one parked sibling per level: depth <= 63 key bits + 30 tie-break
// bits = 93), so 160 covers the worst case of both together; only the entries actually used are ever touched
#define B200CD_MAX_STACK 160
#define B200CD_QUERY_BLOCK 256  // consecutive sorted leaves per traversal block (query chunks are multiples of this)
#define B200CD_QUERY_GROUP 128  // smallest run of consecutive sorted leaves that shares one list of start subtrees
#define B200CD_MAX_ENTRIES 64   // start subtrees recorded per group
#define B200CD_MAX_TRIS_LOG2 30 // at most 2^30 triangles / vertices per mesh

namespace b200cd {

// ---------------------------------------------------------------- device data layout
//
// Traversal node: one child of an internal node. Internal node s is "the split between the
// sorted leaves s and s+1"; its two children sit side by side in pairs[s] (64 B, one aligned
// fetch gives both boxes, both links and the node's leaf range).
//   link >= 0 : child is internal node `link`  (visit pairs[link] next)
//   link <  0 : child is the leaf at sorted position ~link
//   ext       : the far end of the child's leaf range - FIRST leaf for the left child c[0]
//               (its last leaf is s), LAST leaf for the right child c[1] (its first is s+1).
//               In a traversal entry list (collide.cu) ext is always the LAST leaf.
struct __align__(32) Node32 {
    float lo[3];
    float hi[3];
    int32_t link;
    int32_t ext;
};
struct __align__(64) NodePair {
    Node32 c[2];
};

// Everything the narrow phase needs about one triangle, at its sorted position:
// vertex coordinates (copies, so no indirection), vertex indices for the
// shared-vertex filter (reference triangle.cuh:18-30) and the triangle ID.
struct __align__(64) LeafRec {
    float v[9];       // v0.xyz v1.xyz v2.xyz in the triangle's own vIdx order
    uint32_t vi[3];
    uint32_t id;
    uint32_t pad[3];
};
static_assert(sizeof(Node32) == 32 && sizeof(NodePair) == 64 && sizeof(LeafRec) == 64, "layout");

// Quantised traversal node (round 2): the same two children in HALF the bytes - one 32-byte sector, one LDG.256 per
// node visit instead of two (the traversal is bound by the L1 data pipe: 92 % of its peak wavefront rate, profiles/
// r02_ncu_broad_l1.md). c[k] = {x, y, z, link}; an axis word holds the child's box on a 15-bit grid over the Morton box
// (b200cd_params): lo cell in bits 0-14, (32767 - hi cell) in bits 16-30, bits 15 and 31 zero. Cells are CONSERVATIVE:
// lo = floor(u), hi = floor(u) + 1 with u = (x - origin) * scale, clamped to [0, 32766] / [1, 32767]; rounding is
// monotone, so a.lo < b.hi in floats implies cell_lo(a) < cell_hi(b) - a quantised test never misses an overlap, it
// only admits a few more candidates, and the narrow phase re-tests the exact boxes (box.cuh:40-43) first.
// With the query packed as ((hi | 0x8000) | ((32767 - lo) | 0x8000) << 16) - 0x00010001 per axis, BOTH strict
// comparisons of an axis are the two guard bits of ONE subtraction: t = query - node, bit 15 = node.lo < query.hi,
// bit 31 = query.lo < node.hi (neither half can borrow), and the three axes AND together.
struct __align__(32) QNodePair {
    uint4 c[2];
};
struct QFrame {
    float o[3], s[3];  // cell = floor((x - o) * s)
};
constexpr uint32_t Q_CELLS = 32767u;
constexpr uint32_t Q_GUARD = 0x80008000u;
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t qcell_lo(float x, float o, float s) {
    const float u = floorf((x - o) * s);
    return (uint32_t)(int)fminf(fmaxf(u, 0.f), (float)(Q_CELLS - 1u));
}
__device__ __forceinline__ uint32_t qcell_hi(float x, float o, float s) {
    const float u = floorf((x - o) * s) + 1.f;
    return (uint32_t)(int)fminf(fmaxf(u, 1.f), (float)Q_CELLS);
}
__device__ __forceinline__ uint32_t qnode_word(float lo, float hi, float o, float s) {
    return qcell_lo(lo, o, s) | ((Q_CELLS - qcell_hi(hi, o, s)) << 16);
}
__device__ __forceinline__ uint32_t qquery_word(float lo, float hi, float o, float s) {
    return ((qcell_hi(hi, o, s) | 0x8000u) | (((Q_CELLS - qcell_lo(lo, o, s)) | 0x8000u) << 16)) - 0x00010001u;
}
__device__ __forceinline__ bool qoverlap(uint32_t qx, uint32_t qy, uint32_t qz, uint32_t nx, uint32_t ny, uint32_t nz) {
    return (((qx - nx) & (qy - ny) & (qz - nz)) & Q_GUARD) == Q_GUARD;
}
// a child as stored in a NodePair half (a = lo.xyz hi.x, b = hi.yz link ext) -> its quantised half
__device__ __forceinline__ uint4 qnode_half(const float4& a, const float4& b, const QFrame& f) {
    return make_uint4(qnode_word(a.x, a.w, f.o[0], f.s[0]), qnode_word(a.y, b.x, f.o[1], f.s[1]),
                      qnode_word(a.z, b.y, f.o[2], f.s[2]), __float_as_uint(b.z));
}
#endif
inline QFrame make_qframe(const b200cd_params& p) {
    QFrame f;
    for (int k = 0; k < 3; ++k) {
        f.o[k] = (float)p.morton_origin[k];
        const double e = p.morton_extent[k];
        f.s[k] = (e > 0.0 && e < 1e300) ? (float)((double)Q_CELLS / e) : 0.f;
        if (!(f.s[k] > 0.f) || !(f.s[k] < 3.0e38f) || !(f.o[k] == f.o[k])) { f.o[k] = 0.f; f.s[k] = 0.f; }  // degenerate box: one cell
    }
    return f;
}

// ---------------------------------------------------------------- host-side objects

}  // namespace b200cd

struct b200cd_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // b200cd_mesh_update_async: uploads that overlap the build / query of another mesh
    cudaStream_t push_stream[4] = {};    // b200cd_mesh_update_slice_async: peer pushes (several copy engines at once)
    cudaEvent_t push_ev[4] = {};         // chunk c's H2D has landed / push stream p has drained
    int sm_count = 148;
    b200cd_stats stats{};
    std::string last_error;
    cudaEvent_t ev[20]{};
    // scratch shared by builds/queries on this context
    uint32_t* d_scalars = nullptr;   // small device scratch (counters, flags, bbox)
    uint32_t* h_scalars = nullptr;   // pinned mirror
    // scratch of b200cd_sort_pairs_device (grow-only)
    uint2* d_sort_tmp = nullptr;  uint64_t sort_tmp_cap = 0;
    uint32_t* d_sort_hist = nullptr;
    uint32_t* d_sort_status = nullptr;  uint64_t sort_status_words = 0;
    // scratch of b200cd_unique_triangles_device (grow-only): ID bitmap, per-block counts, result
    uint32_t* d_uniq_bits = nullptr;  uint64_t uniq_words = 0;
    uint32_t* d_uniq_sums = nullptr;  uint64_t uniq_blocks = 0;
    uint32_t* d_uniq_out = nullptr;   uint64_t uniq_out_cap = 0;
};

struct b200cd_mesh {
    b200cd_ctx* ctx = nullptr;
    uint32_t nverts = 0, ntris = 0;
    float4* d_verts = nullptr;   // nverts float4 (xyz, w = 0)
    uint32_t* d_idx = nullptr;   // ntris * 3
    float* d_stage = nullptr;    // nverts * 3: H2D landing area of b200cd_mesh_update (kept between frames)
    // asynchronous upload (b200cd_mesh_update_async / b200cd_mesh_wait)
    cudaEvent_t ev_ready = nullptr;     // recorded on the copy stream after the upload's last operation
    cudaEvent_t ev_consumed = nullptr;  // recorded on the work stream after the last kernel that read this mesh
    uint32_t* d_async_flag = nullptr;   // index check of the pending upload (device word + pinned mirror)
    uint32_t* h_async_flag = nullptr;
    bool pending = false;               // an asynchronous upload has been enqueued and not yet waited for
    bool consumed_valid = false;
    // multi-GPU: the same mesh object on the other ranks' GPUs (CUDA-IPC mappings), b200cd_mesh_set_peers
    uint32_t npeers = 0, my_rank = 0;
    float4* peer_verts[16] = {};
    uint32_t* peer_idx[16] = {};
};

namespace b200cd {
struct PeerTable;
}

struct b200cd_bvh {
    b200cd_ctx* ctx = nullptr;
    uint32_t n = 0;         // triangles (leaves) currently in the tree
    uint32_t cap = 0;       // leaves the buffers were sized for (== n except for partitioned builds)
    uint64_t ghost_cap = 0; // ghost leaf records that fit after the local leaves in d_leaves
    float* d_block_boxes = nullptr;           // [cap / 256 + 1][8] union box of every 256-leaf block (partitioned builds)
    b200cd::LeafRec* d_ghost_out = nullptr;   // [peers][ghost_out_cap] outgoing ghost lists
    uint32_t* d_cut_scratch = nullptr;        // coarse-box reduction scratch (256*6+1 words)
    b200cd::PeerTable* d_peers = nullptr;     // peer-memory destinations (b200cd_bvh_set_peers)
    unsigned long long* d_ghost_in_count = nullptr;  // ghosts appended to MY ghost records by the peers (and by me)
    uint32_t* d_ghost_list = nullptr;         // ghost_list_words(cap, peers): (block, peer, box mask) items of the ghost selection's pre-filter
    uint32_t ghost_list_cap = 0;
    uint64_t ghost_out_cap = 0;
    uint32_t max_peers = 0;  // outgoing ghost lists allocated (b200cd_bvh_alloc_partial)
    uint32_t nranks = 0;     // ranks given to b200cd_bvh_set_peers (0: not set)
    uint32_t nverts = 0;
    bool built = false;
    bool unshared_verts = false;  // mesh with (mostly) unshared vertices, V >= 1.5 N: a triangle soup (set by the build)
    b200cd_params params{};
    // sort buffers (ping-pong); sorted result is in d_keys[cur] / d_ids[cur]
    uint64_t* d_keys[2] = {nullptr, nullptr};
    uint32_t* d_ids[2] = {nullptr, nullptr};
    int cur = 0;
    uint32_t* d_hist = nullptr;        // radix histograms / digit bases
    uint32_t* d_tile_status = nullptr; // decoupled look-back words
    uint64_t tile_status_words = 0;
    // hybrid sort (radix_sort.cu): radix passes over the key bits >= 64 - 8 * sort_high only, per-run fix-up below
    uint32_t* d_fix = nullptr;         // [0] fallback ran, [1] longest run of equal high bits, [2] items in runs >= 2
    uint32_t* h_fix = nullptr;         // pinned copy, read by the NEXT build to adapt sort_high
    cudaEvent_t ev_fix = nullptr;      // the copy above has landed
    bool fix_pending = false;
    int sort_high = 5;                 // digits sorted by radix passes (8 = plain full sort)
    int sort_top = 0;                  // significant key bits seen by the previous build (0 = unknown: all 63)
    bool sort_locked = false;          // a longer prefix was needed once: never try a shorter one again
    int sort_slow_streak = 0;          // consecutive builds whose measured fix-up cost more than 1.3 radix passes
    bool sort_trial = false;           // the last build tried one digit more because of that: its timing decides
    bool sort_time_frozen = false;     // the trial did not pay off: stop escalating on time
    float sort_fix_before = 0.f;       // the fix-up time that triggered the trial
    cudaEvent_t ev_sort[4] = {};       // around the last radix pass and around the fix-up of the previous hybrid sort (timed)
    // hierarchy
    uint32_t* d_flags = nullptr;       // n-1 arrival counters of the splits merged through global memory
    void* d_build_scratch = nullptr;   // pending-subtree list of the tree build (lbvh.cu)
    b200cd::NodePair* d_pairs = nullptr;  // n-1
    b200cd::QNodePair* d_qpairs = nullptr;  // n-1: the same nodes on the 15-bit grid (what the traversal reads)
    float* d_qframe = nullptr;         // 8 floats: QFrame of d_qpairs
    bool root_valid = false;           // d_root_box holds the root box of a build of THIS handle (frame of the next build's grid)
    bool qvalid = false;               // the last build wrote d_qpairs (soups; B200CD_BROAD_QUANT): the traversal walks them
    b200cd::LeafRec* d_leaves = nullptr;  // n
    b200cd::LeafRec* d_recs = nullptr;    // n, face order: written by K1, moved into sorted order by the tree build (full builds only)
    float* d_root_box = nullptr;       // 6 floats + [6] = index of the root node (int)
    // query
    b200cd::Node32* d_entries = nullptr;  // [blocks][B200CD_MAX_ENTRIES] traversal start subtrees
    uint32_t* d_entry_count = nullptr;    // [blocks]
    uint64_t entry_blocks = 0;
    uint2* d_cand = nullptr;  uint64_t cand_cap = 0;
    uint2* d_out = nullptr;   uint64_t out_cap = 0;
    uint2* d_out_tmp = nullptr; uint64_t out_tmp_cap = 0;
    uint64_t npairs = 0;     // pairs currently in d_out (last local query + appended ghost queries)
    uint32_t id_space = 0;   // triangle ids are < id_space (mesh size for a partitioned build, else n)
    unsigned long long* d_counters = nullptr;  // [0] candidates, [1] pairs, [2] error flags
    unsigned long long* h_counters = nullptr;  // pinned
};

namespace b200cd {

// ---------------------------------------------------------------- 256-bit global accesses (sm_100 LDG/STG.256)
// A 32-byte Node32 / half a LeafRec moves in ONE instruction that covers a whole 32-byte DRAM
// sector. Two 16-byte stores per thread instead leave every sector half written per warp
// instruction, and the B200 memory system then reads the sector back before writing it
// (measured: 2.9 GB of extra DRAM reads per 16 M-triangle build, profiles/r01_*).
#ifdef __CUDACC__
#define B200CD_V8_OUT(a, b) "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
#define B200CD_V8_IN(a, b) "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
__device__ __forceinline__ void ld256_nc(const void* p, float4& a, float4& b) {  // read-only path
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : B200CD_V8_OUT(a, b) : "l"(p));
}
__device__ __forceinline__ void ld256_cg(const void* p, float4& a, float4& b) {  // through L2 (data another SM just wrote)
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : B200CD_V8_OUT(a, b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(void* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), B200CD_V8_IN(a, b) : "memory");
}
__device__ __forceinline__ void st256_cg(void* p, const float4& a, const float4& b) {
    asm volatile("st.global.cg.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), B200CD_V8_IN(a, b) : "memory");
}
__device__ __forceinline__ void st256_cs(void* p, const float4& a, const float4& b) {  // streaming: written once, read later by another kernel
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), B200CD_V8_IN(a, b) : "memory");
}
#endif

// every kernel launch of this library is counted (b200cd_stats::kernel_launches)
extern unsigned long long g_kernel_launches;
inline void count_launch(unsigned n = 1) { g_kernel_launches += n; }

// Kernel timeline for runs that cannot go under a profiler (several ranks): with B200CD_TRACE set in the environment
// every trace_mark records a CUDA event behind the launches it names; b200cd_trace_dump writes "name,ms since the
// previous mark" rows (device time, gaps included). Off (one predictable branch per call) otherwise.
void trace_mark(const char* name, cudaStream_t s);

inline int set_error(b200cd_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

#define CD_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return b200cd::set_error((ctx), e__ == cudaErrorMemoryAllocation ? B200CD_E_NOMEM      \
                                                                              : B200CD_E_CUDA,    \
                                     std::string(#expr) + ": " + cudaGetErrorString(e__));         \
    } while (0)

// obj_parse.cu: multi-threaded host parser of the reference's OBJ dialect; returns a B200CD_* status
}  // namespace b200cd
#include <vector>
namespace b200cd {
int obj_parse(const char* path, std::vector<float>& xyz, std::vector<uint32_t>& idx, std::string& err);

// kernels' host launchers (each enqueues on `s`, no synchronisation)
// morton.cu
void launch_expand_verts(const float* d_xyz, float4* d_verts, uint32_t nverts, cudaStream_t s);
void launch_check_idx(const uint32_t* d_idx, uint32_t ntris, uint32_t nverts, uint32_t* d_flag, int sms, cudaStream_t s);
void launch_check_idx_accumulate(const uint32_t* d_idx, uint32_t ntris, uint32_t nverts, uint32_t* d_flag, int sms, cudaStream_t s);
void launch_bbox(const float4* d_verts, uint32_t nverts, uint32_t* d_bbox6 /*ordered-uint min3,max3*/, int sms, cudaStream_t s);
// d_recs (optional): also write every triangle's LeafRec, in face order (slice-relative like d_keys)
// The digit histograms of the radix passes that WILL run on the keys K1 is writing (radix_hist_plan): K1 holds every key
// in a register anyway, so it counts the digits itself and the sort's separate histogram pass (one more read of the
// keys) is skipped (radix_sort hist_done). hist: [npass][256] words, zeroed by the caller.
struct RadixHistPlan {
    int npass;
    int shift[8];
    uint32_t mask[8];
};
void launch_morton(const float4* d_verts, const uint32_t* d_idx, uint32_t first, uint32_t n, const b200cd_params& p,
                   const uint32_t* d_bbox6_or_null, uint64_t* d_keys, cudaStream_t s, LeafRec* d_recs = nullptr,
                   const RadixHistPlan* hist_plan = nullptr, uint32_t* d_hist = nullptr, int sms = 148);
// radix_sort.cu
struct RadixPass { int shift; int bits; };
// sorts n (key,value) items; values may be null (keys only); if iota_values the
// first pass generates value = index. Result lands in buffer index returned.
// high_passes > 0 (with d_fix, 4 device words): hybrid sort - only the top `high_passes` digits are sorted with
// radix passes, the low bits by a per-run fix-up (radix_sort.cu, rs_fixup); needs an EVEN npass.
// d_fix afterwards: [0] the fallback (all passes) ran, [1] longest run of equal high bits, [2] items in runs >= 2,
// [3] significant key bits (highest set bit + 1). top_bits (0 = all): where the window of sorted digits ends.
int radix_sort(uint64_t* keys[2], uint32_t* vals[2], uint32_t n, const RadixPass* passes, int npass,
               bool iota_values, uint32_t* d_hist, uint32_t* d_tile_status, uint64_t tile_status_words,
               int sms, cudaStream_t s, int high_passes = 0, uint32_t* d_fix = nullptr, int top_bits = 0,
               cudaEvent_t* ev4 = nullptr /* hybrid: events recorded around the last radix pass [0,1] and the fix-up [2,3] */,
               bool hist_done = false /* d_hist already holds the digit histograms of radix_hist_plan (and zeroed tickets) */);
// the passes radix_sort will run first for these arguments (hybrid: the window of high digits), for a fused histogram
void radix_hist_plan(const RadixPass* passes, int npass, bool has_values, int high_passes, bool has_fix, int top_bits,
                     RadixHistPlan* out);
// stable range partition (multi-GPU): bucket = number of device-resident splitters <= key; needs
// d_hist >= 2*256+1 words and d_tile_status >= radix_tile_status_words(n, 1); counts land in d_hist[256..]
void radix_partition(const uint64_t* keys_in, const uint32_t* vals_in, uint32_t iota_base, uint64_t* keys_out,
                     uint32_t* vals_out, uint32_t n, const uint64_t* d_splitters, int nsplit, uint32_t* d_hist,
                     uint32_t* d_tile_status, int sms, cudaStream_t s);
#define RS_MAX_SPLIT_P1 16  // ranks a partitioned build can address
// Peer-memory destinations of a partitioned build (own memory for the rank itself, cudaIpc-mapped
// memory of the other ranks' GPUs otherwise; kernels store / atomically append through them over NVLink):
// bucket d of the range partition goes into rank d's (key, id) receive buffers, ghosts for rank d are
// appended to its ghost records under its counter.
struct PeerTable {
    uint64_t* keys[RS_MAX_SPLIT_P1];
    uint32_t* ids[RS_MAX_SPLIT_P1];
    LeafRec* ghosts[RS_MAX_SPLIT_P1];
    unsigned long long* ghost_count[RS_MAX_SPLIT_P1];
    unsigned long long ghost_cap;
};
void radix_partition_counts(const uint64_t* keys_in, uint32_t n, const uint64_t* d_splitters, int nsplit, uint32_t* d_counts,
                            int sms, cudaStream_t s);
void radix_partition_to_peers(const uint64_t* keys_in, uint32_t iota_base, uint32_t n, const uint64_t* d_splitters, int nsplit,
                              const PeerTable* d_peers, const uint32_t* d_recv_offsets, uint32_t* d_hist,
                              uint32_t* d_tile_status, cudaStream_t s);
// lists of at most 8192 pairs: one block, bitonic network in shared memory, in place; false = too long (use radix_sort)
bool small_pair_sort(uint2* d_pairs, uint64_t count, cudaStream_t s);
uint64_t radix_tile_status_words(uint32_t n, int npass);
int radix_digit_bits();  // digit width of one pass (status / histogram rows hold 1 << bits words)
uint32_t radix_hist_words(int npass);
// lbvh.cu
uint32_t build_tree_pending_capacity(uint32_t n);
uint64_t build_tree_scratch_bytes(uint32_t n);
// d_recs (optional): face-ordered leaf records from launch_morton - the leaves are then copied from there
// (one 64-byte gather per leaf) instead of being assembled from d_idx / d_verts
void launch_build_tree(const float4* d_verts, const uint32_t* d_idx, const uint32_t* d_sorted_ids, const uint64_t* d_keys,
                       uint32_t n, uint32_t* d_flags, NodePair* d_pairs, LeafRec* d_leaves, float* d_root_box,
                       void* d_scratch, cudaStream_t s, const LeafRec* d_recs = nullptr, float* d_block_boxes = nullptr,
                       QNodePair* d_qpairs = nullptr, float* d_qframe = nullptr, const QFrame* frame = nullptr,
                       bool frame_from_root = false /* d_root_box holds the previous build's root box: lay the grid over it */);
// d_scratch: 2 * (2n-1) words
void launch_export_nodes(const NodePair* d_pairs, const float* d_root_box, uint32_t n, uint32_t* d_scratch,
                         b200cd_node32* d_nodes_out, cudaStream_t s);
void launch_validate(const NodePair* d_pairs, const LeafRec* d_leaves, const float* d_root_box, const uint64_t* d_keys,
                     uint32_t n, uint32_t nverts, uint32_t* d_scratch, uint32_t* d_checks9, cudaStream_t s);
// partition.cu (partitioned multi-GPU build)
void launch_partition_plan(const uint32_t* d_ghist, const uint32_t* d_lhist, int shift, int world, uint64_t* d_splitters,
                           int32_t* d_counts, cudaStream_t s);
void launch_key_hist16(const uint64_t* d_keys, uint32_t n, int shift, uint32_t* d_hist65536, int sms, cudaStream_t s);
// dist.cu's range plan: splitters, [source][owner] counts, receive offsets and range sizes from ALL ranks' histograms
constexpr int DIST_HIST_BLOCK_BINS = 1024;
constexpr int DIST_HIST_BLOCKS = 65536 / DIST_HIST_BLOCK_BINS;
struct DistPlan {
    uint64_t splitters[RS_MAX_SPLIT_P1];                 // world-1 used, ascending
    uint32_t recv_off[RS_MAX_SPLIT_P1];                  // where my segment starts in every owner's receive buffers
    uint32_t totals[RS_MAX_SPLIT_P1];                    // triangles each rank owns
    uint32_t counts[RS_MAX_SPLIT_P1][RS_MAX_SPLIT_P1];   // [source rank][owner rank]
};
// d_hists: [world][65536]; d_ghist: 65536 words; d_part: scratch of RS_MAX_SPLIT_P1 * DIST_HIST_BLOCKS + 1024 words
void launch_dist_plan(const uint32_t* d_hists, int world, int rank, int shift, uint32_t* d_ghist, uint32_t* d_part,
                      DistPlan* d_plan, cudaStream_t s);
// d_scratch: K*6 + 1 words
void launch_chunk_boxes(const NodePair* d_pairs, const float* d_root_box, uint32_t n, uint32_t K, uint32_t* d_scratch,
                        float* d_boxes, cudaStream_t s);
int ghost_max_k();
void launch_ghosts(const LeafRec* d_leaves, uint32_t n, const float* d_peer_boxes, uint32_t npeers, uint32_t K,
                   uint32_t peer_mask, LeafRec* d_ghosts, uint64_t cap_per_peer, unsigned long long* d_counts,
                   float* d_overall /* scratch: 6 * npeers floats */, const float* d_block_boxes, cudaStream_t s);
// same selection, but the records are appended straight into the peers' ghost buffers (remote atomics + stores)
void launch_ghosts_to_peers(const LeafRec* d_leaves, uint32_t n, const float* d_peer_boxes, uint32_t npeers, uint32_t K,
                            uint32_t peer_mask, const PeerTable* d_peers, float* d_overall, const float* d_block_boxes,
                            cudaStream_t s, uint32_t* d_list = nullptr /* scratch of ghost_list_words(): enables the block pre-filter */,
                            uint32_t list_cap = 0 /* (block, peer) items the scratch holds */, int sms = 148);
inline uint32_t ghost_list_items(uint32_t n, uint32_t peers) { return (n / 256 + 1) * (peers ? peers : 1); }  // worst case: all pairs
inline uint64_t ghost_list_words(uint32_t n, uint32_t peers) { return 4 + 10ull * ghost_list_items(n, peers); }
// unique.cu: sorted set of the triangle IDs in a pair list (reference main.cu:33-45). d_bits: ceil(id_space / 32) words,
// d_sums: ceil(words / 1024) + 1 words; *d_sums_total (last word of d_sums) receives the count; d_out: ids, ascending
void launch_unique_mark(const uint2* d_pairs, uint64_t count, uint32_t id_space, uint32_t* d_bits, cudaStream_t s);
void launch_unique_count(const uint32_t* d_bits, uint64_t words, uint32_t* d_sums, cudaStream_t s);
void launch_unique_emit(const uint32_t* d_bits, uint64_t words, const uint32_t* d_sums, uint32_t* d_out, uint64_t out_cap,
                        cudaStream_t s);
// collide.cu
// foreign != 0: the queries are the nquery ghost records stored at leaves[ghost_base ...]; they start at the
// root and are tested against every local leaf (no "only later positions" rule)
void launch_broad(const NodePair* d_pairs, const LeafRec* d_leaves, const float* d_root_box, uint32_t n, uint32_t shard,
                  uint32_t nshards, uint32_t chunk, uint32_t nquery, int foreign, uint32_t ghost_base, Node32* d_entries, uint32_t* d_entry_count, uint2* d_cand,
                  uint64_t cand_cap, unsigned long long* d_counters, cudaStream_t s,
                  const unsigned long long* d_nquery = nullptr /* foreign only: device-side query count (nquery = cap) */, int sms = 148,
                  bool shared_vertices = false /* a mesh (V < 1.5 N): drop vertex-sharing candidates in the traversal */,
                  const QNodePair* d_qpairs = nullptr, const float* d_qframe = nullptr /* quantised nodes + their frame */);
bool broad_uses_quantised_nodes(bool shared_vertices);  // whether a build should write d_qpairs for the traversal
void launch_narrow(const LeafRec* d_leaves, const uint2* d_cand, uint64_t cand_cap, uint2* d_out, uint64_t out_cap,
                   unsigned long long* d_counters, int sms, cudaStream_t s, bool unshared_vertices = false);

}  // namespace b200cd
