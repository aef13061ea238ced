/*
 * b200cd.h — C ABI of libb200cd.so: B200-native (sm_100a) triangle-mesh
 * self-collision detection. Drop-in boundary for the one hot path of
 * Asichurter/GPU-Computing-Course's CollisionDetection project:
 *
 *     OBJ / arrays in -> LBVH build -> self-collision query -> colliding pair list out
 *
 * The reference has no FFI/plugin interface; its "entry-point surface" is the call
 * sequence inside main() (reference CollisionDetection/main.cu:47-174). Each entry
 * point below names the reference call it replaces. All citations are relative to
 * /root/reference/CollisionDetection/.
 *
 * Conventions
 *   - plain C, opaque handles, every function returns an int status (0 = OK);
 *     nothing calls exit() (reference: common/book.h:21-31, load_obj.h:34,60,73)
 *     and nothing throws across the boundary;
 *   - one context = one GPU = one host thread at a time (externally synchronised);
 *     multi-GPU = one process (rank) per GPU, each with its own context, sharding
 *     the query with b200cd_self_collide_shard;
 *   - the library owns device memory behind handles; the caller owns host buffers;
 *   - there is NO CPU fallback: without a CUDA device b200cd_create fails.
 */
#ifndef B200CD_H
#define B200CD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: + b200cd_mesh_update_async / _wait / _update_slice_async / _ipc_export / _set_peers, b200cd_partition_plan_device
 * 3: + b200cd_dist_* (the multi-GPU step in C++), b200cd_unique_triangles[_device], b200cd_nccl_unique_id, B200CD_E_PEER
 * (additions only: every earlier entry point keeps its signature and meaning) */
#define B200CD_ABI_VERSION 3

typedef struct b200cd_ctx b200cd_ctx;
typedef struct b200cd_mesh b200cd_mesh;
typedef struct b200cd_bvh b200cd_bvh;
typedef struct b200cd_dist b200cd_dist;

enum {
    B200CD_OK = 0,
    B200CD_E_INVALID = 1,   /* bad argument */
    B200CD_E_CUDA = 2,      /* CUDA runtime error, see b200cd_last_error */
    B200CD_E_NOMEM = 3,     /* host or device allocation failed */
    B200CD_E_IO = 4,        /* file could not be opened / read */
    B200CD_E_PARSE = 5,     /* OBJ line not in the accepted dialect */
    B200CD_E_CAPACITY = 6,  /* caller's pair buffer too small; *count_out holds the true count */
    B200CD_E_DEPTH = 7,     /* traversal stack exhausted (tree deeper than B200CD_MAX_STACK) */
    B200CD_E_NODEVICE = 8,  /* no usable sm_100 device */
    B200CD_E_TOOBIG = 9,    /* mesh exceeds 2^30 triangles or vertices */
    B200CD_E_PEER = 10      /* multi-GPU step: a barrier between the ranks timed out (a peer failed) */
};

/* Morton normalisation and key width.
 * Reference: morton.h:43-58 hard-codes origin/extent for the flag mesh and
 * morton.h:70-89 builds 63-bit codes (20 bits per axis used); morton.h:31-40 holds
 * the unused 30-bit variant. b200cd_default_params() yields exactly those
 * constants; auto_box != 0 replaces them with the mesh's own bounding box. */
typedef struct b200cd_params {
    double morton_origin[3];
    double morton_extent[3];
    int32_t key_bits;            /* 63 or 30 */
    int32_t auto_box;            /* 0: use origin/extent above; 1: mesh bounding box */
    uint64_t pair_capacity_hint; /* expected number of colliding pairs (0 = let the library guess) */
} b200cd_params;

/* One BVH node as exported by b200cd_bvh_download, in the reference's Karras
 * numbering (bvh.cuh:161-198): nodes[0 .. n-2] are the internal nodes with their
 * Karras index (root = 0), nodes[n-1+j] is the leaf of sorted position j.
 * left/right are such unified node numbers (-1 for leaves). 32 bytes: the
 * reference's 112-byte Node (bvh.cuh:25-43) holds six doubles that are always
 * exact copies of fp32 vertex coordinates (load_obj.h:38,50-52), so float
 * storage is lossless. */
typedef struct b200cd_node32 {
    float lo[3];
    float hi[3];
    int32_t left;
    int32_t right;
} b200cd_node32;

/* Per-stage device times of the most recent build / query on a context, from CUDA
 * events on the context's stream (reference: printElapsedTime, main.cu:19-24). */
typedef struct b200cd_stats {
    float ms_upload;      /* H2D + vertex expansion (mesh_from_arrays / load_obj) */
    float ms_morton;      /* K1: centroid + Morton keys (+ bounding box when auto_box) */
    float ms_sort;        /* K2: onesweep radix sort of (key, id) */
    float ms_hierarchy;   /* K3: Karras hierarchy */
    float ms_refit;       /* K4: leaf records + bottom-up AABB refit */
    float ms_build;       /* K1..K4 end to end */
    float ms_traverse;    /* K5: broad phase (BVH traversal -> candidate pairs) */
    float ms_narrow;      /* K6: fp64 triangle-triangle tests -> pairs */
    float ms_pair_sort;   /* optional lexicographic sort of the pair list */
    float ms_query;       /* K5 + K6 (+ sort) end to end */
    float ms_download;    /* D2H of the pair list */
    uint32_t ntris, nverts;
    uint64_t candidates;  /* leaf pairs handed to the narrow phase: the AABB-overlapping ones (collision.cuh:36), on a triangle soup a
                             few per cent more - its traversal walks 15-bit quantised boxes and the narrow phase re-tests the exact ones */
    uint64_t pairs;       /* colliding pairs found */
    uint32_t sort_passes; /* radix passes actually run */
    uint32_t query_retries; /* times a stage was re-run after growing a buffer */
    uint64_t kernel_launches; /* cumulative count of this library's kernel launches on the context */
    uint64_t nodes_visited;   /* K5: internal nodes fetched, summed over queries (nodes-visited/s = this / ms_traverse) */
    uint64_t warp_steps;      /* K5: traversal loop iterations summed over warps (x32 = lane slots) */
    uint64_t start_entries;   /* K5: start subtrees kept per warp, summed over warps */
} b200cd_stats;

/* Structural self-checks, the counters the reference prints on every run
 * (check.cuh:64-96, main.cu:113-136): all must be 0 except null_parent_internal,
 * which is 1 (the root). */
typedef struct b200cd_checks {
    uint32_t null_parent_internal; /* check.cuh:74 */
    uint32_t wrong_bound_count;    /* check.cuh:73  (visit counter != 2) */
    uint32_t null_child;           /* check.cuh:75-76 */
    uint32_t uninit_box_internal;  /* check.cuh:77 */
    uint32_t null_parent_leaf;     /* check.cuh:91 */
    uint32_t bad_triangle;         /* check.cuh:92, :29-50 (vertex index out of range) */
    uint32_t uninit_box_leaf;      /* check.cuh:93 */
    uint32_t unsorted_keys;        /* load_obj.h:109-115 (strictly increasing; ties count) */
    uint32_t box_not_enclosing;    /* ours: a parent box that does not contain its children */
} b200cd_checks;

/* ---- context ------------------------------------------------------------ */

/* Replaces the implicit "device 0, default stream" of main.cu. */
int b200cd_create(int device, b200cd_ctx** out);
int b200cd_destroy(b200cd_ctx* ctx);
/* Run all of the context's work on a caller-owned cudaStream_t (e.g. the stream a benchmark
 * records its CUDA events on, or the one NCCL transfers are ordered against). NULL is the legacy
 * default stream (where the reference runs everything, main.cu:92-142);
 * B200CD_PRIVATE_STREAM restores the context's own non-blocking stream. */
#define B200CD_PRIVATE_STREAM ((void*)(intptr_t)-1)
int b200cd_set_stream(b200cd_ctx* ctx, void* cuda_stream);
int b200cd_synchronize(b200cd_ctx* ctx);
int b200cd_get_stats(const b200cd_ctx* ctx, b200cd_stats* out);
const char* b200cd_strerror(int status);
const char* b200cd_last_error(const b200cd_ctx* ctx);
int b200cd_abi_version(void);

/* Debug aid, not a reference entry point: with B200CD_TRACE set in the environment the library records a CUDA event
 * behind its kernel launches; this appends "name,milliseconds since the previous mark" rows (device time) to `path`
 * and clears the list. For runs that cannot go under a profiler (several ranks). */
int b200cd_trace_dump(b200cd_ctx* ctx, const char* path);
void b200cd_trace_enable(int on); /* switch the recording on / off at run time (initial state: B200CD_TRACE set or not) */

void b200cd_default_params(b200cd_params* p);

/* For callers without the CUDA runtime of their own (examples/dist_main.cpp): the number of usable GPUs, and a
 * copy of library-owned DEVICE results (b200cd_self_collide_device, b200cd_dist_step, b200cd_unique_triangles_device)
 * to host memory, ordered behind the context's stream and complete on return. */
int b200cd_device_count(int* count_out);
int b200cd_copy_to_host(b200cd_ctx* ctx, void* dst, const void* d_src, uint64_t bytes);

/* Page-locked host buffers for callers that want full-rate H2D/D2H. */
int b200cd_host_alloc(void** out, uint64_t bytes);
int b200cd_host_free(void* p);

/* ---- mesh in ------------------------------------------------------------ */

/* Replaces loadObj (load_obj.h:24-103) + the H2D copies of main.cu:78-88.
 * Same dialect and quirks: only "v %f %f %f" and "f %d/%d %d/%d %d/%d" lines are
 * read (load_obj.h:50,68), indices are 1-based (load_obj.h:81-83), triangle ID =
 * face order (load_obj.h:94), a last line without '\n' is dropped (load_obj.h:41).
 * Malformed lines and out-of-range / forward vertex references return
 * B200CD_E_PARSE instead of exit()/undefined behaviour. */
int b200cd_mesh_load_obj(b200cd_ctx* ctx, const char* path, b200cd_mesh** out);

/* The parser behind b200cd_mesh_load_obj on its own (host only, multi-threaded, no GPU needed):
 * *xyz_out = nverts*3 floats, *idx_out = ntris*3 0-based indices, both malloc'ed - release with
 * b200cd_host_array_free. err_out (optional) receives "line N: ..." on B200CD_E_PARSE / B200CD_E_IO. */
int b200cd_obj_parse_host(const char* path, float** xyz_out, uint32_t* nverts_out, uint32_t** idx_out, uint32_t* ntris_out,
                          char* err_out, uint64_t err_len);
void b200cd_host_array_free(void* p);

/* Array twin of the above: xyz = nverts*3 floats, tri_idx = ntris*3 0-based
 * indices, triangle ID = array order. Host pointers. */
int b200cd_mesh_from_arrays(b200cd_ctx* ctx, const float* xyz, uint32_t nverts,
                            const uint32_t* tri_idx, uint32_t ntris, b200cd_mesh** out);
/* Same, but the two arrays are already in device memory on the context's GPU. */
int b200cd_mesh_from_device(b200cd_ctx* ctx, const void* d_xyz, uint32_t nverts,
                            const void* d_tri_idx, uint32_t ntris, b200cd_mesh** out);
/* Overwrite the contents of an existing mesh in place (same nverts / ntris):
 * xyz and/or tri_idx may be NULL to keep that array. Host (on_device = 0) or
 * device pointers. This is the per-frame update of a cloth / flag simulation and
 * the steady-state upload path (no allocation). */
int b200cd_mesh_update(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, const uint32_t* tri_idx,
                       int on_device);
/* Double-buffered frames (a simulation that streams meshes through the library): same as
 * b200cd_mesh_update from HOST pointers, but enqueued on the context's copy stream and returning at
 * once, so the PCIe transfer of frame k+1 overlaps the build + query of frame k running from
 * another mesh object. The host arrays must stay valid until b200cd_mesh_wait(mesh) returns; every
 * other call that takes this mesh fails with B200CD_E_INVALID until then. The copy is ordered
 * after the last build / refit that read this mesh. b200cd_mesh_wait blocks the host until the
 * upload has landed and returns the index check's verdict (B200CD_E_INVALID for an index >= nverts). */
int b200cd_mesh_update_async(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, const uint32_t* tri_idx);
int b200cd_mesh_wait(b200cd_ctx* ctx, b200cd_mesh* mesh);
/* Multi-GPU upload: overwrite only vertices [first_vert, +nverts) and triangles [first_tri, +ntris)
 * from host memory; the ranks then all-gather b200cd_mesh_device_buffers over NVLink (the buffers
 * carry 16 elements of padding so equal chunks of ceil(n / ranks) fit). */
int b200cd_mesh_update_slice(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, uint32_t first_vert, uint32_t nverts,
                             const uint32_t* tri_idx, uint32_t first_tri, uint32_t ntris);
/* Several GPUs, double-buffered frames: b200cd_mesh_ipc_export gives the CUDA-IPC handles (2 x 64 bytes) and
 * offsets (2) of the mesh's vertex and index buffers; every other rank maps them (b200cd_ipc_open) and hands the
 * mapped addresses to b200cd_mesh_set_peers (peers[2 * r + {0, 1}] = rank r's vertex / index buffer as seen from
 * this GPU). b200cd_mesh_update_slice_async then uploads this rank's slice over its own PCIe link AND copies it into
 * every peer's mesh with the copy engines over NVLink, all on the copy stream, returning at once;
 * b200cd_mesh_wait(mesh) blocks until this rank's slice has landed everywhere (index check as for
 * b200cd_mesh_update_async), and a collective across the ranks after it means the whole frame is in place.
 * Not before every rank has finished its last build from this mesh object. */
int b200cd_mesh_ipc_export(b200cd_ctx* ctx, b200cd_mesh* mesh, uint8_t* handles_out, uint64_t* offsets_out);
int b200cd_mesh_set_peers(b200cd_ctx* ctx, b200cd_mesh* mesh, uint32_t nranks, uint32_t my_rank, void* const* peers);
int b200cd_mesh_update_slice_async(b200cd_ctx* ctx, b200cd_mesh* mesh, const float* xyz, uint32_t first_vert, uint32_t nverts,
                                   const uint32_t* tri_idx, uint32_t first_tri, uint32_t ntris);
/* float4[nverts + 16] (xyz, w = 0) and uint32[3 * (ntris + 16)] on the context's GPU */
int b200cd_mesh_device_buffers(b200cd_ctx* ctx, b200cd_mesh* mesh, void** d_verts4, void** d_idx);
int b200cd_mesh_info(const b200cd_mesh* mesh, uint32_t* nverts, uint32_t* ntris);
int b200cd_mesh_download(b200cd_ctx* ctx, const b200cd_mesh* mesh, float* xyz, uint32_t* tri_idx);
int b200cd_mesh_destroy(b200cd_mesh* mesh);

/* ---- BVH build ---------------------------------------------------------- */

/* Replaces: centroid + morton3D (load_obj.h:89-91, morton.h:70-89), the host
 * thrust::sort_by_key (load_obj.h:107), fillLeafNodes (bvh.cuh:125, main.cu:92),
 * generateHierarchyParallel (bvh.cuh:146, main.cu:99) and calBoundingBox
 * (bvh.cuh:258, main.cu:107). Asynchronous on the context's stream. */
int b200cd_bvh_build(b200cd_ctx* ctx, const b200cd_mesh* mesh, const b200cd_params* params,
                     b200cd_bvh** out);
/* Rebuild in place, reusing every allocation of an existing BVH (steady state). */
int b200cd_bvh_rebuild(b200cd_ctx* ctx, b200cd_bvh* bvh, const b200cd_mesh* mesh,
                       const b200cd_params* params);
/* Keep topology and order, recompute leaf records and all boxes from the mesh's
 * current vertex positions (K4 only). */
int b200cd_bvh_refit(b200cd_ctx* ctx, b200cd_bvh* bvh, const b200cd_mesh* mesh);
/* Parity hooks: any pointer may be NULL. nodes: 2n-1 entries (see b200cd_node32),
 * sorted_keys / sorted_ids: n entries (load_obj.h:107 result). */
int b200cd_bvh_download(b200cd_ctx* ctx, const b200cd_bvh* bvh, b200cd_node32* nodes,
                        uint64_t* sorted_keys, uint32_t* sorted_ids);
/* The reference's per-run self-checks (check.cuh:29-96) on the device. */
int b200cd_bvh_validate(b200cd_ctx* ctx, const b200cd_bvh* bvh, const b200cd_mesh* mesh,
                        b200cd_checks* out);
int b200cd_bvh_destroy(b200cd_bvh* bvh);

/* ---- self-collision query ------------------------------------------------ */

/* Replaces findCollisions (collision.cuh:73-88, main.cu:142) + the D2H of
 * main.cu:145-146. pairs_out receives *count_out pairs as uint32_t[2], lower
 * triangle ID first (tri_contact.cuh:81); sorted != 0 => lexicographic order.
 * The pair SET equals the reference's: {(a,b): a<b, no shared vertex index
 * (triangle.cuh:18-30), AABBs strictly overlap (box.cuh:40-43),
 * checkTriangleContact(a,b) (tri_contact.cuh:19-78)}. If more than `cap` pairs
 * exist, nothing is written, *count_out holds the count and
 * B200CD_E_CAPACITY is returned (the reference overruns its 500-pair buffer,
 * main.cu:81, collision.cuh:40-42). pairs_out may be NULL with cap = 0 to count. */
int b200cd_self_collide(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t* pairs_out, uint64_t cap,
                        uint64_t* count_out, int sorted);

/* Query-sharded form for one rank of a multi-GPU job: only the query triangles of
 * sorted-leaf chunks c with c % nshards == shard are traversed (chunk = leaves per
 * chunk; 0 = one contiguous slice per shard). The union over all shards is the
 * full pair set; shards are disjoint. */
int b200cd_self_collide_shard(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t shard, uint32_t nshards,
                              uint32_t chunk, uint32_t* pairs_out, uint64_t cap,
                              uint64_t* count_out, int sorted);

/* Device-resident result: *d_pairs_out points at library-owned device memory
 * (valid until the next query on this BVH), *count_out pairs of uint32_t[2]. */
int b200cd_self_collide_device(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t shard, uint32_t nshards,
                               uint32_t chunk, int sorted, const void** d_pairs_out,
                               uint64_t* count_out);

/* Replaces makeAndPrintSet (main.cu:33-45, called at main.cu:154): the second half of the reference's output, the
 * sorted set of the IDs of every triangle that takes part in at least one colliding pair. Computed on the device
 * (bitmap over the ID space, per-word population counts, scan, ordered compaction). The _device form takes any
 * device-resident pair list (uint32_t[2] per pair, IDs < id_space) and returns library-owned device memory valid
 * until the next call on this context; the host form works on the pair list of the last query of `bvh` and follows
 * b200cd_self_collide's capacity protocol (B200CD_E_CAPACITY with the true count; ids_out NULL + cap 0 to count). */
int b200cd_unique_triangles_device(b200cd_ctx* ctx, const void* d_pairs, uint64_t count, uint32_t id_space,
                                   const void** d_ids_out, uint64_t* count_out);
int b200cd_unique_triangles(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t* ids_out, uint64_t cap, uint64_t* count_out);

/* Lexicographic sort of a device-resident pair list (e.g. after gathering the
 * per-rank lists on rank 0). id_bits = number of significant bits in a triangle
 * ID (0 = 32). In place; asynchronous on the context's stream. */
int b200cd_sort_pairs_device(b200cd_ctx* ctx, void* d_pairs, uint64_t count, uint32_t id_bits);

/* An empty BVH of the shape of one built over `ntris` triangles: the receiving side of b200cd_dist_broadcast_bvh
 * (replicated multi-GPU mode: rank 0 builds, the others receive nodes, leaf records and sorted ids over NCCL). */
int b200cd_bvh_alloc_like(b200cd_ctx* ctx, uint32_t ntris, b200cd_bvh** out);

/* ---- partitioned multi-GPU build ------------------------------------------
 * The reference is single-GPU (no collective anywhere). With a replicated BVH
 * every rank repeats the whole build; these entry points let each rank (one process per GPU,
 * NCCL through the caller's torch.distributed) own ONE Morton range of the triangles: it builds
 * and queries only that range and exchanges the thin layer of triangles whose boxes reach into
 * another rank's range ("ghosts"). gpu-computing-course_b200/multigpu.py drives the sequence;
 * the pair set is identical to b200cd_self_collide on one GPU.
 *   1. b200cd_morton_keys_device      keys of this rank's slice of the INPUT triangles
 *   2. b200cd_key_histogram_device    65536-bin histogram of the keys' top bits (all-reduced by the caller -> splitters)
 *   3. b200cd_partition_keys_device   (key, id) bucketed by destination rank; caller all-to-alls them
 *                                     into b200cd_bvh_key_buffers of the receiving rank
 *   4. b200cd_bvh_build_partial       sort + tree over the received triangles
 *   5. b200cd_self_collide_device     pairs inside the rank
 *   6. b200cd_bvh_chunk_boxes_device  K coarse boxes of the rank = a cut through its tree (all-gathered by the caller)
 *   7. b200cd_select_ghosts_device    local leaves overlapping a peer's coarse boxes -> per-peer lists;
 *                                     caller sends them into the peer's b200cd_bvh_ghost_buffer
 *   8. b200cd_collide_ghosts_device   received ghosts against the local tree, pairs appended
 * All pointers named d_* are device memory on the context's GPU; work is enqueued on the context's stream. */
int b200cd_morton_keys_device(b200cd_ctx* ctx, const b200cd_mesh* mesh, const b200cd_params* params, uint32_t first,
                              uint32_t count, void* d_keys_out /* count x u64 */);
int b200cd_key_histogram_device(b200cd_ctx* ctx, const void* d_keys, uint32_t count, int32_t shift,
                                void* d_hist65536 /* 65536 x u32, accumulated into */);
/* BVH with room for `capacity` local triangles, `ghost_capacity` received ghost records and
 * `max_peers` outgoing ghost lists of `ghost_capacity` records each. */
/* The range plan in one launch: from the all-reduced histogram (b200cd_key_histogram_device, summed over the ranks) the
 * world-1 splitters (uint64, ascending; keys >= splitter r-1 belong to rank >= r) that give every rank an equal share,
 * and from this rank's own histogram how many of ITS keys each rank will own (int32[world]). */
int b200cd_partition_plan_device(b200cd_ctx* ctx, const void* d_global_hist65536, const void* d_local_hist65536, int32_t shift,
                                 uint32_t world, void* d_splitters_out, void* d_counts_out);
int b200cd_bvh_alloc_partial(b200cd_ctx* ctx, uint32_t capacity, uint64_t ghost_capacity, uint32_t max_peers,
                             b200cd_bvh** out);
int b200cd_bvh_key_buffers(b200cd_ctx* ctx, b200cd_bvh* bvh, void** d_keys, void** d_ids, uint32_t* capacity);
/* Stable bucketing: bucket = number of splitters <= key (nsplit <= 15 ascending u64 keys in device
 * memory); ids are first_id + position. counts_out[nsplit + 1] (host) receives the bucket sizes. */
int b200cd_partition_keys_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_keys, uint32_t first_id, uint32_t count,
                                 const void* d_splitters, uint32_t nsplit, void* d_keys_out, void* d_ids_out,
                                 uint64_t* counts_out);
/* Sort + tree over the `count` (key, triangle id) items already placed in b200cd_bvh_key_buffers. */
int b200cd_bvh_build_partial(b200cd_ctx* ctx, b200cd_bvh* bvh, const b200cd_mesh* mesh, const b200cd_params* params,
                             uint32_t count);
/* K <= 256 boxes (lo xyz, hi xyz floats) covering every local leaf: the boxes of a cut through the tree
 * (subtrees of at most ~2n/K leaves; unused slots hold lo = +inf, hi = -inf). */
int b200cd_bvh_chunk_boxes_device(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t K, void* d_boxes_out /* K x 6 floats */);
/* d_peer_boxes: [npeers][K][6] floats. For every peer p with bit p set in peer_mask, the local
 * leaves whose box strictly overlaps one of p's boxes are copied (64-byte records) to
 * (*d_ghosts_out) + p * (*stride_out) records; counts_out[npeers] (host). */
int b200cd_select_ghosts_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_peer_boxes, uint32_t npeers, uint32_t K,
                                uint32_t peer_mask, const void** d_ghosts_out, uint64_t* stride_out, uint64_t* counts_out);
/* Where received ghost records go (64 bytes each), valid until the next build of this BVH. */
int b200cd_bvh_ghost_buffer(b200cd_ctx* ctx, b200cd_bvh* bvh, void** d_ptr, uint64_t* capacity);
/* The first nghost records of the ghost buffer are queried against the local tree; keep_pairs != 0
 * appends to the list of the preceding b200cd_self_collide_device call. */
int b200cd_collide_ghosts_device(b200cd_ctx* ctx, b200cd_bvh* bvh, uint64_t nghost, int keep_pairs,
                                 const void** d_pairs_out, uint64_t* count_out);

/* ---- peer memory (NVLink / NVSwitch) variant of steps 3 and 7 -----------------
 * One process per GPU: each rank exports its receive buffers with CUDA IPC once
 * (b200cd_ipc_export -> exchange the bytes -> b200cd_ipc_open -> b200cd_bvh_set_peers). After that
 *   b200cd_partition_to_peers_device    ranks and scatters the (key, id) pairs of the range partition
 *                                       in ONE kernel whose stores land directly in the owning rank's
 *                                       buffers over NVLink (fused partition + all-to-all);
 *   b200cd_send_ghosts_to_peers_device  appends ghost records to the peers' ghost buffers with remote
 *                                       atomics + 256-bit stores (fused selection + exchange).
 * The caller separates the phases with a collective on the same stream (e.g. a 1-word all-reduce). */
/* handles_out: 4 x 64 bytes (key buffer, id buffer, leaf/ghost buffer, ghost counter);
 * offsets_out[4]: byte offset of each buffer inside the mapping b200cd_ipc_open returns */
int b200cd_ipc_export(b200cd_ctx* ctx, b200cd_bvh* bvh, uint8_t* handles_out, uint64_t* offsets_out);
int b200cd_ipc_open(b200cd_ctx* ctx, const uint8_t* handle64, void** d_ptr_out);
int b200cd_ipc_close(b200cd_ctx* ctx, void* d_ptr);
/* peers[4 * r + i]: rank r's buffer i (mapped pointer + offset); the entries of my_rank are ignored */
int b200cd_bvh_set_peers(b200cd_ctx* ctx, b200cd_bvh* bvh, uint32_t nranks, uint32_t my_rank, void* const* peers);
/* u32 counts[nsplit + 1] (device): how many of my keys each rank owns */
int b200cd_partition_counts_device(b200cd_ctx* ctx, const void* d_keys, uint32_t count, const void* d_splitters,
                                   uint32_t nsplit, void* d_counts_out);
/* d_recv_offsets: u32[nsplit + 1] (device): where my segment starts in each rank's receive buffers */
int b200cd_partition_to_peers_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_keys, uint32_t first_id, uint32_t count,
                                     const void* d_splitters, uint32_t nsplit, const void* d_recv_offsets);
int b200cd_send_ghosts_to_peers_device(b200cd_ctx* ctx, b200cd_bvh* bvh, const void* d_peer_boxes, uint32_t npeers, uint32_t K,
                                       uint32_t peer_mask);
int b200cd_ghost_counter_reset(b200cd_ctx* ctx, b200cd_bvh* bvh);
int b200cd_ghost_counter_read(b200cd_ctx* ctx, b200cd_bvh* bvh, uint64_t* count_out);

/* ---- the multi-GPU step as ONE call per rank (one process per GPU) -----------------------------------
 * The partitioned build + query above, driven by the library itself: b200cd_dist_step enqueues every phase on the
 * context's stream and the ranks exchange histograms, (key, id) pairs, coarse boxes, ghost records, verdicts and
 * the final pair lists through CUDA-IPC peer memory over NVLink, ordered by flag barriers over the same mappings
 * (no collective library on the data path, two host reads per step). The pair list rank 0 gets is identical to
 * b200cd_self_collide(sorted = 1) on one GPU.
 *   1. every rank: b200cd_dist_create(ctx, rank, world, ntris, ...)
 *   2. every rank: b200cd_dist_export -> the caller exchanges the blobs between the processes however it likes
 *      (MPI, torch.distributed, files ...) -> b200cd_dist_connect(all blobs in rank order)
 *   3. per frame, every rank: b200cd_dist_step(dist, mesh, params, &d_pairs, &count); only rank 0 receives the
 *      sorted list (library-owned device memory, valid in stream order until the next step)
 *   4. a barrier of the caller's, then b200cd_dist_destroy (peers must not be inside a step any more).
 * Every rank must hold the whole mesh (same ntris, same contents). slack >= 1: room for uneven Morton ranges
 * (capacity per rank = ntris / world * slack + 65536); pair_capacity = room in rank 0's gather buffer
 * (0 = ntris + 65536 pairs; the buffer is peer-mapped and cannot grow: a step that finds more returns
 * B200CD_E_CAPACITY on rank 0 with the count in the error text - create the object again with more room).
 * world = 1 runs the same code on one GPU. */
#define B200CD_DIST_BLOB_BYTES 512
typedef struct b200cd_dist_stats {
    uint32_t rank, world;
    uint32_t local_triangles;  /* size of this rank's Morton range in the last step */
    uint32_t retries;          /* steps re-run after growing a buffer (cumulative) */
    uint64_t ghosts;           /* ghost records received from lower ranks */
    uint64_t candidates, local_pairs;
    uint64_t total_pairs;      /* rank 0: pairs in the gathered list */
    /* device times (CUDA events on the context's stream) of the last step's phases, barriers included */
    float ms_keys_hist, ms_plan, ms_exchange, ms_build, ms_ghost_send, ms_local_query, ms_ghost_query, ms_gather,
          ms_sort, ms_step;
} b200cd_dist_stats;
int b200cd_dist_create(b200cd_ctx* ctx, uint32_t rank, uint32_t world, uint32_t ntris_total, double slack,
                       uint64_t pair_capacity, b200cd_dist** out);
int b200cd_dist_export(b200cd_dist* dist, uint8_t* blob_out /* B200CD_DIST_BLOB_BYTES */);
int b200cd_dist_connect(b200cd_dist* dist, const uint8_t* blobs /* world x B200CD_DIST_BLOB_BYTES, rank order */);
int b200cd_dist_step(b200cd_dist* dist, const b200cd_mesh* mesh, const b200cd_params* params, const void** d_pairs_out,
                     uint64_t* count_out);
/* Pipelined frames: with on != 0 rank 0 sorts the gathered list on a side stream and goes straight on to the next frame
 * (the other ranks would otherwise wait for that sort at the next step's first barrier). The list the last
 * b200cd_dist_step returned is then valid after b200cd_dist_wait_sorted, which makes the context's stream wait for the
 * sort without blocking the host; the next b200cd_dist_step does it implicitly before it reuses the buffers. */
int b200cd_dist_set_async_sort(b200cd_dist* dist, int on);
int b200cd_dist_wait_sorted(b200cd_dist* dist);
/* a barrier across the ranks on the context's stream (flag barrier over peer memory) */
int b200cd_dist_barrier(b200cd_dist* dist);
int b200cd_dist_get_stats(b200cd_dist* dist, b200cd_dist_stats* out);
/* the rank's partial BVH (owned by dist): b200cd_bvh_validate / b200cd_get_stats work on it */
int b200cd_dist_bvh(b200cd_dist* dist, b200cd_bvh** out);
int b200cd_dist_destroy(b200cd_dist* dist);

/* Replicated mode ("each GPU ... receives the BVH, broadcast via NCCL over NVLink"): NCCL is loaded at run time
 * (dlopen of libnccl.so.2; B200CD_E_NODEVICE if absent). Rank 0 calls b200cd_nccl_unique_id and ships the 128 bytes
 * to the other ranks; every rank calls b200cd_dist_nccl_init; b200cd_dist_broadcast_bvh then sends the three device
 * blobs of a BVH built on `root` (traversal nodes, leaf records, sorted ids) to BVHs made with b200cd_bvh_alloc_like
 * on the other ranks, which query their shard with b200cd_self_collide_shard / _device. */
int b200cd_nccl_unique_id(uint8_t* id128);
int b200cd_dist_nccl_init(b200cd_dist* dist, const uint8_t* id128);
int b200cd_dist_broadcast_bvh(b200cd_dist* dist, b200cd_bvh* bvh, uint32_t root);

#ifdef __cplusplus
}
#endif
#endif /* B200CD_H */
