"""ctypes binding of libb200cd.so (include/b200cd.h) — the call a Python user makes.

Mirrors the reference's entry-point sequence (reference CollisionDetection/main.cu:47-174):

    ctx  = Context(device)                      # device 0 / default stream in main.cu
    mesh = ctx.mesh_load_obj(path)              # loadObj, load_obj.h:24   (or mesh_from_arrays)
    bvh  = ctx.bvh_build(mesh, params)          # Morton + sort + fillLeafNodes + hierarchy + refit
    pairs = ctx.self_collide(bvh, sorted=True)  # findCollisions, collision.cuh:73 -> (count, 2) uint32

There is no CPU fallback: if the library or a B200 is missing this module raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200CD_LIB_PATH: a tuning build of the library (make ... EXTRA_NVFLAGS=-D...), for A/B runs of bench.py
LIB_PATH = os.environ.get("B200CD_LIB_PATH") or os.path.join(_HERE, "lib", "libb200cd.so")
_LIB = None

OK, E_INVALID, E_CUDA, E_NOMEM, E_IO, E_PARSE, E_CAPACITY, E_DEPTH, E_NODEVICE, E_TOOBIG, E_PEER = range(11)

# every symbol include/b200cd.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "b200cd_create", "b200cd_destroy", "b200cd_set_stream", "b200cd_synchronize", "b200cd_get_stats",
    "b200cd_strerror", "b200cd_last_error", "b200cd_abi_version", "b200cd_default_params", "b200cd_host_alloc",
    "b200cd_host_free", "b200cd_mesh_load_obj", "b200cd_mesh_from_arrays", "b200cd_mesh_from_device",
    "b200cd_mesh_update", "b200cd_mesh_info", "b200cd_mesh_download", "b200cd_mesh_destroy",
    "b200cd_bvh_build", "b200cd_bvh_rebuild", "b200cd_bvh_refit", "b200cd_bvh_download", "b200cd_bvh_validate",
    "b200cd_bvh_destroy", "b200cd_self_collide", "b200cd_self_collide_shard", "b200cd_self_collide_device",
    "b200cd_sort_pairs_device", "b200cd_bvh_alloc_like",
    "b200cd_morton_keys_device", "b200cd_key_histogram_device", "b200cd_bvh_alloc_partial", "b200cd_bvh_key_buffers",
    "b200cd_partition_keys_device", "b200cd_bvh_build_partial", "b200cd_bvh_chunk_boxes_device",
    "b200cd_select_ghosts_device", "b200cd_bvh_ghost_buffer", "b200cd_collide_ghosts_device",
    "b200cd_ipc_export", "b200cd_ipc_open", "b200cd_ipc_close", "b200cd_bvh_set_peers", "b200cd_partition_counts_device",
    "b200cd_partition_to_peers_device", "b200cd_send_ghosts_to_peers_device", "b200cd_ghost_counter_reset",
    "b200cd_ghost_counter_read", "b200cd_mesh_update_slice", "b200cd_mesh_device_buffers",
    "b200cd_obj_parse_host", "b200cd_host_array_free", "b200cd_mesh_update_async", "b200cd_mesh_wait",
    "b200cd_mesh_ipc_export", "b200cd_mesh_set_peers", "b200cd_mesh_update_slice_async", "b200cd_partition_plan_device",
    "b200cd_unique_triangles", "b200cd_unique_triangles_device",
    "b200cd_dist_create", "b200cd_dist_export", "b200cd_dist_connect", "b200cd_dist_step", "b200cd_dist_barrier",
    "b200cd_dist_get_stats", "b200cd_dist_bvh", "b200cd_dist_destroy", "b200cd_dist_set_async_sort", "b200cd_dist_wait_sorted", "b200cd_nccl_unique_id", "b200cd_dist_nccl_init",
    "b200cd_dist_broadcast_bvh", "b200cd_trace_dump", "b200cd_trace_enable", "b200cd_device_count", "b200cd_copy_to_host",
]

DIST_BLOB_BYTES = 512


class Params(C.Structure):
    _fields_ = [("morton_origin", C.c_double * 3), ("morton_extent", C.c_double * 3), ("key_bits", C.c_int32),
                ("auto_box", C.c_int32), ("pair_capacity_hint", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("ms_upload", "ms_morton", "ms_sort", "ms_hierarchy", "ms_refit", "ms_build",
                                         "ms_traverse", "ms_narrow", "ms_pair_sort", "ms_query", "ms_download")] + \
               [("ntris", C.c_uint32), ("nverts", C.c_uint32), ("candidates", C.c_uint64), ("pairs", C.c_uint64),
                ("sort_passes", C.c_uint32), ("query_retries", C.c_uint32), ("kernel_launches", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("warp_steps", C.c_uint64), ("start_entries", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Checks(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("null_parent_internal", "wrong_bound_count", "null_child",
                                          "uninit_box_internal", "null_parent_leaf", "bad_triangle",
                                          "uninit_box_leaf", "unsorted_keys", "box_not_enclosing")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DistStats(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32), ("local_triangles", C.c_uint32), ("retries", C.c_uint32),
                ("ghosts", C.c_uint64), ("candidates", C.c_uint64), ("local_pairs", C.c_uint64), ("total_pairs", C.c_uint64)] + \
               [(k, C.c_float) for k in ("ms_keys_hist", "ms_plan", "ms_exchange", "ms_build", "ms_ghost_send",
                                         "ms_local_query", "ms_ghost_query", "ms_gather", "ms_sort", "ms_step")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


NODE32 = np.dtype([("lo", np.float32, 3), ("hi", np.float32, 3), ("left", np.int32), ("right", np.int32)])


class B200cdError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        msg = lib().b200cd_strerror(status).decode()
        super().__init__(f"{where}: {msg}" + (f" ({detail})" if detail else ""))


def lib():
    """Load libb200cd.so; fail loudly if it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                               f"(make -C gpu-computing-course_b200/csrc). There is no CPU fallback.")
        _LIB = C.CDLL(LIB_PATH)
        _LIB.b200cd_strerror.restype = C.c_char_p
        _LIB.b200cd_last_error.restype = C.c_char_p
        _LIB.b200cd_last_error.argtypes = [C.c_void_p]
    return _LIB


def default_params(key_bits=63):
    p = Params()
    lib().b200cd_default_params(C.byref(p))
    p.key_bits = key_bits
    return p


def make_params(origin=None, extent=None, key_bits=63, auto_box=False, pair_capacity_hint=0):
    p = default_params(key_bits)
    if origin is not None:
        p.morton_origin[:] = list(origin)
    if extent is not None:
        p.morton_extent[:] = list(extent)
    p.auto_box = 1 if auto_box else 0
    p.pair_capacity_hint = pair_capacity_hint
    return p


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def parse_obj(path):
    """The library's OBJ parser on its own (host only; reference dialect, load_obj.h:24-103).
    -> (xyz float32 [V,3], idx uint32 [N,3]); raises B200cdError(E_IO / E_PARSE) with the line number."""
    xyz_p, idx_p = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)()
    nv, nt = C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(256)
    rc = lib().b200cd_obj_parse_host(os.fsencode(path), C.byref(xyz_p), C.byref(nv), C.byref(idx_p), C.byref(nt), err,
                                     C.c_uint64(256))
    if rc != OK:
        raise B200cdError(rc, "obj_parse", err.value.decode(errors="replace"))
    xyz = np.ctypeslib.as_array(xyz_p, shape=(max(nv.value, 1), 3))[:nv.value].copy()
    idx = np.ctypeslib.as_array(idx_p, shape=(max(nt.value, 1), 3))[:nt.value].copy()
    lib().b200cd_host_array_free(xyz_p)
    lib().b200cd_host_array_free(idx_p)
    return xyz, idx


class Mesh:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        nv, nt = C.c_uint32(), C.c_uint32()
        lib().b200cd_mesh_info(self.h, C.byref(nv), C.byref(nt))
        self.nverts, self.ntris = nv.value, nt.value

    def update(self, xyz=None, idx=None):
        """overwrite vertices and/or indices in place from host arrays (same sizes)"""
        if xyz is not None:
            xyz = np.ascontiguousarray(xyz, np.float32)
            assert xyz.shape == (self.nverts, 3)
        if idx is not None:
            idx = np.ascontiguousarray(idx, np.uint32)
            assert idx.shape == (self.ntris, 3)
        self.ctx._chk(lib().b200cd_mesh_update(self.ctx.h, self.h, _ptr(xyz, C.c_float) if xyz is not None else None,
                                               _ptr(idx, C.c_uint32) if idx is not None else None, C.c_int(0)),
                      "mesh_update")

    def update_from_ptr(self, xyz_ptr, idx_ptr, on_device=False):
        """raw pointers (pinned host buffers or device memory); either may be 0/None"""
        self.ctx._chk(lib().b200cd_mesh_update(self.ctx.h, self.h, C.c_void_p(xyz_ptr or None),
                                               C.c_void_p(idx_ptr or None), C.c_int(1 if on_device else 0)),
                      "mesh_update")

    def update_async_from_ptr(self, xyz_ptr, idx_ptr):
        """enqueue the upload on the context's copy stream and return; the (pinned) host buffers must stay valid
        until wait(). Double-buffered frames: upload mesh B while mesh A is being built / queried."""
        self.ctx._chk(lib().b200cd_mesh_update_async(self.ctx.h, self.h, C.c_void_p(xyz_ptr or None),
                                                     C.c_void_p(idx_ptr or None)), "mesh_update_async")

    def update_slice_async_from_ptr(self, xyz_ptr, first_vert, nverts, idx_ptr, first_tri, ntris):
        """multi-GPU twin of update_async_from_ptr: my slice goes host -> my mesh -> every peer's mesh (set_peers),
        all on the copy stream; wait() returns when it has landed everywhere"""
        self.ctx._chk(lib().b200cd_mesh_update_slice_async(self.ctx.h, self.h, C.c_void_p(xyz_ptr or None), C.c_uint32(first_vert),
                                                           C.c_uint32(nverts), C.c_void_p(idx_ptr or None), C.c_uint32(first_tri),
                                                           C.c_uint32(ntris)), "mesh_update_slice_async")

    def ipc_export(self):
        """(128 handle bytes, [2 offsets]) of the vertex and index buffers, for the other ranks' ipc_open"""
        handles = (C.c_uint8 * 128)()
        offsets = (C.c_uint64 * 2)()
        self.ctx._chk(lib().b200cd_mesh_ipc_export(self.ctx.h, self.h, handles, offsets), "mesh_ipc_export")
        return bytes(handles), [int(o) for o in offsets]

    def set_peers(self, nranks, my_rank, peers):
        """peers[2 * r + {0, 1}] = rank r's vertex / index buffer as mapped on this GPU (0 for my own rank)"""
        arr = (C.c_void_p * (2 * nranks))(*[C.c_void_p(p or None) for p in peers])
        self.ctx._chk(lib().b200cd_mesh_set_peers(self.ctx.h, self.h, C.c_uint32(nranks), C.c_uint32(my_rank), arr), "mesh_set_peers")

    def wait(self):
        """block until a pending asynchronous upload has landed (and raise if its index check failed)"""
        self.ctx._chk(lib().b200cd_mesh_wait(self.ctx.h, self.h), "mesh_wait")

    def update_slice_from_ptr(self, xyz_ptr, first_vert, nverts, idx_ptr, first_tri, ntris):
        """raw host pointers to the slice's first vertex / triangle"""
        self.ctx._chk(lib().b200cd_mesh_update_slice(self.ctx.h, self.h, C.c_void_p(xyz_ptr or None), C.c_uint32(first_vert),
                                                     C.c_uint32(nverts), C.c_void_p(idx_ptr or None), C.c_uint32(first_tri),
                                                     C.c_uint32(ntris)), "mesh_update_slice")

    def device_buffers(self):
        v, i = C.c_void_p(), C.c_void_p()
        self.ctx._chk(lib().b200cd_mesh_device_buffers(self.ctx.h, self.h, C.byref(v), C.byref(i)), "mesh_device_buffers")
        return v.value, i.value

    def download(self):
        xyz = np.empty((self.nverts, 3), np.float32)
        idx = np.empty((self.ntris, 3), np.uint32)
        self.ctx._chk(lib().b200cd_mesh_download(self.ctx.h, self.h, _ptr(xyz, C.c_float), _ptr(idx, C.c_uint32)), "mesh_download")
        return xyz, idx

    def destroy(self):
        if self.h:
            lib().b200cd_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Bvh:
    def __init__(self, ctx, handle, ntris):
        self.ctx, self.h, self.ntris = ctx, handle, ntris

    def download(self, nodes=True, keys=True, ids=True):
        """parity hooks: (nodes[2n-1] NODE32, sorted_keys[n] u64, sorted_ids[n] u32)"""
        n = self.ntris
        a_nodes = np.empty(max(2 * n - 1, 0), NODE32) if nodes else None
        a_keys = np.empty(n, np.uint64) if keys else None
        a_ids = np.empty(n, np.uint32) if ids else None
        self.ctx._chk(lib().b200cd_bvh_download(
            self.ctx.h, self.h,
            a_nodes.ctypes.data_as(C.c_void_p) if nodes else None,
            _ptr(a_keys, C.c_uint64) if keys else None,
            _ptr(a_ids, C.c_uint32) if ids else None), "bvh_download")
        return a_nodes, a_keys, a_ids

    def validate(self, mesh=None):
        out = Checks()
        self.ctx._chk(lib().b200cd_bvh_validate(self.ctx.h, self.h, mesh.h if mesh else None, C.byref(out)), "bvh_validate")
        return out.as_dict()

    def destroy(self):
        if self.h:
            lib().b200cd_bvh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Context:
    """One GPU. Not thread-safe (same contract as the C ABI)."""

    def __init__(self, device=0, stream=None):
        self.h = C.c_void_p()
        rc = lib().b200cd_create(C.c_int(device), C.byref(self.h))
        if rc != OK:
            self.h = None
            raise B200cdError(rc, "b200cd_create")
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def _chk(self, rc, where):
        if rc != OK:
            raise B200cdError(rc, where, lib().b200cd_last_error(self.h).decode())

    def set_stream(self, cuda_stream_handle):
        """cudaStream_t handle as an int (0 = the legacy default stream); None = the context's private stream"""
        h = C.c_void_p(-1) if cuda_stream_handle is None else C.c_void_p(cuda_stream_handle)
        self._chk(lib().b200cd_set_stream(self.h, h), "set_stream")

    def synchronize(self):
        self._chk(lib().b200cd_synchronize(self.h), "synchronize")

    def trace_enable(self, on=True):
        lib().b200cd_trace_enable(C.c_int(1 if on else 0))

    def trace_dump(self, path):
        """append the kernel timeline collected since the last dump (needs B200CD_TRACE in the environment)"""
        self._chk(lib().b200cd_trace_dump(self.h, os.fsencode(path)), "trace_dump")

    def stats(self):
        s = Stats()
        self._chk(lib().b200cd_get_stats(self.h, C.byref(s)), "get_stats")
        return s.as_dict()

    # ---- mesh in
    def mesh_load_obj(self, path):
        m = C.c_void_p()
        self._chk(lib().b200cd_mesh_load_obj(self.h, os.fsencode(path), C.byref(m)), "mesh_load_obj")
        return Mesh(self, m)

    def mesh_from_arrays(self, xyz, idx):
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(idx, np.uint32).reshape(-1, 3)
        m = C.c_void_p()
        self._chk(lib().b200cd_mesh_from_arrays(self.h, _ptr(xyz, C.c_float), C.c_uint32(xyz.shape[0]),
                                                _ptr(idx, C.c_uint32), C.c_uint32(idx.shape[0]), C.byref(m)),
                  "mesh_from_arrays")
        return Mesh(self, m)

    def mesh_from_host_ptr(self, xyz_ptr, nverts, idx_ptr, ntris):
        """raw host pointers (e.g. pinned buffers from host_alloc)"""
        m = C.c_void_p()
        self._chk(lib().b200cd_mesh_from_arrays(self.h, C.c_void_p(xyz_ptr), C.c_uint32(nverts), C.c_void_p(idx_ptr),
                                                C.c_uint32(ntris), C.byref(m)), "mesh_from_arrays")
        return Mesh(self, m)

    def mesh_from_device(self, d_xyz_ptr, nverts, d_idx_ptr, ntris):
        m = C.c_void_p()
        self._chk(lib().b200cd_mesh_from_device(self.h, C.c_void_p(d_xyz_ptr), C.c_uint32(nverts),
                                                C.c_void_p(d_idx_ptr), C.c_uint32(ntris), C.byref(m)),
                  "mesh_from_device")
        return Mesh(self, m)

    # ---- build
    def bvh_build(self, mesh, params=None):
        params = params or default_params()
        b = C.c_void_p()
        self._chk(lib().b200cd_bvh_build(self.h, mesh.h, C.byref(params), C.byref(b)), "bvh_build")
        return Bvh(self, b, mesh.ntris)

    def bvh_rebuild(self, bvh, mesh, params=None):
        params = params or default_params()
        self._chk(lib().b200cd_bvh_rebuild(self.h, bvh.h, mesh.h, C.byref(params)), "bvh_rebuild")

    def bvh_refit(self, bvh, mesh):
        self._chk(lib().b200cd_bvh_refit(self.h, bvh.h, mesh.h), "bvh_refit")

    def bvh_alloc_like(self, ntris):
        b = C.c_void_p()
        self._chk(lib().b200cd_bvh_alloc_like(self.h, C.c_uint32(ntris), C.byref(b)), "bvh_alloc_like")
        return Bvh(self, b, ntris)

    # ---- query
    def self_collide(self, bvh, sorted=True, shard=0, nshards=1, chunk=0, cap=None):
        """-> (count, 2) uint32 array, lower ID first. Grows the host buffer on E_CAPACITY."""
        cnt = C.c_uint64()
        if cap is None:
            cap = 1 << 16
        while True:
            out = np.empty((cap, 2), np.uint32)
            rc = lib().b200cd_self_collide_shard(self.h, bvh.h, C.c_uint32(shard), C.c_uint32(nshards),
                                                 C.c_uint32(chunk), _ptr(out, C.c_uint32), C.c_uint64(cap),
                                                 C.byref(cnt), C.c_int(1 if sorted else 0))
            if rc == E_CAPACITY:
                cap = int(cnt.value)
                continue
            self._chk(rc, "self_collide")
            return out[:cnt.value]

    def self_collide_into(self, bvh, out_ptr, cap, sorted=True, shard=0, nshards=1, chunk=0):
        """result straight into a caller-owned host buffer (raw pointer); returns the pair count"""
        cnt = C.c_uint64()
        rc = lib().b200cd_self_collide_shard(self.h, bvh.h, C.c_uint32(shard), C.c_uint32(nshards), C.c_uint32(chunk),
                                             C.c_void_p(out_ptr), C.c_uint64(cap), C.byref(cnt),
                                             C.c_int(1 if sorted else 0))
        self._chk(rc, "self_collide")
        return int(cnt.value)

    def self_collide_device(self, bvh, sorted=True, shard=0, nshards=1, chunk=0):
        """-> (device pointer, count); the buffer is owned by the BVH until its next query"""
        cnt = C.c_uint64()
        ptr = C.c_void_p()
        self._chk(lib().b200cd_self_collide_device(self.h, bvh.h, C.c_uint32(shard), C.c_uint32(nshards),
                                                   C.c_uint32(chunk), C.c_int(1 if sorted else 0), C.byref(ptr),
                                                   C.byref(cnt)), "self_collide_device")
        return ptr.value, int(cnt.value)

    # ---- partitioned multi-GPU build (include/b200cd.h, "partitioned multi-GPU build"); raw device pointers
    def morton_keys_device(self, mesh, params, first, count, d_keys_out):
        self._chk(lib().b200cd_morton_keys_device(self.h, mesh.h, C.byref(params), C.c_uint32(first), C.c_uint32(count),
                                                  C.c_void_p(d_keys_out)), "morton_keys_device")

    def key_histogram_device(self, d_keys, count, shift, d_hist):
        self._chk(lib().b200cd_key_histogram_device(self.h, C.c_void_p(d_keys), C.c_uint32(count), C.c_int32(shift),
                                                    C.c_void_p(d_hist)), "key_histogram_device")

    def partition_plan_device(self, d_global_hist, d_local_hist, shift, world, d_splitters, d_counts):
        self._chk(lib().b200cd_partition_plan_device(self.h, C.c_void_p(d_global_hist), C.c_void_p(d_local_hist), C.c_int32(shift),
                                                     C.c_uint32(world), C.c_void_p(d_splitters or None), C.c_void_p(d_counts)),
                  "partition_plan_device")

    def bvh_alloc_partial(self, capacity, ghost_capacity, max_peers):
        b = C.c_void_p()
        self._chk(lib().b200cd_bvh_alloc_partial(self.h, C.c_uint32(capacity), C.c_uint64(ghost_capacity),
                                                 C.c_uint32(max_peers), C.byref(b)), "bvh_alloc_partial")
        return Bvh(self, b, 0)

    def bvh_key_buffers(self, bvh):
        k, i, cap = C.c_void_p(), C.c_void_p(), C.c_uint32()
        self._chk(lib().b200cd_bvh_key_buffers(self.h, bvh.h, C.byref(k), C.byref(i), C.byref(cap)), "bvh_key_buffers")
        return k.value, i.value, cap.value

    def partition_keys_device(self, bvh, d_keys, first_id, count, d_splitters, nsplit, d_keys_out, d_ids_out):
        counts = (C.c_uint64 * (nsplit + 1))()
        self._chk(lib().b200cd_partition_keys_device(self.h, bvh.h, C.c_void_p(d_keys), C.c_uint32(first_id), C.c_uint32(count),
                                                     C.c_void_p(d_splitters), C.c_uint32(nsplit), C.c_void_p(d_keys_out),
                                                     C.c_void_p(d_ids_out), counts), "partition_keys_device")
        return [int(c) for c in counts]

    def bvh_build_partial(self, bvh, mesh, params, count):
        self._chk(lib().b200cd_bvh_build_partial(self.h, bvh.h, mesh.h, C.byref(params), C.c_uint32(count)), "bvh_build_partial")
        bvh.ntris = count

    def bvh_chunk_boxes_device(self, bvh, K, d_boxes_out):
        self._chk(lib().b200cd_bvh_chunk_boxes_device(self.h, bvh.h, C.c_uint32(K), C.c_void_p(d_boxes_out)), "bvh_chunk_boxes_device")

    def select_ghosts_device(self, bvh, d_peer_boxes, npeers, K, peer_mask):
        """-> (device pointer of the per-peer lists, stride in records, [count per peer])"""
        ptr, stride = C.c_void_p(), C.c_uint64()
        counts = (C.c_uint64 * npeers)()
        self._chk(lib().b200cd_select_ghosts_device(self.h, bvh.h, C.c_void_p(d_peer_boxes), C.c_uint32(npeers), C.c_uint32(K),
                                                    C.c_uint32(peer_mask), C.byref(ptr), C.byref(stride), counts),
                  "select_ghosts_device")
        return ptr.value, int(stride.value), [int(c) for c in counts]

    def bvh_ghost_buffer(self, bvh):
        ptr, cap = C.c_void_p(), C.c_uint64()
        self._chk(lib().b200cd_bvh_ghost_buffer(self.h, bvh.h, C.byref(ptr), C.byref(cap)), "bvh_ghost_buffer")
        return ptr.value, int(cap.value)

    def collide_ghosts_device(self, bvh, nghost, keep_pairs=True):
        cnt, ptr = C.c_uint64(), C.c_void_p()
        self._chk(lib().b200cd_collide_ghosts_device(self.h, bvh.h, C.c_uint64(nghost), C.c_int(1 if keep_pairs else 0),
                                                     C.byref(ptr), C.byref(cnt)), "collide_ghosts_device")
        return ptr.value, int(cnt.value)

    # ---- peer-memory variant (CUDA IPC between the ranks' processes)
    def ipc_export(self, bvh):
        handles = (C.c_uint8 * 256)()
        offsets = (C.c_uint64 * 4)()
        self._chk(lib().b200cd_ipc_export(self.h, bvh.h, handles, offsets), "ipc_export")
        return bytes(handles), [int(o) for o in offsets]

    def ipc_open(self, handle64):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle64)
        ptr = C.c_void_p()
        self._chk(lib().b200cd_ipc_open(self.h, buf, C.byref(ptr)), "ipc_open")
        return ptr.value

    def ipc_close(self, ptr):
        lib().b200cd_ipc_close(self.h, C.c_void_p(ptr))

    def bvh_set_peers(self, bvh, nranks, my_rank, peers):
        arr = (C.c_void_p * (4 * nranks))(*[C.c_void_p(p) for p in peers])
        self._chk(lib().b200cd_bvh_set_peers(self.h, bvh.h, C.c_uint32(nranks), C.c_uint32(my_rank), arr), "bvh_set_peers")

    def partition_counts_device(self, d_keys, count, d_splitters, nsplit, d_counts_out):
        self._chk(lib().b200cd_partition_counts_device(self.h, C.c_void_p(d_keys), C.c_uint32(count), C.c_void_p(d_splitters),
                                                       C.c_uint32(nsplit), C.c_void_p(d_counts_out)), "partition_counts_device")

    def partition_to_peers_device(self, bvh, d_keys, first_id, count, d_splitters, nsplit, d_recv_offsets):
        self._chk(lib().b200cd_partition_to_peers_device(self.h, bvh.h, C.c_void_p(d_keys), C.c_uint32(first_id), C.c_uint32(count),
                                                         C.c_void_p(d_splitters), C.c_uint32(nsplit), C.c_void_p(d_recv_offsets)),
                  "partition_to_peers_device")

    def send_ghosts_to_peers_device(self, bvh, d_peer_boxes, npeers, K, peer_mask):
        self._chk(lib().b200cd_send_ghosts_to_peers_device(self.h, bvh.h, C.c_void_p(d_peer_boxes), C.c_uint32(npeers),
                                                           C.c_uint32(K), C.c_uint32(peer_mask)), "send_ghosts_to_peers_device")

    def ghost_counter_reset(self, bvh):
        self._chk(lib().b200cd_ghost_counter_reset(self.h, bvh.h), "ghost_counter_reset")

    def ghost_counter_read(self, bvh):
        cnt = C.c_uint64()
        self._chk(lib().b200cd_ghost_counter_read(self.h, bvh.h, C.byref(cnt)), "ghost_counter_read")
        return int(cnt.value)

    # ---- the sorted set of triangle IDs in a pair list (reference makeAndPrintSet, main.cu:33-45), on the device
    def unique_triangles(self, bvh):
        """IDs of every triangle in at least one colliding pair of the last query on `bvh`, ascending (uint32 array)"""
        cnt = C.c_uint64()
        rc = lib().b200cd_unique_triangles(self.h, bvh.h, None, C.c_uint64(0), C.byref(cnt))
        if rc not in (OK, E_CAPACITY):
            self._chk(rc, "unique_triangles")
        out = np.empty(int(cnt.value), np.uint32)
        if cnt.value:
            self._chk(lib().b200cd_unique_triangles(self.h, bvh.h, _ptr(out, C.c_uint32), C.c_uint64(out.size), C.byref(cnt)),
                      "unique_triangles")
        return out[:cnt.value]

    def unique_triangles_device(self, d_pairs, count, id_space):
        """-> (device pointer to ascending uint32 IDs, count); library-owned until the next call on this context"""
        cnt, ptr = C.c_uint64(), C.c_void_p()
        self._chk(lib().b200cd_unique_triangles_device(self.h, C.c_void_p(d_pairs or None), C.c_uint64(count), C.c_uint32(id_space),
                                                       C.byref(ptr), C.byref(cnt)), "unique_triangles_device")
        return ptr.value, int(cnt.value)

    # ---- the multi-GPU step in C++ (include/b200cd.h, b200cd_dist_*)
    def dist_create(self, rank, world, ntris_total, slack=1.5, pair_capacity=0):
        return Dist(self, rank, world, ntris_total, slack, pair_capacity)

    def sort_pairs_device(self, d_ptr, count, id_bits=0):
        self._chk(lib().b200cd_sort_pairs_device(self.h, C.c_void_p(d_ptr), C.c_uint64(count), C.c_uint32(id_bits)),
                  "sort_pairs_device")

    def destroy(self):
        if self.h:
            lib().b200cd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Dist:
    """One rank of the partitioned multi-GPU self-collision (b200cd_dist_*). The caller moves the export blobs
    between the processes (exchange = a callable: my blob -> list of every rank's blob in rank order), e.g.
    torch.distributed.all_gather_object; nothing else of a step goes through the host language."""

    def __init__(self, ctx, rank, world, ntris_total, slack=1.5, pair_capacity=0):
        self.ctx, self.rank, self.world, self.ntris = ctx, rank, world, ntris_total
        self.h = C.c_void_p()
        ctx._chk(lib().b200cd_dist_create(ctx.h, C.c_uint32(rank), C.c_uint32(world), C.c_uint32(ntris_total), C.c_double(slack),
                                          C.c_uint64(pair_capacity), C.byref(self.h)), "dist_create")

    def export(self):
        blob = (C.c_uint8 * DIST_BLOB_BYTES)()
        self.ctx._chk(lib().b200cd_dist_export(self.h, blob), "dist_export")
        return bytes(blob)

    def connect(self, blobs):
        assert len(blobs) == self.world and all(len(b) == DIST_BLOB_BYTES for b in blobs)
        buf = (C.c_uint8 * (DIST_BLOB_BYTES * self.world)).from_buffer_copy(b"".join(blobs))
        self.ctx._chk(lib().b200cd_dist_connect(self.h, buf), "dist_connect")

    def step(self, mesh, params):
        """-> (device pointer of the sorted pair list, count) on rank 0, (None, 0) elsewhere"""
        ptr, cnt = C.c_void_p(), C.c_uint64()
        self.ctx._chk(lib().b200cd_dist_step(self.h, mesh.h, C.byref(params), C.byref(ptr), C.byref(cnt)), "dist_step")
        return ptr.value, int(cnt.value)

    def set_async_sort(self, on=True):
        """rank 0's final sort on a side stream (pipelined frames); call wait_sorted() before reading a step's list"""
        self.ctx._chk(lib().b200cd_dist_set_async_sort(self.h, C.c_int(1 if on else 0)), "dist_set_async_sort")

    def wait_sorted(self):
        self.ctx._chk(lib().b200cd_dist_wait_sorted(self.h), "dist_wait_sorted")

    def barrier(self):
        self.ctx._chk(lib().b200cd_dist_barrier(self.h), "dist_barrier")

    def stats(self):
        s = DistStats()
        self.ctx._chk(lib().b200cd_dist_get_stats(self.h, C.byref(s)), "dist_get_stats")
        return s.as_dict()

    def bvh(self):
        """the rank's partial BVH (owned by this object: do not destroy)"""
        b = C.c_void_p()
        self.ctx._chk(lib().b200cd_dist_bvh(self.h, C.byref(b)), "dist_bvh")
        out = Bvh(self.ctx, b, 0)
        out.destroy = lambda: None
        return out

    def nccl_init(self, id128):
        buf = (C.c_uint8 * 128).from_buffer_copy(id128)
        self.ctx._chk(lib().b200cd_dist_nccl_init(self.h, buf), "dist_nccl_init")

    def broadcast_bvh(self, bvh, root=0):
        self.ctx._chk(lib().b200cd_dist_broadcast_bvh(self.h, bvh.h, C.c_uint32(root)), "dist_broadcast_bvh")

    def destroy(self):
        if self.h:
            lib().b200cd_dist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    rc = lib().b200cd_nccl_unique_id(buf)
    if rc != OK:
        raise B200cdError(rc, "nccl_unique_id")
    return bytes(buf)


def host_alloc(nbytes):
    p = C.c_void_p()
    rc = lib().b200cd_host_alloc(C.byref(p), C.c_uint64(nbytes))
    if rc != OK:
        raise B200cdError(rc, "host_alloc")
    return p.value


def host_free(ptr):
    lib().b200cd_host_free(C.c_void_p(ptr))


def pinned_array(shape, dtype):
    """numpy array over page-locked memory (kept alive by the returned object)"""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    ptr = host_alloc(max(n * dtype.itemsize, 1))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    return arr, ptr
