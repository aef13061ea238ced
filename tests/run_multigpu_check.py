"""torchrun --nproc-per-node N tests/run_multigpu_check.py  (needs N >= 2 B200s; not collected by pytest -
tests/test_gpu_dist.py covers the multi-rank step on the driver's single-GPU box with ranks sharing the GPU)

Checks on real GPUs, over NVLink, that every multi-GPU mode gives exactly the single-GPU pair list: the C++ step
(b200cd_dist_step), the replicated BVH (built everywhere, or built on rank 0 and broadcast with NCCL by the library),
the Python-orchestrated partitioned build (peer-memory and NCCL send/recv variants), and the sharded uploads."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    del os.environ["NCCL_DEBUG"]
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
mgpu = importlib.import_module("gpu-computing-course_b200.multigpu")
ctx = cd.Context(lr)
for name, (xyz, idx), box in (("soup4m", mg.soup(1 << 22, seed=5), ((0,0,0),(1,1,1))), ("sheets1024", mg.two_sheets(1024), ((0,0,0),(1,1,1))),
                              ("cloth", mg.cloth_fold(500, 500), None)):
    p = cd.make_params(*box) if box else cd.default_params()
    mesh = ctx.mesh_from_arrays(np.zeros_like(xyz), np.zeros_like(idx))
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    xyz = np.ascontiguousarray(xyz, np.float32); idx = np.ascontiguousarray(idx, np.uint32)
    mgpu.upload_mesh_sharded(ctx, mesh, xyz.ctypes.data, idx.ctypes.data)
    torch.cuda.synchronize()
    x2, i2 = mesh.download()
    assert np.array_equal(x2, xyz) and np.array_equal(i2, idx), "sharded upload mismatch"
    if rank == 0: print(name, "sharded upload OK", flush=True)
    # double-buffered frames over peer memory: every rank pushes its slice of each frame into all the others
    fr = [ctx.mesh_from_arrays(np.zeros_like(xyz), np.zeros_like(idx)) for _ in range(2)]
    pm = mgpu.PeerMeshFrames(cd, ctx, fr)
    if pm.ok:
        hx, hx_ptr = cd.pinned_array(xyz.shape, np.float32); hi, hi_ptr = cd.pinned_array(idx.shape, np.uint32)
        hx2, hx2_ptr = cd.pinned_array(xyz.shape, np.float32)
        hx[:] = xyz; hi[:] = idx; hx2[:] = xyz[::-1]
        pm.upload_async(0, hx_ptr, hi_ptr); pm.upload_async(1, hx2_ptr, hi_ptr)
        for k, want in ((0, xyz), (1, xyz[::-1])):
            x2, i2 = pm.wait(k).download()
            assert np.array_equal(x2, want) and np.array_equal(i2, idx), f"peer-memory frame {k} mismatch"
        pm.upload_async(0, hx2_ptr, hi_ptr)              # reuse of a frame object
        x2, _ = pm.wait(0).download()
        assert np.array_equal(x2, xyz[::-1])
        if rank == 0: print(name, "peer-memory frames OK", flush=True)
        torch.cuda.synchronize(); dist.barrier(); pm.close(); dist.barrier()
    elif rank == 0:
        print(name, "peer-memory frames SKIPPED:", pm.error, flush=True)
    for m in fr:
        m.destroy()
    bvh = ctx.bvh_build(mesh, p)
    single = ctx.self_collide(bvh, sorted=True)            # what ONE GPU gives (every rank computes it for itself)
    sh = mgpu.ShardedSelfCollision(cd, ctx)
    ref = sh.step(bvh, mesh, p)
    ref_np = mgpu.unpack_pairs(ref) if rank == 0 else None
    if rank == 0:
        assert np.array_equal(ref_np, single), "replicated + sharded query differs from the single-GPU list"
        print(name, "replicated (every rank builds) OK", len(single), flush=True)
    # replicated, BVH received: rank 0 builds, the library broadcasts nodes / leaf records / ids with NCCL
    bc = mgpu.BvhBroadcaster(cd, ctx, mesh.ntris)
    rbvh = bvh if rank == 0 else None
    for it in range(2):
        if rank == 0:
            ctx.bvh_rebuild(bvh, mesh, p)
        rbvh = bc.receive_or_send(rbvh, 0)
        got = sh.step(rbvh, mesh, p, rebuild=False)
        if rank == 0:
            assert np.array_equal(mgpu.unpack_pairs(got), single), "broadcast BVH: sharded query differs"
    if rank != 0:
        chk = rbvh.validate()                               # the received tree passes the reference's self-checks
        assert chk["null_parent_internal"] == 1 and chk["wrong_bound_count"] == 0 and chk["null_child"] == 0, chk
        rbvh.destroy()
    bc.close()
    if rank == 0: print(name, "replicated (BVH broadcast over NCCL by the library) OK", flush=True)
    # the multi-GPU step in C++ (b200cd_dist_step): one call per rank and frame
    ds = mgpu.DistSelfCollision(cd, ctx, mesh, p)
    for it in range(3):
        got = ds.step()
        torch.cuda.synchronize()
        if rank == 0:
            g = mgpu.unpack_pairs(got)
            ok = np.array_equal(g, single)
            st = ds.stats
            print(name, "b200cd_dist_step iter", it, "pairs", len(g), "OK" if ok else "MISMATCH",
                  {k: st[k] for k in ("local_triangles", "ghosts", "retries", "ms_step")}, flush=True)
            assert ok
    ds.close()
    for pm in (True, False):
        pr = mgpu.PartitionedSelfCollision(cd, ctx, mesh, p, peer_memory=pm)
        for it in range(3):
            got = pr.step()
            if rank == 0:
                g = mgpu.unpack_pairs(got)
                ok = np.array_equal(g, ref_np)
                print(name, "peer_memory" if pm else "nccl", "iter", it, "pairs", len(g), "ref", len(ref_np), "OK" if ok else "MISMATCH", pr.stats, flush=True)
                assert ok
        torch.cuda.synchronize(); dist.barrier()
        pr.close(); pr.part.bvh.destroy()
    bvh.destroy(); mesh.destroy()
dist.barrier()
dist.destroy_process_group()
