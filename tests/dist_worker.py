"""One rank of a b200cd_dist job, started by tests/test_gpu_dist.py (and usable by hand):

    python tests/dist_worker.py RANK WORLD RENDEZVOUS_DIR WORKLOAD [STEPS]

The ranks find each other through files in RENDEZVOUS_DIR (export blobs, then "done" markers): the library asks
nothing more of the host language than moving B200CD_DIST_BLOB_BYTES per rank. Device = RANK modulo the number of
visible GPUs, so WORLD processes can share ONE GPU (the driver's single-GPU test box): the kernels of the ranks are
then time-sliced, the peer-memory stores, remote atomics and flag barriers are the same code as over NVLink.
Rank 0 writes the sorted pair list of every step to RENDEZVOUS_DIR/pairs_<step>.npy and the per-rank statistics to
stats_<rank>.json."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def workload(mg, cd, name):
    """(xyz, idx, params) - small and medium meshes of the bench generators"""
    unit = ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))
    if name.endswith("/30"):  # the 30-bit key variant (morton.h:31-40): many duplicate keys, 14-bit histogram shift
        xyz, idx, params = workload(mg, cd, name[:-3])
        params.key_bits = 30
        return xyz, idx, params
    if name.startswith("soup"):
        xyz, idx = mg.soup(int(name[4:]), seed=5)
        return xyz, idx, cd.make_params(*unit)
    if name.startswith("sheets"):
        xyz, idx = mg.two_sheets(int(name[6:]), seed=7)
        return xyz, idx, cd.make_params(*unit)
    if name.startswith("cloth"):
        s = int(name[5:])
        xyz, idx = mg.cloth_fold(s, s)
        return xyz, idx, cd.default_params()
    if name.startswith("dup"):
        # edge case: clusters of IDENTICAL triangles (equal Morton keys, runs far longer than the sort's fix-up handles,
        # ties ordered by triangle id) - every rank must order them exactly like one GPU does
        k = int(name[3:])
        bx, bi = mg.soup(max(k // 64, 4), seed=11)
        xyz = np.tile(bx, (64, 1))
        idx = (np.tile(bi, (64, 1)) + (np.arange(64, dtype=np.uint32).repeat(len(bi)) * np.uint32(len(bx)))[:, None]).astype(np.uint32)
        return xyz, idx, cd.make_params(*unit)
    if name.startswith("tiny"):
        # edge case: fewer triangles than ranks x 2 (ranks whose Morton range is empty or a single leaf)
        xyz, idx = mg.soup(int(name[4:]), h=0.4, seed=3)
        return xyz, idx, cd.make_params(*unit)
    raise SystemExit(f"unknown workload {name}")


def exchange(rdv, tag, rank, world, payload, timeout=120.0):
    """all-gather of one bytes object per rank through files"""
    tmp = os.path.join(rdv, f".{tag}_{rank}.tmp")
    with open(tmp, "wb") as f:
        f.write(payload)
    os.replace(tmp, os.path.join(rdv, f"{tag}_{rank}.bin"))
    out, t0 = [], time.time()
    for r in range(world):
        p = os.path.join(rdv, f"{tag}_{r}.bin")
        while not os.path.exists(p):
            if os.path.exists(os.path.join(rdv, "FAILED")):
                raise SystemExit(f"rank {rank}: another rank failed")
            if time.time() - t0 > timeout:
                raise SystemExit(f"rank {rank}: rank {r} never published {tag}")
            time.sleep(0.005)
        out.append(open(p, "rb").read())
    return out


def main():
    rank, world, rdv, name = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    import torch
    cd = importlib.import_module("gpu-computing-course_b200.binding")
    mg = importlib.import_module("gpu-computing-course_b200.meshgen")
    mgpu = importlib.import_module("gpu-computing-course_b200.multigpu")
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = cd.Context(dev)
    xyz, idx, params = workload(mg, cd, name)
    mesh = ctx.mesh_from_arrays(xyz, idx)
    dist = ctx.dist_create(rank, world, mesh.ntris, slack=float(os.environ.get("B200CD_TEST_SLACK", "1.5")),
                           pair_capacity=int(os.environ.get("B200CD_TEST_PAIR_CAP", "0")))
    dist.connect(exchange(rdv, "blob", rank, world, dist.export()))
    exchange(rdv, "connected", rank, world, b"1")      # nobody steps before everybody has mapped everybody
    stats = []
    for k in range(steps):
        if k == 1:                                      # a second frame: same topology, moved vertices
            xyz = (xyz + np.float32(1e-3) * np.sin(37.0 * xyz[:, ::-1])).astype(np.float32)
            mesh.update(xyz=xyz)
        if k == 1:
            dist.set_async_sort(True)                   # second frame: rank 0's final sort on a side stream
        ptr, count = dist.step(mesh, params)
        dist.wait_sorted()                              # (no-op unless the sort went to the side stream)
        ctx.synchronize()                               # the list is valid in stream order on the CONTEXT's stream
        if rank == 0:
            pairs = mgpu.unpack_pairs(mgpu.device_pairs_as_tensor(ptr, count, torch.device("cuda", dev)))
            np.save(os.path.join(rdv, f"pairs_{k}.npy"), pairs)
        stats.append(dist.stats())
    chk = dist.bvh().validate(mesh)                     # the rank's own tree passes the reference's self-checks
    ctx.synchronize()
    with open(os.path.join(rdv, f"stats_{rank}.json"), "w") as f:
        json.dump({"steps": stats, "checks": chk}, f)
    exchange(rdv, "done", rank, world, b"1")            # peers may still be storing into my buffers before this
    dist.destroy()
    mesh.destroy()
    ctx.destroy()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        try:  # tell the other ranks at once (they would otherwise wait for their barrier / rendezvous timeouts)
            open(os.path.join(sys.argv[3], "FAILED"), "w").close()
        except Exception:
            pass
        raise
