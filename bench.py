#!/usr/bin/env python
"""bench.py — Mtri/s end-to-end self-collision (BVH build + query) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one pass of the hot path over one synthetic mesh: BVH build (Morton keys, radix
sort, Karras hierarchy, refit) + self-collision query (traversal, fp64 narrow phase, sorted pair
list). Workloads (BASELINE.json `configs`):
    soup16m    C4  random triangle soup, 2^24 triangles in the unit cube   (default at N = 1)
    sheets64m  C5  two 4096x4096-quad sheets, 2^26 triangles               (default at N > 1)
    cloth1m    C3  accordion-folded sheet, 1 002 528 triangles, dense contacts
    flag1m     C1/C2 stand-in for the missing flag mesh, 1 262 460 triangles
N > 1: one rank per GPU (torchrun). Default: the build and the query are PARTITIONED - each rank owns
one Morton range of the triangles (distributed sort over NCCL, local tree + query, ghost exchange);
--mode replicated: every rank builds the whole BVH and the queries are sharded. Either way the
per-rank pair lists are gathered over NCCL and sorted on rank 0; strong scaling (the mesh is fixed).

Prints ONE JSON line on rank 0. `value` = triangles / device time with the mesh resident in HBM;
`e2e` = the same through the host-buffer C-ABI calls (H2D of the mesh and D2H of the pair list
inside the timed region). `--impl reference` times the reference's own CPU functions
(oracle/_ref, else the oracle port) on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

# rank 0 must print exactly ONE line on stdout; NCCL_DEBUG=VERSION (set in this image) makes NCCL
# printf its banner there, so drop that level (warnings and above are kept if asked for)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    del os.environ["NCCL_DEBUG"]

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "gpu-computing-course_b200"

METRIC = "Mtri/s end-to-end self-collision (build+query)"
UNIT = "Mtri/s"
UNIT_CUBE = ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))
SCALING_WORKLOAD = "sheets64m"  # default workload at N > 1

WORKLOADS = {
    # name: (kind, args, morton box or None = reference constants)
    "soup16m": ("soup", dict(n=1 << 24, seed=1234), UNIT_CUBE),
    "soup1m": ("soup", dict(n=1 << 20, seed=1234), UNIT_CUBE),
    "sheets64m": ("two_sheets", dict(nq=4096, seed=7), UNIT_CUBE),
    "sheets16m": ("two_sheets", dict(nq=2048, seed=7), UNIT_CUBE),
    "cloth1m": ("cloth_fold", dict(nx=708, ny=708), None),
    "flag1m": ("flag", dict(nx=795, nz=794), None),
}


def workload_sizes(mg, name):
    kind, kw, _ = WORKLOADS[name]
    if kind == "soup":
        return 3 * kw["n"], kw["n"]
    if kind == "two_sheets":
        return mg.grid_sizes(kw["nq"], kw["nq"], 2)
    if kind == "cloth_fold":
        return mg.grid_sizes(kw["nx"], kw["ny"])
    return mg.grid_sizes(kw["nx"], kw["nz"])


def generate(mg, name, out=None):
    kind, kw, _ = WORKLOADS[name]
    return getattr(mg, kind)(**kw, out=out)


def algorithmic_bytes(n, nverts, ncand, npairs, sort_passes, recs=True, quant=False):
    """Compulsory HBM bytes per stage for OUR data layout, each array read or
    written once per pass. n triangles, nverts vertices. recs: K1 also writes the face-ordered 64 B leaf
    records and the tree build moves those (single-GPU full builds) instead of gathering indices + vertices.
    quant: the tree build also writes the 32 B quantised node pairs and the traversal reads those instead of the
    64 B exact ones (triangle soups, csrc/collide.cu broad_uses_quantised_nodes)."""
    return {
        # K1: 12 B indices + one 16 B float4 per vertex (each vertex is read at least once) + 8 B key out (+ 64 B record)
        "morton": 12 * n + 16 * nverts + 8 * n + (64 * n if recs else 0),
        # K2: histogram reads the keys once; every pass reads and writes (8 B key + 4 B id)
        #     (first pass generates the ids: no 4 B read)
        #     hybrid sort (< 8 passes): the fix-up reads the keys once more
        "sort": 8 * n + sort_passes * 24 * n - 4 * n + (8 * n if sort_passes < 8 else 0),
        # K3+K4 (fused): sorted ids + keys + indices + vertices in, 64 B leaf record + 64 B node pair out
        "tree": 4 * n + 8 * n + (64 * n if recs else 12 * n + 16 * nverts) + 64 * n + 64 * n + (32 * n if quant else 0),
        # K5: every node pair (64 B, or 32 B quantised) and every query record (64 B) once, 8 B per candidate out
        "traverse": (32 if quant else 64) * n + 64 * n + 8 * ncand,
        # K6: candidate list in, every leaf record it names at most once from HBM (compulsory traffic; repeats are
        #     cache hits), 8 B per pair out
        "narrow": 8 * ncand + 64 * min(2 * ncand, n) + 8 * npairs,
    }


def uses_quantised_nodes(ntris, nverts):
    """mirror of csrc/collide.cu broad_uses_quantised_nodes: soups (V >= 1.5 N) unless B200CD_BROAD_QUANT says otherwise"""
    e = os.environ.get("B200CD_BROAD_QUANT")
    if os.environ.get("B200CD_TRAVERSAL", "2")[:1] != "2":
        return False
    if e is not None:
        return e[:1] != "0"
    return 2 * nverts >= 3 * ntris


def contract_bytes(n, nverts, npairs):
    """SURVEY.md section 8(d)'s compulsory bytes per stage for the layout the survey planned (8-pass sort of 12-byte items,
    32-byte nodes, separate AABB / parent arrays) - the contract the judge divides by, independent of OUR layout:
    K1 = 48N + 16V, K2 = 200N, K3 = 24N, K4 = 108N, K5 (traversal + narrow phase) = 100N + 16V + 8 pairs;
    build = 380N + 16V, end to end = 480N + 32V + 8 pairs."""
    return {"morton": 48 * n + 16 * nverts, "sort": 200 * n, "tree": 24 * n + 108 * n,
            "query": 100 * n + 16 * nverts + 8 * npairs,
            "e2e": 480 * n + 32 * nverts + 8 * npairs}


def golden_checksums():
    path = os.path.join(ROOT, "tests", "golden", "checksums.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def matches_oracle(workload, npairs, checksum):
    """the sorted pair list equals the one the reference's host functions and the CPU oracle produce for this workload
    (count + two wrap-around sums, tests/golden/checksums.json made by tests/golden/make_checksums.py); None = no golden"""
    g = golden_checksums().get(workload)
    if not g or checksum is None:
        return None
    return bool(int(npairs) == int(g["pairs"]) and [int(c) for c in checksum] == [int(c) for c in g["checksum"]])


STAGE_KERNEL = {"morton": "morton_kernel", "sort": "rs_pass (x passes) + rs_histogram + rs_fixup", "tree": "build_kernel (+ upper_kernel)",
                "traverse": "broad_kernel (+ entry_kernel)", "narrow": "narrow_kernel"}


class ClockSampler(threading.Thread):
    """nvidia-smi's clocks / throttle reasons, sampled through NVML while the timed region runs"""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.sm_max = index, period, [], set(), None
        self._stop_evt = threading.Event()
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------ CPU arms

def cpu_reference_run(name, sample_tris, threads, repeats=1):
    """The reference's own host functions (oracle/_ref) — or the oracle port — on a bounded
    sample of workload `name`: same generator and contact density, `sample_tris` triangles placed
    inside the reference's hard-coded Morton box (morton.h:43-58) where its keys are valid.
    Returns (Mtri/s list per repeat, info dict)."""
    import numpy as np
    mg = importlib.import_module(f"{PKG}.meshgen")
    from oracle import refcd, cdoracle  # bench.py's CPU legs are one of the places allowed to run oracle/
    kind, kw, _ = WORKLOADS[name]
    if kind == "soup":
        xyz, idx = mg.soup(sample_tris, seed=kw["seed"], origin=(0.1, -0.4, -0.3), extent=(2.8, 0.6, 2.2))
        sample = f"{sample_tris}-triangle soup, same generator/density as {name}, inside the reference Morton box"
    elif kind == "two_sheets":
        nq = max(8, int(round((sample_tris / 4) ** 0.5)))
        xyz, idx = mg.two_sheets(nq, seed=kw["seed"])
        o, e = np.array(mg.REF_ORIGIN), np.array(mg.REF_EXTENT)
        xyz = (o + 0.05 * e + xyz.astype(np.float64) * (0.9 * e.min())).astype(np.float32)  # uniform scale into the box
        sample = f"two_sheets nq={nq} ({len(idx)} triangles), uniformly scaled into the reference Morton box"
    elif kind == "cloth_fold":
        s = max(8, int(round((sample_tris / 2) ** 0.5)))
        xyz, idx = mg.cloth_fold(s, s)
        sample = f"cloth_fold {s}x{s} ({len(idx)} triangles)"
    else:
        s = max(8, int(round((sample_tris / 2) ** 0.5)))
        xyz, idx = mg.flag(s, s)
        sample = f"flag {s}x{s} ({len(idx)} triangles)"
    n = len(idx)
    vals, stages = [], None
    if refcd.available():
        kind_s = "reference"
        for _ in range(repeats):
            m = refcd.RefMesh.from_arrays(xyz, idx)   # centroid + morton3D + thrust::sort_by_key (load_obj.h:89-107)
            m.build()                                 # fillLeafNodesCpu, generateHierarchyParallelCpu, calBoundingBoxCpu
            npairs = len(m.collide(threads))          # findCollisionIterativeCpu per sorted leaf
            t = m.timing()
            total_ms = t["load"] + t["fill"] + t["hierarchy"] + t["refit"] + t["query"]
            vals.append(n / total_ms / 1e3)
            stages = dict(t, pairs=npairs)
            m.close()
        how = ("reference host functions via oracle/ref_driver.cu; build stages serial as in the reference "
               f"(cpu.cuh:103,125,171), query loop over {threads} thread(s)")
    else:
        kind_s = "port"
        threads = 1
        op = cdoracle.default_params()
        for _ in range(repeats):
            pairs, tm = cdoracle.run(xyz, idx, op)
            total_ms = tm.ms_morton + tm.ms_sort + tm.ms_hierarchy + tm.ms_refit + tm.ms_query
            vals.append(n / total_ms / 1e3)
            stages = dict(morton=tm.ms_morton, sort=tm.ms_sort, hierarchy=tm.ms_hierarchy, refit=tm.ms_refit,
                          query=tm.ms_query, pairs=int(tm.pairs))
        how = "oracle/cd_oracle.c (plain-C restatement), one thread"
    return vals, {"kind": kind_s, "cores": threads, "sample": sample, "how": how, "stages_ms": stages,
                  "host_cores_available": os.cpu_count(), "triangles": n}


def run_reference_arm(args, workload):
    """The reference's own CPU functions on the box's host cores, all threads in the query loop. Each step is a bounded
    SAMPLE of the workload (same generator, same contact density): the sample size is chosen from a short calibration run
    so that steps + warmup fit `--ref-budget` seconds (whole workload when that fits). The line says what it timed:
    config.triangles is the SAMPLE's triangle count, config.sample_of names the full workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm runs on rank 0 alone
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    mg = importlib.import_module(f"{PKG}.meshgen")
    nverts, ntris = workload_sizes(mg, workload)
    t0 = time.time()
    cal_n = 1 << 18
    tc = time.time()
    cal, _ = cpu_reference_run(workload, cal_n, threads, repeats=1)
    per_tri = (time.time() - tc) / cal_n                       # wall seconds per triangle incl. mesh generation
    budget = max(20.0, float(args.ref_budget))
    want = budget / max(total, 1) / per_tri / 1.25              # 25 % head-room (log N growth of the tree walk)
    sample_tris = ntris if want >= ntris else max(1 << 16, 1 << int(want).bit_length() - 1)
    vals, info = cpu_reference_run(workload, sample_tris, threads, repeats=total)
    timed = vals[args.warmup:] or vals
    v = len(timed) / sum(1.0 / x for x in timed)  # total triangles / total time over the timed steps
    n_timed = int(info.get("triangles", sample_tris))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e-3 * n_timed / v, 3) if v else None,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "triangles": n_timed,
                   "sample_of": {"workload": workload, "triangles": ntris, "vertices": nverts},
                   "full_workload": bool(n_timed == ntris),
                   "note": "Mtri/s of a CPU BVH pipeline falls slowly (log N) with size: a sample slightly flatters the CPU arm"},
        "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"], "how": info["how"], "stages_ms": info["stages_ms"],
                         "host_cores_available": info["host_cores_available"]},
        "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.time() - t0, 1),
        "calibration": {"triangles": cal_n, "mtri_per_s": round(cal[0], 4)},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm

def pairs_checksum(words):
    """two wrap-around sums over the sorted packed pair list (int64 words = hi_id << 32 | lo_id): equal lists give equal
    sums, so the records of different GPU counts can be compared without shipping the lists"""
    if words is None or words.numel() == 0:
        return [0, 0]
    w = words.to("cuda") if not words.is_cuda else words
    return [int(w.sum().item()) & 0xFFFFFFFFFFFFFFFF, int((w * w + (w >> 7)).sum().item()) & 0xFFFFFFFFFFFFFFFF]


def device_value(cd, mg, mgpu, ctx, name, steps, warmup=3, host=None):
    """triangles / CUDA-event time of `steps` build + query steps of workload `name` on ONE GPU, mesh resident in HBM.
    host = (xyz_ptr, idx_ptr): the workload's arrays already sit in pinned host memory"""
    import numpy as np
    import torch
    _, _, box = WORKLOADS[name]
    nverts, ntris = workload_sizes(mg, name)
    if host is None:
        xyz, xyz_ptr = cd.pinned_array((nverts, 3), np.float32)
        idx, idx_ptr = cd.pinned_array((ntris, 3), np.uint32)
        generate(mg, name, out=(xyz, idx))
    else:
        xyz_ptr, idx_ptr = host
    params = cd.make_params(*box) if box else cd.default_params()
    mesh = ctx.mesh_from_host_ptr(xyz_ptr, nverts, idx_ptr, ntris)
    dev = torch.device("cuda", ctx.device)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)  # the events below are recorded on torch's current stream
    bvh = ctx.bvh_build(mesh, params)

    def one_step():  # ONE GPU, whatever process group exists: rebuild + whole query + sorted list
        ctx.bvh_rebuild(bvh, mesh, params)
        ptr, count = ctx.self_collide_device(bvh, sorted=True)
        return mgpu.device_pairs_as_tensor(ptr, count, dev)
    for _ in range(warmup):
        pairs = one_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        pairs = one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = ctx.stats()
    csum = pairs_checksum(pairs)
    out = {"workload": name, "triangles": ntris, "vertices": nverts, "n_gpus": 1, "steps": steps,
           "value": round(ntris / ms / 1e3, 2), "unit": UNIT, "ms_per_step": round(ms, 4), "pairs": int(pairs.numel()),
           "pairs_checksum": csum, "pairs_match_oracle": matches_oracle(name, int(pairs.numel()), csum),
           "bvh_build_ms": round(st["ms_build"], 4), "query_ms": round(st["ms_query"], 4),
           "sort_passes": int(st["sort_passes"])}
    bvh.destroy()
    mesh.destroy()
    if host is None:
        cd.host_free(xyz_ptr)
        cd.host_free(idx_ptr)
    return out


def run_gpu_arm(args, workload):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The collision path has no CPU fallback "
                         "(use --impl reference for the CPU arm).")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        # (a short collective timeout: a rank that dies must not leave the others waiting for ten minutes)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    cd = importlib.import_module(f"{PKG}.binding")
    mg = importlib.import_module(f"{PKG}.meshgen")
    mgpu = importlib.import_module(f"{PKG}.multigpu")

    kind, kw, box = WORKLOADS[workload]
    nverts, ntris = workload_sizes(mg, workload)
    # the mesh a caller would hand us: page-locked host arrays (every rank holds the whole mesh)
    xyz, xyz_ptr = cd.pinned_array((nverts, 3), np.float32)
    idx, idx_ptr = cd.pinned_array((ntris, 3), np.uint32)
    generate(mg, workload, out=(xyz, idx))
    params = cd.make_params(*box) if box else cd.default_params()

    ctx = cd.Context(local_rank)
    # the strong-scaling denominator, measured in THIS run: rank 0 times a few single-GPU steps of the same workload
    # before the distributed objects exist (the other ranks wait at the barrier)
    scaling_base = None
    if world > 1 and not args.no_scaling_base:
        if rank == 0:
            scaling_base = device_value(cd, mg, mgpu, ctx, workload, steps=5, host=(xyz_ptr, idx_ptr))
        dist.barrier()
    mesh = ctx.mesh_from_host_ptr(xyz_ptr, nverts, idx_ptr, ntris)
    partitioned = world > 1 and args.mode in ("partitioned", "partitioned-py")
    cxx_step = partitioned and args.mode == "partitioned"
    if cxx_step:
        # each rank owns one Morton range; the whole step is ONE library call per rank (b200cd_dist_step, csrc/dist.cu)
        # (pipelined frames: rank 0 sorts frame k's gathered list on a side stream while all ranks start frame k + 1;
        #  --sync-sort keeps that sort on the main stream)
        prunner = mgpu.DistSelfCollision(cd, ctx, mesh, params, async_sort=not args.sync_sort)
        bvh = prunner.dist.bvh()

        class _Step:
            def step(self, bvh_, mesh_, params_):
                return prunner.step()

            def wait(self):  # the last step's list is complete in stream order after this
                prunner.wait_sorted()
        runner = _Step()
    elif partitioned:
        # the same algorithm orchestrated from Python over torch.distributed (round 1; kept for A/B and as the NCCL
        # send/recv fallback where CUDA IPC is unavailable)
        prunner = mgpu.PartitionedSelfCollision(cd, ctx, mesh, params, peer_memory=not args.no_peer_memory)
        bvh = prunner.part.bvh

        class _Step:
            counts = property(lambda self: prunner.counts)

            def step(self, bvh_, mesh_, params_):
                return prunner.step()
        runner = _Step()
    else:
        runner = mgpu.ShardedSelfCollision(cd, ctx, chunk=args.chunk)  # binds the library to torch's current stream
        if world > 1 and args.replicate == "broadcast":  # rank 0 alone builds; the others receive the BVH over NCCL
            runner.broadcaster = mgpu.BvhBroadcaster(cd, ctx, ntris)
            bvh = ctx.bvh_build(mesh, params) if rank == 0 else ctx.bvh_alloc_like(ntris)
        else:
            bvh = ctx.bvh_build(mesh, params)

    if not hasattr(runner, "wait"):
        runner.wait = lambda: None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the candidate / pair buffers, so timed steps never reallocate)
    merged = None
    for _ in range(max(args.warmup, 3)):
        merged = runner.step(bvh, mesh, params)
    runner.wait()
    npairs_total = int(merged.numel()) if rank == 0 else 0
    checksum = pairs_checksum(merged) if rank == 0 else None
    host_pairs = torch.empty(max(npairs_total, 1) + 1024, dtype=torch.int64).pin_memory() if rank == 0 else None
    # count_launch() is process-wide; sample before/after the timed region
    launches0 = ctx.stats()["kernel_launches"]

    stage_keys = ("ms_morton", "ms_sort", "ms_hierarchy", "ms_refit", "ms_build", "ms_traverse", "ms_narrow",
                  "ms_pair_sort", "ms_query")
    acc = {k: 0.0 for k in stage_keys}
    last = {}

    # ---- timed region 1: mesh resident in HBM
    sampler = ClockSampler(physical_gpu_index(local_rank))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        merged = runner.step(bvh, mesh, params)
        st = ctx.stats()  # per-stage CUDA-event times of this step (events on the same stream)
        for k in stage_keys:
            acc[k] += st[k]
        last = st
    runner.wait()  # the final step's sorted list is part of the timed work
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.finish()
    launches = ctx.stats()["kernel_launches"] - launches0

    # ---- timed region 2: end to end through the host-buffer calls
    # every step: H2D of that step's mesh from pinned host memory, build, query, D2H of the sorted pair list
    h2d = 12 * nverts + 12 * ntris
    d2h = 0

    def e2e_serial_step():
        nonlocal merged, d2h
        if world == 1:
            mesh.update_from_ptr(xyz_ptr, idx_ptr)        # b200cd_mesh_update: H2D of vertices + indices
            ctx.bvh_rebuild(bvh, mesh, params)            # b200cd_bvh_rebuild
            cnt = ctx.self_collide_into(bvh, host_pairs.data_ptr(), host_pairs.numel(), sorted=True)  # + D2H
            d2h = 8 * cnt
        else:                                             # 1/world of the mesh per PCIe link + NVLink all-gather
            mgpu.upload_mesh_sharded(ctx, mesh, xyz_ptr, idx_ptr)
            merged = runner.step(bvh, mesh, params)
            runner.wait()
            if rank == 0:
                host_pairs[: merged.numel()].copy_(merged, non_blocking=True)
                d2h = 8 * int(merged.numel())
            torch.cuda.current_stream().synchronize()

    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record()
    for _ in range(args.steps):
        e2e_serial_step()
    ev3.record()
    barrier()
    ms_e2e_serial = ev2.elapsed_time(ev3)
    ms_e2e = ms_e2e_serial
    e2e_mode = "serial calls: b200cd_mesh_update, b200cd_bvh_rebuild, b200cd_self_collide (one frame at a time)"
    ms_e2e_verts = None  # the same stream of frames when only the VERTICES change (a deforming mesh: static topology)
    if world == 1:
        # double-buffered frames (b200cd_mesh_update_async / b200cd_mesh_wait): the H2D of frame k+1 runs on the
        # copy stream while frame k is built and queried from the other mesh object. Still one H2D of the whole
        # mesh and one D2H of the pair list per step, all inside the timed region.
        frames = [mesh, ctx.mesh_from_host_ptr(xyz_ptr, nverts, idx_ptr, ntris)]
        for k in range(2):                                # warm-up: allocates the second staging buffer, events, copy stream
            frames[k].update_async_from_ptr(xyz_ptr, idx_ptr)
            frames[k].wait()

        def stream_frames(index_ptr):
            nonlocal d2h
            e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e_a.record()
            frames[0].update_async_from_ptr(xyz_ptr, index_ptr)
            for k in range(args.steps):
                cur = frames[k % 2]
                if k + 1 < args.steps:
                    frames[(k + 1) % 2].update_async_from_ptr(xyz_ptr, index_ptr)
                cur.wait()
                ctx.bvh_rebuild(bvh, cur, params)
                cnt = ctx.self_collide_into(bvh, host_pairs.data_ptr(), host_pairs.numel(), sorted=True)
                d2h = 8 * cnt
            e_b.record()
            barrier()
            return e_a.elapsed_time(e_b)
        ms_e2e = stream_frames(idx_ptr)
        ms_e2e_verts = stream_frames(None)
        e2e_mode = ("double-buffered frames: b200cd_mesh_update_async of frame k+1 overlaps b200cd_bvh_rebuild + "
                    "b200cd_self_collide of frame k; one full H2D and one D2H per step inside the timed region")
        frames[1].destroy()
    elif partitioned and (cxx_step or prunner.peer_memory):
        # several GPUs, double-buffered frames: every rank pushes its 1/world slice of frame k+1 over its own PCIe
        # link and on into the peers' mesh buffers with the copy engines over NVLink (copy stream, no NCCL, no SMs)
        # while frame k is built and queried. Still one whole mesh H2D (summed over the ranks) and one D2H per step.
        frames = [mesh, ctx.mesh_from_host_ptr(xyz_ptr, nverts, idx_ptr, ntris)]
        pm = mgpu.PeerMeshFrames(cd, ctx, frames)
        if pm.ok:
            for k in range(2):                            # warm-up: staging buffers, events, copy stream, peer mappings
                pm.upload_async(k, xyz_ptr, idx_ptr)
                pm.wait(k)
            torch.cuda.synchronize()

            def set_mesh(m):
                if cxx_step:
                    prunner.mesh = m
                else:
                    prunner.part.mesh = m

            def stream_frames(index_ptr):
                nonlocal merged, d2h
                e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e_a.record()
                pm.upload_async(0, xyz_ptr, index_ptr)
                for k in range(args.steps):
                    if k + 1 < args.steps:
                        pm.upload_async((k + 1) % 2, xyz_ptr, index_ptr)
                    set_mesh(pm.wait(k % 2))
                    merged = prunner.step()
                    runner.wait()
                    if rank == 0:
                        host_pairs[: merged.numel()].copy_(merged, non_blocking=True)
                        d2h = 8 * int(merged.numel())
                    torch.cuda.current_stream().synchronize()
                e_b.record()
                barrier()
                return e_a.elapsed_time(e_b)
            ms_e2e = stream_frames(idx_ptr)
            ms_e2e_verts = stream_frames(None)
            e2e_mode = ("double-buffered frames on every rank: 1/N of frame k+1 per PCIe link, then copy-engine pushes into the "
                        "peers' mesh buffers over NVLink (b200cd_mesh_update_slice_async), overlapping build + query of frame k; "
                        "one whole-mesh H2D (summed over ranks) and one D2H per step inside the timed region")
            set_mesh(mesh)
            torch.cuda.synchronize()
            dist.barrier()
            pm.close()
            dist.barrier()                                # nobody frees a frame another rank still has mapped
        frames[1].destroy()

    if args.trace:  # three more (untimed) steps with an event behind every launch: per-kernel device timeline of every rank
        barrier()
        ctx.trace_enable(True)
        for _ in range(3):
            runner.step(bvh, mesh, params)
        runner.wait()
        torch.cuda.synchronize()
        ctx.trace_enable(False)
        ctx.trace_dump(f"{args.trace}.rank{rank}.csv")
        barrier()

    # ---- max over ranks
    t = torch.tensor([ms_total, ms_e2e, ms_e2e_serial, ms_e2e_verts or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_e2e_serial = float(t[0]), float(t[1]), float(t[2])
    ms_e2e_verts = float(t[3]) if ms_e2e_verts else None

    if partitioned and not cxx_step:
        prunner.step(profile=True)  # one extra, untimed step (all ranks) with a synchronise after every phase
    pstats = dict(prunner.stats) if partitioned else {}
    all_pstats = None
    if partitioned and cxx_step:  # every rank's view of the last step (range size, ghosts received, phase times)
        all_pstats = [None] * world
        dist.all_gather_object(all_pstats, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in pstats.items()})

    if rank == 0:
        K = args.steps
        ms_step = ms_total / K
        value = ntris / ms_step / 1e3
        e2e_value = ntris / (ms_e2e / K) / 1e3
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        ncand, npairs = int(last.get("candidates", 0)), int(last.get("pairs", 0))
        recs = os.environ.get("B200CD_RECS", "1") != "0"
        recs = recs and 2 * nverts >= 3 * ntris  # api.cu run_build: only for (mostly) unshared vertices
        quant = uses_quantised_nodes(ntris, nverts)
        abytes = algorithmic_bytes(ntris, nverts, ncand, npairs, int(last.get("sort_passes", 8)), recs, quant)
        nloc = ntris
        if partitioned:  # rank 0's own Morton range (keys arrive from the exchange: no face-ordered records)
            nloc = int(pstats.get("local_triangles", ntris // world))
            abytes = algorithmic_bytes(nloc, min(nverts, 3 * nloc), ncand, npairs, int(last.get("sort_passes", 8)), False, quant)
        elif world > 1:  # per-rank share of the query stages
            abytes["traverse"] = (32 if quant else 64) * ntris + 64 * ntris // world + 8 * ncand
        stages = {}
        for s, key in (("morton", "ms_morton"), ("sort", "ms_sort"), ("tree", "ms_refit"),
                       ("traverse", "ms_traverse"), ("narrow", "ms_narrow")):
            ms = acc[key] / K
            gbs = abytes[s] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            kname = "broad_kernel_q (quantised nodes) (+ entry_kernel)" if (s == "traverse" and quant) else STAGE_KERNEL[s]
            stages[s] = {"kernel": kname, "ms": round(ms, 4), "algorithmic_bytes": int(abytes[s]),
                         "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        if partitioned:  # K1 ran on the rank's input slice before the exchange: not part of the local build's events
            stages["morton"] = {"kernel": STAGE_KERNEL["morton"], "ms": None, "note": "runs before the key exchange (phase_ms: keys+hist+allreduce)"}
        dominant = max(("tree", "traverse", "narrow") if partitioned else ("morton", "tree", "traverse", "narrow"),
                       key=lambda s: stages[s]["ms"])
        sort_launch_ms = stages["sort"]["ms"] / max(int(last.get("sort_passes", 8)), 1)
        if sort_launch_ms > stages[dominant]["ms"]:
            dominant = "sort"
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        traffic_note = None
        traffic_file = json.load(open(tpath)) if os.path.exists(tpath) else {}
        if traffic_file:
            traffic = traffic_file.get(workload, {}).get(dominant)
            if traffic is not None and world > 1:
                # ncu may not wrap a multi-rank run: the single-GPU capture of the same workload, divided by the ranks
                # (each rank traverses 1/N of the leaves over 1/N of the nodes)
                traffic = int(traffic / world)
                traffic_note = f"single-GPU ncu capture of {workload} / {world} ranks (a multi-rank run cannot go under ncu)"
        # the same stages against SURVEY.md section 8(d)'s CONTRACT bytes (the survey's planned layout, not ours): a fatter
        # layout of ours cannot raise these fractions
        cb = contract_bytes(nloc, min(nverts, 3 * nloc) if partitioned else nverts, npairs)
        cms = {"morton": acc["ms_morton"] / K, "sort": acc["ms_sort"] / K, "tree": acc["ms_refit"] / K,
               "query": (acc["ms_traverse"] + acc["ms_narrow"]) / K, "e2e": (acc["ms_build"] + acc["ms_query"]) / K}
        contract = {}
        for st_, ms in cms.items():
            gbs = cb[st_] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            contract[st_] = {"bytes": int(cb[st_]), "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        contract_stage = {"traverse": "query", "narrow": "query"}.get(dominant, dominant)
        roofline = {"bound": "hbm", "kernel": stages[dominant]["kernel"], "stage": dominant,
                    "achieved": stages[dominant]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": stages[dominant]["frac"], "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                    "launch_ms": stages[dominant]["ms"], "algorithmic_bytes": stages[dominant]["algorithmic_bytes"],
                    "stages": stages,
                    "frac_contract": contract[contract_stage]["frac"],
                    "contract": dict(contract, note="SURVEY.md 8(d) bytes: K1 48N+16V, K2 200N, K3+K4 132N, K5 (traverse+narrow) "
                                                    "100N+16V+8*pairs, e2e 480N+32V+8*pairs; `frac_contract` is the dominant "
                                                    "kernel's stage (traverse and narrow share K5)"),
                    # what actually bounds the traversal (ncu, profiles/r02_ncu_broad_l1.md): the L1 data pipe, not HBM
                    "traverse_l1_data_pipe": traffic_file.get("_l1_data_pipe", {}).get(workload) if dominant == "traverse" else None,
                    "e2e_frac_contract": contract["e2e"]["frac"],
                    "e2e_algorithmic_bytes": int(sum(abytes.values())),
                    "e2e_frac": round(sum(abytes.values()) / ((acc["ms_build"] + acc["ms_query"]) / K * 1e-3) / 1e9 / peak, 4)
                    if acc["ms_build"] + acc["ms_query"] > 0 else None}
        trav_ms = acc["ms_traverse"] / K
        traversal = {"nodes_visited_per_query": round(last.get("nodes_visited", 0) / max(ntris // world, 1), 2),
                     "gnodes_per_s": round(last.get("nodes_visited", 0) / (trav_ms * 1e-3) / 1e9, 2) if trav_ms > 0 else None,
                     "lane_utilisation": round(last.get("nodes_visited", 0) / max(32 * last.get("warp_steps", 0), 1), 3),
                     "candidates": ncand, "pairs": npairs}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "triangles": ntris, "vertices": nverts,
                       "morton_box": "unit cube" if box else "reference constants (morton.h:45,51,57)",
                       "key_bits": 63, "pairs": npairs_total, "pairs_checksum": checksum,
                       "pairs_match_oracle": matches_oracle(workload, npairs_total, checksum),
                       "sort_passes": int(last.get("sort_passes", 0)),
                       "sort_regime": "hybrid: radix passes over the top digits + per-run fix-up (steady state of this mesh)"
                       if 0 < int(last.get("sort_passes", 0)) < 8 else "full: every digit",
                       "parallelism": "single GPU" if world == 1 else (
                           f"partitioned x{world}, one library call per rank and step (b200cd_dist_step): one Morton range per "
                           f"rank; histograms, (key, id) pairs, coarse boxes, ghost records and the pair lists go through NVLink "
                           f"peer memory (CUDA IPC) ordered by flag barriers; local tree + query; sort on rank 0" if cxx_step else
                           f"partitioned x{world} (Python orchestration): one Morton range per rank; (key, id) exchange and ghost records "
                           f"{'stored straight into the owners buffers over NVLink peer memory (CUDA IPC)' if prunner.peer_memory else 'over grouped NCCL send/recv'}; "
                           f"local tree + query; one padded NCCL all-gather of the pair lists + sort on rank 0" if partitioned else
                           f"query-sharded x{world}, replicated BVH ({'built on rank 0, broadcast with NCCL (b200cd_dist_broadcast_bvh)' if args.replicate == 'broadcast' else 'built on every rank'}), block-cyclic chunks of {args.chunk} sorted leaves, "
                           f"NCCL gather + sort on rank 0"),
                       "l2_policy": "inputs larger than L2 (mesh + BVH >> 126 MB); no explicit flush"},
            "bvh_build_ms": round(acc["ms_build"] / K, 4), "query_ms": round(acc["ms_query"] / K, 4),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(ms_e2e / K, 4), "mode": e2e_mode,
                    "serial_value": round(ntris / (ms_e2e_serial / K) / 1e3, 2),
                    "serial_ms_per_step": round(ms_e2e_serial / K, 4),
                    "h2d_gbs": round(h2d / (ms_e2e / K * 1e-3) / 1e9, 1),
                    # NOT the headline: a deforming mesh keeps its topology, so a caller streams vertex positions only
                    # (12 B x V per frame; the index array was uploaded once, outside the timed region)
                    "vertex_frames": None if not ms_e2e_verts else {
                        "value": round(ntris / (ms_e2e_verts / K) / 1e3, 2), "ms_per_step": round(ms_e2e_verts / K, 4),
                        "h2d_bytes_per_step": 12 * nverts}},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "traversal": traversal,
        }
        if partitioned:
            line["partition"] = dict(pstats, rank=0, build_ms=ctx.stats()["ms_build"])
            if cxx_step:  # device times of rank 0's phases of the LAST timed step (CUDA events, barriers included)
                line["partition"]["phase_ms"] = {k[3:]: round(v, 4) for k, v in pstats.items() if k.startswith("ms_")}
                line["partition"]["ranks"] = [{"rank": r["rank"], "local_triangles": r["local_triangles"], "ghosts_received": r["ghosts"],
                                               "local_pairs": r["local_pairs"], "ms_build": r["ms_build"],
                                               "ms_local_query": r["ms_local_query"], "ms_step": r["ms_step"]} for r in all_pstats]
        if scaling_base is not None:
            line["scaling_base"] = scaling_base
            line["speedup_same_workload"] = round(scaling_base["ms_per_step"] / ms_step, 3)
        if world == 1 and workload != SCALING_WORKLOAD and not args.no_scaling_base:
            # the N > 1 runs use sheets64m (strong scaling of a fixed 2^26-triangle mesh): its single-GPU time, measured
            # here in the same run, is the denominator of that scaling curve (this line's `value` is soup16m)
            line["scaling_base"] = device_value(cd, mg, mgpu, ctx, SCALING_WORKLOAD, steps=max(3, min(args.steps, 5)))
        if world == 1 and box is None and not args.no_gpu_reference:
            # the reference's OWN GPU kernels (bvh.cuh:125,146,258, collision.cuh:73 with main.cu:92-142's launch grids),
            # compiled for sm_100a from /root/reference into oracle/_ref, on this same B200 and mesh: needs a mesh inside the
            # reference's hard-coded Morton box with unique codes (flag1m, cloth1m). Its Morton codes and sort run on the HOST
            # (load_obj.h:91,107) and are not in these four numbers.
            try:
                from oracle import refcd
                if refcd.available():
                    rm = refcd.RefMesh.from_arrays(xyz, idx)
                    rp, rms = rm.gpu_run(repeats=5)
                    host_ms = rm.timing()["load"]
                    rm.close()
                    rsum = sum(rms.values())
                    line["gpu_reference_baseline"] = {
                        "what": "reference CUDA kernels, unmodified, sm_100a build, same GPU, same mesh; minimum of 5 runs",
                        "stages_ms": {k: round(v, 4) for k, v in rms.items()}, "gpu_ms": round(rsum, 4),
                        "mtri_per_s_gpu_stages_only": round(ntris / rsum / 1e3, 2),
                        "host_morton_and_sort_ms": round(host_ms, 1),
                        "pairs": int(len(rp)), "pairs_equal_ours": bool(len(rp) == npairs_total),
                        "ours_ms_per_step": round(ms_step, 4), "ours_over_reference_gpu_stages": round(rsum / ms_step, 2)}
            except Exception as e:  # noqa: BLE001
                line["gpu_reference_baseline"] = {"unavailable": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            vals, info = cpu_reference_run(workload, args.cpu_sample, 1)
            line["cpu_baseline"] = {"value": round(vals[0], 4), "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"], "how": info["how"], "stages_ms": info["stages_ms"],
                                    "host_cores_available": info["host_cores_available"]}
        print(json.dumps(line), flush=True)

    if partitioned:
        torch.cuda.synchronize()
        dist.barrier()
        prunner.close()
    elif world > 1 and getattr(runner, "broadcaster", None) is not None:
        runner.broadcaster.close()
    if not cxx_step:
        bvh.destroy()
    mesh.destroy()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="partitioned", choices=["partitioned", "partitioned-py", "replicated"],
                    help="N > 1: one Morton range per rank through b200cd_dist_step (default), the same orchestrated from Python "
                         "over torch.distributed, or replicated BVH with sharded queries")
    ap.add_argument("--replicate", default="rebuild", choices=["rebuild", "broadcast"],
                    help="--mode replicated: every rank builds the (deterministic) BVH, or rank 0 builds and NCCL broadcasts it")
    ap.add_argument("--sync-sort", action="store_true",
                    help="N > 1: rank 0 sorts the gathered pair list on the main stream instead of a side stream (no frame pipelining)")
    ap.add_argument("--no-peer-memory", action="store_true",
                    help="partitioned mode: exchange (key, id) and ghosts with NCCL send/recv instead of peer-memory stores")
    ap.add_argument("--chunk", type=int, default=1 << 14, help="sorted leaves per block-cyclic query chunk (N > 1)")
    ap.add_argument("--cpu-sample", type=int, default=1 << 21, help="triangles in the cpu_baseline sample")
    ap.add_argument("--trace", default=None, help="write a per-kernel device timeline of 3 extra steps to PATH.rank<r>.csv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="flag1m / cloth1m: skip the reference's own GPU kernels")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work for all steps together")
    ap.add_argument("--no-scaling-base", action="store_true", help="N = 1: skip the single-GPU run of the N > 1 workload")
    args = ap.parse_args()
    workload = args.workload or ("soup16m" if args.gpus == 1 else SCALING_WORKLOAD)

    if args.impl == "reference":
        run_reference_arm(args, workload)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched by hand: re-exec under torchrun, one rank per GPU (the driver does this itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
        raise SystemExit(subprocess.call(cmd + sys.argv[1:]))
    run_gpu_arm(args, workload)


if __name__ == "__main__":
    main()
