"""Full-size expected results for the bench workloads -> tests/golden/checksums.json.

Run in the build container only (needs oracle/_ref/libref_cd.so, i.e. /root/reference, and ~30 GB of RAM for
the 2^26-triangle mesh; the GPU box never runs this script):

    python tests/golden/make_checksums.py [workload ...]        # default: every workload of bench.py

For every workload the colliding-pair list is computed TWICE on the CPU and both must agree:

  (1) "reference": the reference's own host functions behind oracle/ref_driver.cu -
      thrust::sort_by_key (load_obj.h:107), fillLeafNodesCpu (cpu.cuh:89), generateHierarchyParallelCpu
      (cpu.cuh:110), calBoundingBoxCpu (cpu.cuh:167), findCollisionIterativeCpu (cpu.cuh:196) with
      checkBoxOverlap / neighborCount / checkTriangleContactHelper - over all host threads (the queries are
      independent). The sort keys are NOT morton3D's: its box is hard-coded for the flag mesh
      (morton.h:43-58), a unit-cube mesh leaves it (UB at morton.h:80) and 16-64 M triangles give duplicate
      60-bit codes, for which the reference builds a malformed tree. Each triangle's key is its position in
      the oracle's Morton order instead (unique, strictly increasing after the sort), which gives the
      reference hierarchy a valid, spatially coherent binary tree of depth <= 27 (its traversal stack holds
      32 entries, cpu.cuh:198). The emitted pair SET does not depend on the tree.
  (2) "oracle": oracle/cd_oracle.c, the plain-C restatement, whole pipeline with the workload's own Morton
      box, one thread.

The JSON holds, per workload: triangles, vertices, pairs, checksum (gpu-computing-course_b200/pairsum.py)
and which of the two computations produced / confirmed it. bench.py compares every run against it
(`config.pairs_match_oracle`), tests/test_gpu_scale.py asserts it.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cdoracle as co, refcd  # noqa: E402
import bench  # noqa: E402  (WORKLOADS, generate)

mg = importlib.import_module("gpu-computing-course_b200.meshgen")
ps = importlib.import_module("gpu-computing-course_b200.pairsum")

OUT = os.path.join(HERE, "checksums.json")


def one(name, threads):
    _, _, box = bench.WORKLOADS[name]
    t0 = time.time()
    xyz, idx = bench.generate(mg, name)
    n = len(idx)
    params = co.make_params(*box) if box else co.default_params()
    # position of every triangle in the oracle's Morton order (ties: face order, as load_obj.h:107's stable sort)
    keys = co.morton_keys(xyz, idx, params)
    _, sids = co.sort_keys(keys)
    pos = np.empty(n, np.uint64)
    pos[sids] = np.arange(n, dtype=np.uint64)
    dup = int(n - len(np.unique(keys)))
    del keys, sids
    print(f"[{name}] {n} triangles, {len(xyz)} vertices, {dup} duplicate Morton keys, mesh+keys {time.time() - t0:.1f} s",
          flush=True)

    t1 = time.time()
    m = refcd.RefMesh.from_arrays_keyed(xyz, idx, pos)
    wrong_parent = m.build()
    assert wrong_parent == 0
    rp = co.sort_pairs(m.collide(threads))
    tm = m.timing()
    m.close()
    assert np.all(rp[:, 0] < rp[:, 1])
    assert len(rp) < 2 or np.all((rp[1:, 0] > rp[:-1, 0]) | ((rp[1:, 0] == rp[:-1, 0]) & (rp[1:, 1] > rp[:-1, 1]))), \
        "reference pair list is not duplicate-free"
    ref_sum = ps.pairs_checksum_np(rp)
    print(f"[{name}] reference functions: {len(rp)} pairs, checksum {ref_sum}, {time.time() - t1:.1f} s "
          f"(stages ms {tm})", flush=True)

    t2 = time.time()
    op, otm = co.run(xyz, idx, params)
    ora_sum = ps.pairs_checksum_np(op)
    print(f"[{name}] oracle restatement: {len(op)} pairs, checksum {ora_sum}, {time.time() - t2:.1f} s", flush=True)
    same = bool(np.array_equal(op, rp))
    assert same, f"{name}: oracle restatement and reference functions disagree"
    return {"triangles": n, "vertices": int(len(xyz)), "pairs": int(len(rp)), "checksum": ref_sum,
            "duplicate_morton_keys": dup,
            "source": "reference host functions (oracle/ref_driver.cu, position keys) == oracle/cd_oracle.c, full lists compared",
            "reference_query_threads": threads}


def main():
    assert refcd.available(), "build oracle/_ref first (make -C oracle)"
    names = sys.argv[1:] or ["soup1m", "cloth1m", "flag1m", "sheets16m", "soup16m", "sheets64m"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    threads = os.cpu_count() or 1
    for nm in names:
        res[nm] = one(nm, threads)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
            f.write("\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
