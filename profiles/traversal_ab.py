"""A/B of traversal builds (profiles/r02_traversal_variants.md): one subprocess per library build (B200CD_LIB_PATH, no torch).
usage: python profiles/traversal_ab.py v0,v5@B200CD_REFILL=16 soup16m,sheets64m [repeats]  - libraries are looked up as
scratch/var/libb200cd_<name>.so (built with make -C gpu-computing-course_b200/csrc EXTRA_NVFLAGS=-D... and copied there)"""
import importlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    cd = importlib.import_module("gpu-computing-course_b200.binding")
    mg = importlib.import_module("gpu-computing-course_b200.meshgen")
    ps = importlib.import_module("gpu-computing-course_b200.pairsum")
    gold = json.load(open(os.path.join(ROOT, "tests/golden/checksums.json")))
    ctx = cd.Context(0)
    out = {}
    for name in sys.argv[2:]:
        if name == "soup16m": xyz, idx = mg.soup(1 << 24, seed=1234); reps = 12
        elif name == "sheets64m": xyz, idx = mg.two_sheets(4096); reps = 8
        elif name == "sheets16m": xyz, idx = mg.two_sheets(2048); reps = 10
        elif name == "cloth1m": xyz, idx = mg.cloth_fold(); reps = 20
        p = cd.default_params() if name == "cloth1m" else cd.make_params((0, 0, 0), (1, 1, 1))
        mesh = ctx.mesh_from_arrays(xyz, idx)
        bvh = ctx.bvh_build(mesh, p)
        for _ in range(3):
            ctx.bvh_rebuild(bvh, mesh, p); ctx.self_collide_device(bvh, sorted=True)
        tr, na, tot = [], [], []
        for _ in range(reps):
            ctx.bvh_rebuild(bvh, mesh, p)
            ptr, cnt = ctx.self_collide_device(bvh, sorted=True)
            st = ctx.stats()
            tr.append(st["ms_traverse"]); na.append(st["ms_narrow"]); tot.append(st["ms_build"] + st["ms_query"])
        pairs = ctx.self_collide(bvh, sorted=True)
        g = gold[name]
        c = ps.pairs_checksum_np(pairs)
        ok = (len(pairs) == g["pairs"]) and [int(x) for x in c] == [int(x) for x in g["checksum"]]
        out[name] = dict(traverse=float(np.median(tr)), narrow=float(np.median(na)), step=float(np.median(tot)), tmin=float(np.min(tr)),
                         pairs=len(pairs), ok=bool(ok), cand=int(st["candidates"]), visits=int(st["nodes_visited"]), steps=int(st["warp_steps"]))
        bvh.destroy(); mesh.destroy()
    print("RESULT " + json.dumps(out), flush=True)
    sys.exit(0)
variants = sys.argv[1].split(",")
workloads = sys.argv[2].split(",")
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    for v in variants:
        env = dict(os.environ)
        name, _, extra = v.partition("@")
        env["B200CD_LIB_PATH"] = os.path.join(ROOT, "scratch/var", f"libb200cd_{name}.so")
        if extra:
            k, _, val = extra.partition("=")
            env[k] = val
        t0 = time.time()
        r = subprocess.run([sys.executable, __file__, "child"] + workloads, env=env, capture_output=True, text=True, timeout=200)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        print(v, f"{time.time()-t0:.0f}s", line[0][7:] if line else ("FAILED " + r.stderr[-600:]), flush=True)
