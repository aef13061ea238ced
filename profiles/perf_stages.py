import importlib, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
import numpy as np
ctx = cd.Context(0)
sizes = [int(a) for a in sys.argv[1:]] or [18, 20, 22, 24]
reps = 4
for lg in sizes:
    if lg == 0:
        t0 = time.time(); xyz, idx = mg.cloth_fold(); name = "cloth"; p = cd.default_params()
    elif lg == 1:
        t0 = time.time(); xyz, idx = mg.two_sheets(2048); name = "sheets2048"; p = cd.make_params((0,0,0),(1,1,1))
    else:
        n = 1 << lg
        t0 = time.time(); xyz, idx = mg.soup(n, seed=1234); name = f"soup2^{lg}"; p = cd.make_params((0,0,0),(1,1,1))
    tg = time.time() - t0
    mesh = ctx.mesh_from_arrays(xyz, idx)
    bvh = ctx.bvh_build(mesh, p)
    ctx.synchronize()
    for r in range(reps):
        ctx.bvh_rebuild(bvh, mesh, p)
        ptr, cnt = ctx.self_collide_device(bvh, sorted=True)
        st = ctx.stats()
        keys = ("ms_morton", "ms_sort", "ms_hierarchy", "ms_refit", "ms_build", "ms_traverse", "ms_narrow", "ms_pair_sort", "ms_query")
        print(name, f"gen {tg:.1f}s", "n", st["ntris"], "cand", st["candidates"], "pairs", st["pairs"], "retries", st["query_retries"], "passes", st["sort_passes"],
              " ".join(f"{k[3:]}={st[k]:.3f}" for k in keys),
              f"visits/q={st['nodes_visited']/st['ntris']:.1f} lane_util={st['nodes_visited']/max(1,32*st['warp_steps']):.2f} entries/warp={st['start_entries']/(st['ntris']/32):.1f}", flush=True)
    bvh.destroy(); mesh.destroy()
