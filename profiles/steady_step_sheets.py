"""one build + two rebuilds + query of the two-sheet mesh: argv[1] = quads per side (2048 -> 16 M triangles, 4096 -> the 2^26-triangle N > 1 workload)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
ctx = cd.Context(0)
xyz, idx = mg.two_sheets(int(sys.argv[1]) if len(sys.argv) > 1 else 2048)
p = cd.make_params((0, 0, 0), (1, 1, 1))
mesh = ctx.mesh_from_arrays(xyz, idx)
bvh = ctx.bvh_build(mesh, p)
for _ in range(2):
    ctx.synchronize()
    ctx.bvh_rebuild(bvh, mesh, p)
ptr, cnt = ctx.self_collide_device(bvh, sorted=True)
print(cnt, ctx.stats())
