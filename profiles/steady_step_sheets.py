"""one build + two rebuilds + query of the 16 M-triangle two-sheet mesh (the N > 1 workload at a quarter of its size)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
ctx = cd.Context(0)
xyz, idx = mg.two_sheets(2048)
p = cd.make_params((0, 0, 0), (1, 1, 1))
mesh = ctx.mesh_from_arrays(xyz, idx)
bvh = ctx.bvh_build(mesh, p)
for _ in range(2):
    ctx.synchronize()
    ctx.bvh_rebuild(bvh, mesh, p)
ptr, cnt = ctx.self_collide_device(bvh, sorted=True)
print(cnt, ctx.stats())
