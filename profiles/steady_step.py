"""one build + two rebuilds + query of the 2^lg soup (steady state of the adaptive sort) - the ncu launch list's workload"""
import importlib, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = cd.Context(0)
xyz, idx = mg.soup(1 << lg, seed=1234)
p = cd.make_params((0, 0, 0), (1, 1, 1))
mesh = ctx.mesh_from_arrays(xyz, idx)
bvh = ctx.bvh_build(mesh, p)
for _ in range(2):
    ctx.synchronize()
    ctx.bvh_rebuild(bvh, mesh, p)
ptr, cnt = ctx.self_collide_device(bvh, sorted=True)
print(cnt, ctx.stats())
