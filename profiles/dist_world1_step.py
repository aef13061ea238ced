"""three b200cd_dist_step calls with ONE rank on the 16 M-triangle two-sheet mesh - the multi-GPU step's own kernels
(slice keys, histogram, push, plan, fused partition + exchange, coarse boxes, gather) under ncu, which may not wrap a
multi-rank run; with one rank every peer store lands in local memory"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cd = importlib.import_module("gpu-computing-course_b200.binding")
mg = importlib.import_module("gpu-computing-course_b200.meshgen")
ctx = cd.Context(0)
xyz, idx = mg.two_sheets(int(sys.argv[1]) if len(sys.argv) > 1 else 2048)
p = cd.make_params((0, 0, 0), (1, 1, 1))
mesh = ctx.mesh_from_arrays(xyz, idx)
dist = ctx.dist_create(0, 1, mesh.ntris)
for _ in range(3):
    ptr, count = dist.step(mesh, p)
ctx.synchronize()
print(count, dist.stats())
