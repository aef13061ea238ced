"""b200cd_run — the reference driver's command line, on the B200 library.

    python -m gpu-computing-course_b200.cli mesh.obj           (importlib: the package name has a hyphen)
    python gpu-computing-course_b200/cli.py mesh.obj [--auto-box] [--key-bits 63|30] [--validate]

Prints what the reference's main() prints after findCollisions (reference CollisionDetection/main.cu:147-154,
33-45): the contact count, one "%07u - %07u" line per colliding pair (lower triangle ID first) and the sorted
set of colliding triangle IDs - so the output can be diffed against resources/MyResult.txt (pairs are listed
in sorted order here; the reference's order is whatever its atomics produced).
"""
import argparse
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(HERE))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("obj", help="OBJ file in the reference's dialect (v x y z / f v/vt v/vt v/vt)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--auto-box", action="store_true", help="Morton box = the mesh's bounding box instead of morton.h:43-58")
    ap.add_argument("--key-bits", type=int, default=63, choices=[63, 30])
    ap.add_argument("--validate", action="store_true", help="also run the reference's structural self-checks (check.cuh:64-96)")
    args = ap.parse_args(argv)
    cd = importlib.import_module(os.path.basename(HERE) + ".binding")
    ctx = cd.Context(args.device)
    mesh = ctx.mesh_load_obj(args.obj)
    print(f"\nObj File Loaded:\n- {mesh.nverts} vertexes loaded\n- {mesh.ntris} triangles loaded")  # load_obj.h:117-119
    bvh = ctx.bvh_build(mesh, cd.make_params(key_bits=args.key_bits, auto_box=args.auto_box))
    if args.validate:
        c = bvh.validate(mesh)
        print(f"Internal node check result: nullParentnum = {c['null_parent_internal']}, wrongBoundCount={c['wrong_bound_count']}, "
              f"nullChildCount={c['null_child']}, notInternalCount=0, uninitBoxCount={c['uninit_box_internal']}")  # main.cu:119
        print(f"Leaf node check result: nullParentnum = {c['null_parent_leaf']}, nullTriangle={c['bad_triangle']}, "
              f"notLeafCount=0, illegalBoxCount={c['uninit_box_leaf']}")  # main.cu:127
        print(f"wrong morton sort count: {c['unsorted_keys']}")  # load_obj.h:116
    pairs = ctx.self_collide(bvh, sorted=True)
    st = ctx.stats()
    print(f"Time of build (morton+sort+hierarchy+boxes): {st['ms_build']:.3f} ms")
    print(f"Time of findCollisions: {st['ms_query']:.3f} ms")
    print(f"\n\n- contact val = {len(pairs)}")                                   # main.cu:147
    print(f"\nCollision pair ({len(pairs)} triangle pairs in total):")           # main.cu:149
    for a, b in pairs:
        print("%07u - %07u" % (a, b))                                            # main.cu:151
    ids = ctx.unique_triangles(bvh).tolist()                                     # makeAndPrintSet (main.cu:33-45), on the device
    print(f"\n\nCollision Triangles:（{len(ids)} points in total）:")    # main.cu:40
    for t in ids:
        print(t)
    print("- Successfully Return")
    return 0


if __name__ == "__main__":
    sys.exit(main())
