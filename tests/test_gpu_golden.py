"""GPU: the CUDA path, through the C ABI, against the committed golden vectors produced by the
reference's own host functions (tests/golden/make_golden.py). No oracle in the loop here."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIPELINES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not p.endswith("kat.npz"))


def check_against_golden(cd, ctx, mesh, g):
    n = len(g["idx"])
    bvh = ctx.bvh_build(mesh, cd.default_params())  # reference Morton constants, morton.h:45,51,57
    nodes, skeys, sids = bvh.download()
    assert np.array_equal(skeys, g["sorted_keys"])
    assert np.array_equal(sids, g["sorted_ids"])
    assert np.array_equal(nodes["left"][: n - 1], g["left"])
    assert np.array_equal(nodes["right"][: n - 1], g["right"])
    gb = np.concatenate([nodes["lo"], nodes["hi"]], axis=1)
    assert np.array_equal(gb, g["bounds"]), "node bounds: expected 0 ulp"
    chk = bvh.validate(mesh)
    assert chk["null_parent_internal"] == 1 and sum(chk.values()) == 1, chk  # resources/result.png counters
    pairs = ctx.self_collide(bvh, sorted=True)
    assert np.array_equal(pairs, g["pairs"])
    bvh.destroy()


@pytest.mark.parametrize("name", PIPELINES)
def test_golden_from_arrays(cd, ctx, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    mesh = ctx.mesh_from_arrays(g["xyz"], g["idx"])
    check_against_golden(cd, ctx, mesh, g)
    mesh.destroy()


@pytest.mark.parametrize("name", PIPELINES)
def test_golden_through_obj_loader(cd, ctx, mg, tmp_path, name):
    """b200cd_mesh_load_obj reads the reference's OBJ dialect (load_obj.h:48-52,68) to the same mesh"""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    path = os.path.join(tmp_path, name + ".obj")
    mg.write_obj(path, g["xyz"], g["idx"])
    mesh = ctx.mesh_load_obj(path)
    xyz, idx = mesh.download()
    assert np.array_equal(xyz, g["xyz"]) and np.array_equal(idx, g["idx"])
    check_against_golden(cd, ctx, mesh, g)
    mesh.destroy()


def test_obj_loader_quirks_and_errors(cd, ctx, tmp_path):
    p = os.path.join(tmp_path, "q.obj")
    # comment / vt / blank lines are skipped; a last line without '\n' is dropped (load_obj.h:41)
    open(p, "w").write("# c\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\n\nf 1/1 2/1 3/1\nf 1/1 3/1 2/1")
    m = ctx.mesh_load_obj(p)
    assert (m.nverts, m.ntris) == (3, 1)
    m.destroy()
    # the reference exits on these (load_obj.h:60,73,34); the library returns a status
    for text, status in (("v 0 0\n", cd.E_PARSE), ("v 0 0 0\nf 1 2 3\n", cd.E_PARSE), ("v 0 0 0\nf 1/1 2/1 3/1\n", cd.E_PARSE)):
        open(p, "w").write(text)
        with pytest.raises(cd.B200cdError) as e:
            ctx.mesh_load_obj(p)
        assert e.value.status == status
    with pytest.raises(cd.B200cdError) as e:
        ctx.mesh_load_obj(os.path.join(tmp_path, "missing.obj"))
    assert e.value.status == cd.E_IO


def test_cli_prints_the_reference_output_format(cd, co, mg, tmp_path, capsys):
    """python gpu-computing-course_b200/cli.py mesh.obj prints main.cu:147-154's lines: count, %07u pairs, ID set"""
    import importlib
    cli = importlib.import_module("gpu-computing-course_b200.cli")
    g = np.load(os.path.join(GOLDEN, "cloth_20x20.npz"))
    path = os.path.join(tmp_path, "cloth.obj")
    mg.write_obj(path, g["xyz"], g["idx"])
    assert cli.main([path, "--validate"]) == 0
    out = capsys.readouterr().out
    assert f"- contact val = {len(g['pairs'])}" in out
    printed = [tuple(int(x) for x in l.split(" - ")) for l in out.splitlines() if " - " in l and l[:7].isdigit()]
    assert printed == [tuple(p) for p in g["pairs"].tolist()]
    ids = sorted(set(g["pairs"].reshape(-1).tolist()))
    tail = out.split("points in total")[1].splitlines()[1:1 + len(ids)]
    assert [int(t) for t in tail] == ids
    assert "nullParentnum = 1, wrongBoundCount=0, nullChildCount=0" in out
