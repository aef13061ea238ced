"""GPU: the compiled reference-side binding (examples/ref_main.cpp -> lib/b200cd_run) - the reference's main()
(main.cu:47-174) with its body replaced by C-ABI calls - run on a golden mesh through the OBJ route; its printed pair
list and triangle-ID set must be the ones the REFERENCE's own host functions produced (tests/golden/flag_40x40.npz,
cloth_20x20.npz), in the reference's print format (main.cu:147-154, 33-45)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "gpu-computing-course_b200", "lib", "b200cd_run")


@pytest.mark.parametrize("name", ["flag_40x40", "cloth_20x20"])
def test_cxx_binding_prints_the_reference_result(mg, tmp_path, name):
    assert os.path.exists(EXE), "examples/ref_main.cpp has not been built (__graft_entry__.build())"
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    obj = tmp_path / (name + ".obj")
    mg.write_obj(str(obj), g["xyz"], g["idx"])
    out = subprocess.run([EXE, str(obj), "--validate"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    text = out.stdout
    assert f"- {len(g['xyz'])} vertexes loaded" in text and f"- {len(g['idx'])} triangles loaded" in text
    assert "Internal node check result: nullParentnum = 1, wrongBoundCount=0, nullChildCount=0, notInternalCount=0, uninitBoxCount=0" in text
    assert "wrong morton sort count: 0" in text
    pairs = np.array([[int(a), int(b)] for a, b in re.findall(r"^(\d{7}) - (\d{7})$", text, re.M)], np.uint32).reshape(-1, 2)
    assert f"- contact val = {len(g['pairs'])}" in text
    assert np.array_equal(pairs, g["pairs"])
    tail = text.split("Collision Triangles:")[1]
    ids = np.array([int(x) for x in re.findall(r"^(\d+)$", tail, re.M)], np.uint32)
    assert np.array_equal(ids, np.unique(g["pairs"]))
    assert text.rstrip().endswith("- Successfully Return")


def test_cxx_binding_reports_errors_as_statuses(tmp_path):
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/1 2/2 9/9\n")       # forward / out-of-range vertex reference
    out = subprocess.run([EXE, str(bad)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "b200cd_mesh_load_obj" in out.stderr    # the reference exit()s (load_obj.h:34,60,73)
    out = subprocess.run([EXE, str(tmp_path / "missing.obj")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "file could not be read" in out.stderr


DIST_EXE = os.path.join(ROOT, "gpu-computing-course_b200", "lib", "b200cd_dist_run")


@pytest.mark.parametrize("name,ranks", [("flag_40x40", 2), ("cloth_20x20", 3)])
def test_cxx_multi_rank_driver_prints_the_reference_result(mg, tmp_path, name, ranks):
    """examples/dist_main.cpp: the multi-GPU step driven from plain C++ - one forked process per rank, export blobs over
    pipes, b200cd_dist_step - with all ranks pinned to GPU 0 (so it runs on any box); same printed result as one GPU"""
    assert os.path.exists(DIST_EXE), "examples/dist_main.cpp has not been built (__graft_entry__.build())"
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    obj = tmp_path / (name + ".obj")
    mg.write_obj(str(obj), g["xyz"], g["idx"])
    env = dict(os.environ, B200CD_BARRIER_TIMEOUT_MS="20000")
    out = subprocess.run([DIST_EXE, str(obj), "--ranks", str(ranks), "--steps", "2", "--device", "0"], capture_output=True, text=True,
                         timeout=180, env=env)
    assert out.returncode == 0, out.stderr
    text = out.stdout
    assert f"- {len(g['idx'])} triangles loaded" in text and f"{ranks} ranks" in text
    pairs = np.array([[int(a), int(b)] for a, b in re.findall(r"^(\d{7}) - (\d{7})$", text, re.M)], np.uint32).reshape(-1, 2)
    assert f"- contact val = {len(g['pairs'])}" in text
    assert np.array_equal(pairs, g["pairs"])
    ids = np.array([int(x) for x in re.findall(r"^(\d+)$", text.split("Collision Triangles:")[1], re.M)], np.uint32)
    assert np.array_equal(ids, np.unique(g["pairs"]))
