"""CPU: the library's multi-threaded OBJ parser (csrc/obj_parse.cu, b200cd_obj_parse_host) against the
reference's own loadObj (load_obj.h:24-103, through oracle/_ref when it is built) and against the
arrays that were written. No GPU needed: parsing is host code."""
import os

import numpy as np
import pytest

from oracle import refcd


def write(path, text):
    with open(path, "w", newline="") as f:
        f.write(text)
    return path


def test_round_trip_of_generated_meshes(cd, mg, tmp_path):
    for name, (xyz, idx) in (("flag", mg.flag(120, 90)), ("cloth", mg.cloth_fold(64, 64)), ("soup", mg.soup(5000, seed=2))):
        p = os.path.join(tmp_path, name + ".obj")
        mg.write_obj(p, xyz, idx)
        x, i = cd.parse_obj(p)
        assert x.dtype == np.float32 and i.dtype == np.uint32
        assert np.array_equal(x, xyz) and np.array_equal(i, idx), name


@pytest.mark.skipif(not refcd.available(), reason="oracle/_ref/libref_cd.so not built")
def test_same_mesh_as_reference_loadobj_including_quirks(cd, tmp_path, capfd):
    # comments, vt lines, blank lines, CRLF, tabs, signs, exponents, trailing junk, and a last line without '\n'
    text = ("# comment\n"
            "v 0.25 -0.125 1e-3\r\n"          # (faces must stay inside the reference's Morton box, morton.h:43-58,78)
            "v +1.5   0.1875\t-0.0625 extra tokens are ignored\n"
            "v\t9 9 9\n"                       # 'v' + TAB is not a vertex line (load_obj.h:48 wants "v ")
            "vt 0.5 0.5\n"
            "\n"
            "v 1.0e+0 .5 5.\n"
            "v 0.1 0.2 0.3\n"
            "vn 0 0 1\n"
            "f 1/1 2/2 3/3\n"
            "f  4/1   1/7\t2/2 trailing\n"
            "g group\n"
            "f 3/1 2/1 1/1\n"
            "f 1/1 2/1 4/1")          # no trailing newline: dropped (load_obj.h:41)
    p = write(os.path.join(tmp_path, "q.obj"), text)
    x, i = cd.parse_obj(p)
    m = refcd.RefMesh.from_obj(p)
    capfd.readouterr()
    rx, ri = m.mesh()
    m.close()
    assert np.array_equal(x, rx) and np.array_equal(i, ri)
    assert x.shape == (4, 3) and i.tolist() == [[0, 1, 2], [3, 0, 1], [2, 1, 0]]
    assert x[1].tolist() == [1.5, 0.1875, -0.0625]


@pytest.mark.skipif(not refcd.available(), reason="oracle/_ref/libref_cd.so not built")
def test_float_parsing_matches_scanf_bit_for_bit(cd, tmp_path, capfd):
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.normal(0, 1, 3000), rng.uniform(-1e-6, 1e-6, 600), rng.uniform(-1e6, 1e6, 600),
                           [0.0, -0.0, 1e-45, 3.4e38, 1e-39, 0.1, 1 / 3]])
    vals = vals[: len(vals) // 3 * 3].reshape(-1, 3)
    lines = ["v 1 0 0.5\n", "v 2 0.1 1\n", "v 1.5 -0.2 0.25\n"]  # the one face stays inside the reference's Morton box
    for k, (a, b, c) in enumerate(vals):
        fmt = ("v %.17g %.17g %.17g\n", "v %.9e %.3f %.12f\n", "v %g %g %g\n")[k % 3]  # more digits than a float holds
        lines.append(fmt % (a, b, c))
    lines.append("f 1/1 2/1 3/1\n")
    p = write(os.path.join(tmp_path, "f.obj"), "".join(lines))
    x, _ = cd.parse_obj(p)
    m = refcd.RefMesh.from_obj(p)
    capfd.readouterr()
    rx, _ = m.mesh()
    m.close()
    assert np.array_equal(x.view(np.uint32), rx.view(np.uint32))  # same bits, including -0.0 and denormals


def test_errors_are_statuses_with_line_numbers_not_exit(cd, tmp_path):
    cases = [
        ("v 0 0\n", cd.E_PARSE, "line 1"),                                   # load_obj.h:60 exits
        ("v 0 0 0\nv 1 1 1\nv 2 2 x\n", cd.E_PARSE, "line 3"),
        ("v 0 0 0\nf 1 2 3\n", cd.E_PARSE, "line 2"),                        # load_obj.h:73 exits
        ("v 0 0 0\nv 1 0 0\nf 1/1 2/1 3/1\nv 0 1 0\n", cd.E_PARSE, "line 3"),  # vertex 3 is defined AFTER the face
        ("v 0 0 0\nf 0/1 1/1 1/1\n", cd.E_PARSE, "line 2"),                  # indices are 1-based
        ("v 0 0 0\n#" + "x" * 300 + "\n", cd.E_PARSE, "line 2"),             # getline(buffer, 255) would never recover
    ]
    for text, status, where in cases:
        p = write(os.path.join(tmp_path, "bad.obj"), text)
        with pytest.raises(cd.B200cdError) as e:
            cd.parse_obj(p)
        assert e.value.status == status and where in str(e.value), (text[:20], str(e.value))
    with pytest.raises(cd.B200cdError) as e:
        cd.parse_obj(os.path.join(tmp_path, "missing.obj"))
    assert e.value.status == cd.E_IO


def test_empty_and_vertex_only_files(cd, tmp_path):
    x, i = cd.parse_obj(write(os.path.join(tmp_path, "e.obj"), ""))
    assert x.shape == (0, 3) and i.shape == (0, 3)
    x, i = cd.parse_obj(write(os.path.join(tmp_path, "v.obj"), "v 1 2 3\nv 4 5 6"))
    assert x.tolist() == [[1.0, 2.0, 3.0]] and i.shape == (0, 3)


def test_first_error_wins_across_parser_chunks(cd, mg, tmp_path):
    """a multi-megabyte file is parsed in parallel chunks; the FIRST bad line must be the one reported"""
    xyz, idx = mg.flag(300, 300)
    p = os.path.join(tmp_path, "big.obj")
    mg.write_obj(p, xyz, idx)
    lines = open(p).read().split("\n")
    nv = len(xyz)
    bad1, bad2 = nv + 1000, nv + 150000  # two broken face lines far apart
    lines[bad1] = "f 1 2 3"
    lines[bad2] = "f x"
    write(p, "\n".join(lines))
    with pytest.raises(cd.B200cdError) as e:
        cd.parse_obj(p)
    assert f"line {bad1 + 1}:" in str(e.value)
