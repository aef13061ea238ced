"""CPU: the C-ABI library loads and exports exactly what include/b200cd.h declares; argument and
no-device error behaviour. No compute calls (this file runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200cd.h")


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200cd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(cd):
    declared = header_functions()
    assert len(declared) >= 30
    assert sorted(cd.SYMBOLS) == declared, "binding.SYMBOLS out of date with include/b200cd.h"
    lib = cd.lib()
    for s in declared:
        assert hasattr(lib, s), f"libb200cd.so does not export {s}"


def test_no_other_symbols_leak(cd):
    out = subprocess.check_output(["nm", "-D", "--defined-only", cd.LIB_PATH], text=True)
    names = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [n for n in names if not n.startswith("_") and not n.startswith("cuda") and not n.startswith("__")]
    assert sorted(ours) == header_functions()


def test_library_is_sm100a_only(cd):
    out = subprocess.run(["cuobjdump", "-lelf", cd.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_product_does_not_link_or_import_the_oracle(cd):
    ldd = subprocess.check_output(["ldd", cd.LIB_PATH], text=True)
    assert "oracle" not in ldd and "ref_cd" not in ldd
    pkg = os.path.join(ROOT, "gpu-computing-course_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "cd_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_status_strings_and_defaults(cd):
    lib = cd.lib()
    assert lib.b200cd_abi_version() == 3
    assert lib.b200cd_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.b200cd_strerror(cd.E_NODEVICE)
    p = cd.default_params()
    # reference morton.h:45,51,57
    assert list(p.morton_origin) == [0.004501, -0.476622, -0.381965]
    assert list(p.morton_extent) == [3.08, 0.76, 2.36]
    assert p.key_bits == 63 and p.auto_box == 0


def test_create_fails_loudly_without_device_or_with_bad_args(cd):
    lib = cd.lib()
    assert lib.b200cd_create(C.c_int(0), None) == cd.E_INVALID
    h = C.c_void_p()
    rc = lib.b200cd_create(C.c_int(0), C.byref(h))
    if rc == cd.OK:  # a GPU is present (GPU box): a bad ordinal is still rejected
        assert lib.b200cd_create(C.c_int(4096), C.byref(C.c_void_p())) == cd.E_INVALID
        lib.b200cd_destroy(h)
    else:  # CPU container: no fallback, creation fails with a status (never exit(), book.h:21-31)
        assert rc == cd.E_NODEVICE and not h.value
        try:
            cd.Context(0)
            raise AssertionError("Context() must raise without a GPU")
        except cd.B200cdError as e:
            assert e.status == cd.E_NODEVICE
    # NULL handles are rejected, not dereferenced
    assert lib.b200cd_get_stats(None, None) == cd.E_INVALID
    assert lib.b200cd_mesh_info(None, None, None) == cd.E_INVALID
    assert lib.b200cd_destroy(None) == cd.OK


def test_cxx_example_is_built_and_fails_loudly_without_a_gpu(cd):
    """examples/ref_main.cpp (the reference's main() on the C ABI) links libb200cd.so; no GPU -> a status, not exit()"""
    exe = os.path.join(ROOT, "gpu-computing-course_b200", "lib", "b200cd_run")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    assert os.path.exists(exe)
    assert "libb200cd.so" in subprocess.check_output(["ldd", exe], text=True)
    h = C.c_void_p()
    if cd.lib().b200cd_create(C.c_int(0), C.byref(h)) == cd.OK:
        cd.lib().b200cd_destroy(h)
        return  # GPU box: tests/test_gpu_example.py runs it for real
    out = subprocess.run([exe, "whatever.obj"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr
    dexe = os.path.join(ROOT, "gpu-computing-course_b200", "lib", "b200cd_dist_run")   # examples/dist_main.cpp: forked ranks
    assert os.path.exists(dexe) and "libb200cd.so" in subprocess.check_output(["ldd", dexe], text=True)
    out = subprocess.run([dexe, "whatever.obj", "--ranks", "2"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and out.stderr.count("no CPU fallback") == 2
