import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "gpu-computing-course_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _pkg(name):
    return importlib.import_module(f"{PKG}.{name}")


@pytest.fixture(scope="session")
def mg():
    """synthetic workload generators"""
    return _pkg("meshgen")


@pytest.fixture(scope="session")
def cd():
    """ctypes binding of libb200cd.so"""
    return _pkg("binding")


@pytest.fixture(scope="session")
def co():
    """CPU oracle (plain-C restatement) — checker only"""
    from oracle import cdoracle
    cdoracle.lib()
    return cdoracle


@pytest.fixture(scope="session")
def ctx(cd):
    """one context on cuda:0 shared by the GPU tests"""
    c = cd.Context(0)
    yield c
    c.destroy()
