"""b200cd — B200-native triangle-mesh self-collision (LBVH build + query).

The directory name carries a hyphen, so import it with
    importlib.import_module("gpu-computing-course_b200")
(see tests/conftest.py, bench.py, __graft_entry__.py).
"""
