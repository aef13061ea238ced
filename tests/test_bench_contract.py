"""CPU: bench.py's reference arm prints exactly ONE JSON line with the keys the driver reads.

`bench.py --impl reference` times the reference's own host functions (oracle/_ref when it was built here, else the
plain-C oracle port) on a bounded sample of the workload; it needs no GPU. The GPU arm prints the same keys plus
`roofline`, `clocks`, `gpu_launches` (checked on the GPU box by the driver's own run)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-budget", "20"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mtri/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("Mtri/s end-to-end self-collision") and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["config"]["workload"] == "soup16m" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # the line says what it timed: the SAMPLE's size, and which workload it is a sample of
    assert d["config"]["sample_of"]["triangles"] == 1 << 24 and 0 < d["config"]["triangles"] <= 1 << 24
    assert d["config"]["full_workload"] == (d["config"]["triangles"] == 1 << 24)
    assert abs(d["ms_per_step"] - 1e-3 * d["config"]["triangles"] / d["value"]) < 1e-2 * d["ms_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb and cb["unit"] == "Mtri/s"
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "Mtri/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_non_zero_ranks_of_the_reference_arm_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_stage_bytes_follow_the_layout_the_library_uses(monkeypatch):
    """bench.py's per-stage algorithmic bytes: a soup walks the 32-byte quantised nodes (and its build writes them), a
    mesh the exact 64-byte ones; the knob that overrides the library's choice overrides bench.py's too"""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    n = 1 << 20
    for var in ("B200CD_BROAD_QUANT", "B200CD_TRAVERSAL"):
        monkeypatch.delenv(var, raising=False)
    assert bench.uses_quantised_nodes(n, 3 * n) and not bench.uses_quantised_nodes(n, n // 2)
    exact = bench.algorithmic_bytes(n, 3 * n, 5 * n, n // 8, 4, recs=True, quant=False)
    quant = bench.algorithmic_bytes(n, 3 * n, 5 * n, n // 8, 4, recs=True, quant=True)
    assert exact["traverse"] - quant["traverse"] == 32 * n and quant["tree"] - exact["tree"] == 32 * n
    assert all(exact[k] == quant[k] for k in ("morton", "sort", "narrow"))
    monkeypatch.setenv("B200CD_BROAD_QUANT", "0")
    assert not bench.uses_quantised_nodes(n, 3 * n)
    monkeypatch.setenv("B200CD_BROAD_QUANT", "1")
    assert bench.uses_quantised_nodes(n, n // 2)
    monkeypatch.setenv("B200CD_TRAVERSAL", "1")
    assert not bench.uses_quantised_nodes(n, 3 * n)
    # the survey's contract bytes do not depend on our layout
    assert bench.contract_bytes(n, 3 * n, 7)["query"] == 100 * n + 16 * 3 * n + 8 * 7
