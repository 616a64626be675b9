"""The JSON lines bench.py printed on the B200 (committed under profiles/) carry every key of the measurement
contract (bench.py docstring / DESIGN.md section 6); and without a GPU the product arm refuses to run while the
reference arm falls back to the CPU port."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return [json.loads(l) for l in f if l.strip()]


def test_committed_bench_lines_follow_the_contract():
    ours = [d for d in _lines("r1j_bench_n1.jsonl") + _lines("r1k_bench_unary_final.jsonl") if d.get("impl") != "reference"]
    assert ours
    for d in ours:
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert k in d, k
        assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
        assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32"
        assert d["config"]["workload"] and d["warmup"] >= 3 and d["gpu_launches"] > 0
        e = d["e2e"]
        assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert 0 < e["value"] <= 1.05 * d["value"]          # copies inside the timed region never make it faster
        r = d["roofline"]
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r)
        assert 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        c = d["clocks"]
        assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                             "sw_thermal_slowdown"}
    final = _lines("r1k_bench_unary_final.jsonl")[-1]
    cb = final["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(cb) and cb["kind"] == "port" and cb["cores"] >= 1
    refs = [d for d in _lines("r1j_bench_n1.jsonl") if d.get("impl") == "reference"]
    assert len(refs) == 3
    for d in refs:
        assert d["e2e"] == dict(value=d["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
        assert d["cpu_baseline"]["kind"] == "reference"


def test_bench_arms_without_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0
