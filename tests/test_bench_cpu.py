"""bench.py without a GPU: the reference arm (the CPU path, oracle port) prints the contract's JSON line; the product arm
must fail loudly -- there is no CPU fallback to time."""
import json
import os
import subprocess
import sys

import matrixfactorizationsgd.java_b200 as mf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-sample", "300000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                     # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sgd_rating_updates_per_sec" and d["unit"] == "updates/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert "netflix-shaped" in d["config"]["workload"] and "k=128" in d["config"]["workload"]


def test_product_arm_fails_loudly_without_a_gpu():
    if mf.device_count() > 0:
        import pytest
        pytest.skip("a GPU is visible")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-e2e", "--no-cpu"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]          # and no number is printed
