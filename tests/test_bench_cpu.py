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


def test_reference_arm_keeps_to_its_time_budget():
    """A host too slow for the whole run shrinks the LATER steps to a prefix of the records and says so; value stays records / time."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--ref-sample", "4000000", "--ref-budget-s", "0.5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
    assert d["config"]["same_workload_as_product_arm"] is False
    assert "budget" in d["cpu_baseline"]["sample"] and "prefix" in d["cpu_baseline"]["sample"]
    assert d["value"] > 0 and d["cpu_baseline"]["update_loops_only"] >= d["value"]


def test_product_arm_fails_loudly_without_a_gpu():
    if mf.device_count() > 0:
        import pytest
        pytest.skip("a GPU is visible")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-e2e", "--no-cpu"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]          # and no number is printed


def test_reference_arm_does_not_map_the_product_library():
    """The reference process loads only oracle/ (round-1 review: it used to import the GPU package for WORKLOADS/SEED)."""
    code = ("import sys, os; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-sample', '100000'];"
            "import runpy; runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'libmfsgd' not in maps, 'product library mapped'; assert 'liboracle' in maps"
            % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]


def test_cpu_baseline_ml100k_runs_the_oracle_cli():
    sys.path.insert(0, ROOT)
    import bench
    d = bench.cpu_baseline_ml100k()
    assert d is not None and d["host_cores"] >= 1
    modes = [(r["mode"], r["threads"]) for r in d["runs"] if "error" not in r]
    assert ("seq", 1) in modes and any(m == "threads" for m, _ in modes)
    seq = [r for r in d["runs"] if r.get("mode") == "seq"][0]
    assert abs(seq["heldout_rmse"] - 0.393763) < 1e-5          # tests/golden/oracle_rmse_ml100k.json, last epoch


def test_oracle_curve_covers_the_bench_runs_epoch_counts():
    """bench.py prints rmse_vs_oracle only when the committed sequential-oracle curve reaches warm-up + steps epochs: the
    Netflix-shaped fixture (the bench workload) covers the default 3 + 30 and the driver's shorter runs, for one GPU and
    for rings of 2, 4 and 8 (the DSGD-ordered sequential curves)."""
    import pytest
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_rmse_netflix.json")))
    if len(fx["heldout_rmse_per_epoch"]) <= 10:
        pytest.skip("the fixture holds the 10-epoch curves only (tools/oracle_reference_rmse.py netflix 35 --dsgd 2,4,8 extends them)")
    assert len(fx["heldout_rmse_per_epoch"]) >= 33
    for G in (2, 4, 8):
        assert len(fx["dsgd%d" % G]["heldout_rmse_per_epoch"]) >= 33
    curve = fx["heldout_rmse_per_epoch"]
    assert abs(curve[9] - 0.389229263310776) < 1e-12          # the first ten epochs are the values round 1 committed
    assert all(0.38 < v < 0.41 for v in curve) and curve[-1] < curve[9]
