"""bench.py's reference arm runs on the CPU alone: check the JSON contract of that line here (no GPU needed)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra):
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-grid", "24", "--cpu-scale", "12"] + list(extra), capture_output=True, text=True, timeout=300,
                       cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys():
    d = run_bench()
    assert d["impl"] == "reference" and d["metric"] == "spmv_effective_bandwidth" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["all_cores"]["kind"] == "port" and cb["all_cores"]["cores"] >= 1 and cb["all_cores"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None


def test_reference_arm_other_workloads():
    for extra in (["--workload", "rmat"], ["--format", "tjds"], ["--workload", "rmat", "--format", "tjds"]):
        d = run_bench(*extra)
        assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--cpu-grid", "10"], capture_output=True, text=True, timeout=120, cwd=REPO, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
