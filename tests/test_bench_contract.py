"""bench.py's output contract, checked on the CPU through the reference arm (`--impl reference` needs no GPU): one JSON line with the
keys the driver reads, the CPU baseline object and the informational reference-build leg."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-persons", "2", "--timepoints", "12"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "person-timesteps/s"
    assert d["metric"].startswith("SR-UKF person-timesteps/sec") and d["dtype"] == "f64" and d["data"] == "synthetic"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "persons" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    rbld = cb["reference_build"]                      # the reference's own C++ through the stand-in headers, or why it is unavailable
    assert ("value" in rbld and rbld["m2ll_vs_oracle_rel"] < 1e-9 and rbld["cores"] == 1) or "unavailable" in rbld


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
