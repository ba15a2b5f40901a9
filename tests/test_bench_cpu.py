"""CPU: the reference arm of bench.py (the one arm that runs without a GPU) prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pretrain_clips_per_s" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["vs_baseline"] is None and line["data"] == "synthetic"
    cb = line["cpu_baseline"]
    # "reference" = the real reference's STFTLearner.pretrain_epoch (its tree is importable in the build container), "port" = the oracle (GPU box)
    assert cb["kind"] == ("reference" if os.path.isdir("/root/reference/code") else "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "clips" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "micro-batch 256/GPU" in line["config"]["workload"]          # the same workload name as our arm


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
