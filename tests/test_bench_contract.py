"""bench.py on a machine without a GPU: the reference arm (the CPU statement on the host cores) prints the
one JSON line the driver parses, with the keys of the contract; our arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_present():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def run_bench(*flags):
    env = dict(os.environ)
    env.pop("RANK", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "icp_tracked_frames_per_sec" and j["unit"] == "frames/s"
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["dtype"] == "f32" and j["data"] == "synthetic"
    assert j["value"] > 0 and j["ms_per_step"] > 0
    # same configuration as our arm: configs[1], and the reduction geometry our arm uses for 300-pair launches
    assert "configs[1]" in j["config"]["workload"] and "model" not in j["config"]
    assert j["config"]["icp_ppt"] == 128 and j["config"]["batch_frames_per_launch_group"] == 300
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(gpu_present(), reason="checks the behaviour WITHOUT a CUDA device")
def test_our_arm_has_no_cpu_fallback():
    r = run_bench("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
