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


def test_reference_arm_maps_none_of_the_gpu_libraries():
    """the CPU arm must not load the product's CUDA code (nor the facade that links it): frames come from
    libyouth_synth.so (plain C), the configuration from the oracle"""
    code = ("import sys, bench; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; "
            "bench.main(); maps = open('/proc/self/maps').read(); "
            "sys.stderr.write('MAPPED:' + ','.join(sorted({l.split('/')[-1] for l in maps.splitlines() if '.so' in l and "
            "('youth' in l or 'AlgorithmModule' in l or 'libcuda' in l)})) + '\\n')")
    env = dict(os.environ)
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    mapped = [l for l in r.stderr.splitlines() if l.startswith("MAPPED:")][-1][len("MAPPED:"):].split(",")
    assert "libyouth_synth.so" in mapped and any(m.startswith("libyouth_oracle") for m in mapped), mapped
    assert not any(m.startswith(("libyouth_cuda", "libAlgorithmModule", "libcuda")) for m in mapped), mapped


def test_both_arms_describe_the_same_configuration():
    import bench

    class A:
        pass

    a = A()
    a.user_ppt, a.mode = 0, "frame"
    ppt = bench.default_ppt(a, 300, 1)
    ours = bench.config_dict(640, 480, 3, 1, 300, ppt, 1, "frame")
    assert ppt == 128 and "host_placement" not in ours  # nothing arm-specific inside `config`
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert json.loads(r.stdout.strip())["config"] == ours


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


def test_roofline_arithmetic_is_reproducible_by_hand(monkeypatch):
    """roofline.frac = compulsory DRAM bytes per launch / CUDA-event launch time / measured peak, with 299 pairs
    (300 frames touched) per 300-frame launch; the 48 B/px/iteration the kernel requests is reported beside it; the
    ncu traffic is only attached for the geometry AND the kernel sources it was captured with."""
    import numpy as np

    import bench

    class B:  # the profiling class ids of slam-rgbd_b200/binding.py
        PROF_INGEST, PROF_NORMALS, PROF_ICP0, PROF_SOLVE, PROF_MISC, PROF_RAYCAST, PROF_CLASSES = 0, 1, 2, 6, 7, 8, 9

    ms = np.zeros(9)
    n = np.zeros(9, dtype=np.int64)
    ms[B.PROF_ICP0], n[B.PROF_ICP0] = 5.0, 10       # 0.5 ms per level-0 launch
    ms[B.PROF_ICP0 + 1], n[B.PROF_ICP0 + 1] = 0.75, 5
    ms[B.PROF_INGEST], n[B.PROF_INGEST] = 1.8, 1
    ms[B.PROF_NORMALS], n[B.PROF_NORMALS] = 0.25, 1
    monkeypatch.setattr(bench, "measured_peak", lambda: (6547.5, "measured (test)"))
    out = bench.rooflines(B, ms, n, 640, 480, 3, 1, [(0, 300)], 300, 128, "frame")
    top = out["roofline"]
    comp = 24 * 640 * 480 * 300  # every frame of the group once
    assert top["compulsory_bytes_per_launch"] == comp and top["pairs_per_launch"] == 299
    assert top["requested_bytes_per_launch"] == 48 * 640 * 480 * 299
    assert abs(top["frac"] - comp / 0.5e-3 / 1e9 / 6547.5) < 1e-12 and 0.6 < top["frac"] < 0.72
    assert top["frac"] <= 1.0 and top["peak"] == 6547.5 and top["bound"] == "hbm"
    kinds = [r["kernel"] for r in out["roofline_all"]]
    assert kinds == ["k_icp (level 0)", "k_icp (level 1)", "k_ingest + k_normals"]
    ing = out["roofline_all"][-1]
    b_pre = 2 * 640 * 480 + 24 * (640 * 480 + 320 * 240 + 160 * 120)
    assert ing["algorithmic_bytes_per_frame"] == b_pre and abs(ing["achieved"] - b_pre * 300 / 2.05e-3 / 1e9) < 1e-6
    # two groups of 150: 149 + 150 pairs, 150 + 151 frames touched
    out2 = bench.rooflines(B, ms, n, 640, 480, 3, 1, [(0, 150), (150, 150)], 150, 128, "frame")
    assert out2["roofline"]["pairs_per_launch"] == 149.5
    assert out2["roofline"]["compulsory_bytes_per_launch"] == 24 * 640 * 480 * 150.5
    # traffic: attached for the captured geometry, refused when the kernel sources differ
    t, src = bench.load_traffic("640x480_L3_S1_B300_ppt128_frame")
    assert t.get("k_icp_L0", 0) > 2.2e9 and "captured at" in src
    assert bench.load_traffic("641x480_L3_S1_B300_ppt128_frame")[0] == {}
    monkeypatch.setattr(bench, "kernels_sha1", lambda: "0" * 40)
    t, src = bench.load_traffic("640x480_L3_S1_B300_ppt128_frame")
    assert t == {} and "other kernel sources" in src
