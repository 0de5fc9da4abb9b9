"""Native multi-GPU driver (slam-rgbd_b200/host/multi_gpu.c -> bin/youth_multi; SURVEY.md section 8(e), section 4
tier 5): one host pthread + one tracker handle per GPU in one process, sequences sharded over the GPUs, one
ncclAllGather of the trajectories.  The files written from GPU 0's copy of the gathered buffer must be bit-identical
(same text) to what a single GPU writes for the same sequences, and equal to tracking through the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def run_multi(pkg, gpus, spg, frames, prefix):
    exe = os.path.join(os.path.dirname(pkg.lib_paths()["harness"]), "youth_multi")
    assert os.path.exists(exe), "bin/youth_multi not built"
    r = subprocess.run([exe, str(gpus), str(spg), str(frames), prefix], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_one_gpu_two_sequences_matches_the_c_abi(pkg, tmp_path):
    from slam_rgbd_b200.binding import Tracker

    n = 8
    info = run_multi(pkg, 1, 2, n, str(tmp_path / "one"))
    assert info["gpus"] == 1 and info["frames_per_sec"] > 0
    seqs = [pkg.synth_sequence(n, sequence=s) for s in range(2)]
    trk = Tracker(pkg.default_config(batch=n, n_streams=2, traj_capacity=n))
    want = trk.track_batch(seqs)
    trk.close()
    for s in range(2):
        rows = np.loadtxt(str(tmp_path / f"one_seq{s:03d}_trajectory.txt"))
        assert rows.shape == (n, 8)
        assert np.allclose(rows[:, 1:4], want[s][:, [3, 7, 11]], atol=5e-7)
        assert np.allclose(rows[:, 0], 0.033 * np.arange(n))


@pytest.mark.skipif(gpu_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_two_gpus_give_the_single_gpu_files(pkg, tmp_path):
    n = 8
    run_multi(pkg, 1, 2, n, str(tmp_path / "one"))
    info = run_multi(pkg, 2, 1, n, str(tmp_path / "two"))
    assert info["gpus"] == 2
    for s in range(2):
        a = open(tmp_path / f"one_seq{s:03d}_trajectory.txt").read()
        b = open(tmp_path / f"two_seq{s:03d}_trajectory.txt").read()
        assert a == b and len(a.splitlines()) == n, f"sequence {s}: the gathered trajectory differs from the single-GPU one"
