#!/usr/bin/env python
"""Generates tests/golden/golden_ext_160x120.npz from the CPU statements of the two "next" rows
(SURVEY.md section 8(f)): the YD16 codec and frame-to-model tracking.  Same inputs as
make_golden.py (the frames stored in golden_160x120.npz).  Run here, committed with its output;
pins the CPU statements against their own regressions and gives the GPU tests a fixture that does
not depend on re-running them."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import oracle_py as O  # noqa: E402

W, H = 160, 120
TSDF = dict(dim=(64, 32, 64), voxel_m=0.1, origin=(-3.2, -1.6, -1.2), trunc_m=0.3)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    frames = np.load(os.path.join(HERE, "golden_160x120.npz"))["frames"]
    cfg = O.default_config(width=W, height=H, fx=570.3 * W / 640, fy=570.3 * W / 640, cx=W / 2.0, cy=H / 2.0)
    t = O.tsdf_config(**TSDF)
    out = {}
    streams = [O.codec_encode(f) for f in frames]
    out["yd16_stream_0"] = streams[0]
    out["yd16_sizes"] = np.array([len(s) for s in streams], dtype=np.int64)
    out["yd16_sha"] = np.array([sha(s) for s in streams])
    poses, status = O.track_sequence_model(cfg, t, frames)
    out["model_poses"] = poses
    out["model_status"] = status
    # volume and model maps after fusing frame 0 at the identity and ray casting with its own depth as hint
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    f0 = O.OFrame(cfg, frames[0])
    vol = O.tsdf_new(t)
    O.tsdf_integrate(cfg, t, vol, f0.depth(0), ident)
    out["volume_sha"] = np.array(sha(vol))
    out["volume_observed"] = np.array(int((vol[..., 1] > 0).sum()))
    for level in range(cfg.levels):
        vm, nm = O.tsdf_raycast(cfg, t, vol, ident, level, hint=f0.depth(level))
        out[f"model_vmap_sha_l{level}"] = np.array(sha(vm))
        out[f"model_nmap_sha_l{level}"] = np.array(sha(nm))
        out[f"model_valid_l{level}"] = np.array(int((vm[..., 3] > 0).sum()))
    np.savez_compressed(os.path.join(HERE, "golden_ext_160x120.npz"), **out)
    print("wrote golden_ext_160x120.npz; model pose", poses[-1], "sizes", out["yd16_sizes"])


if __name__ == "__main__":
    main()
