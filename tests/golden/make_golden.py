#!/usr/bin/env python
"""Generates tests/golden/golden_160x120.npz from the CPU oracle (run here, committed with
its output).  The reference ships no golden vectors; these pin the
oracle against its own regressions and give the GPU tests a fixture that does not depend
on re-running the oracle.  Inputs come from the C generator (slam-rgbd_b200/host/youth_synth.c)
at 160x120 with scaled Astra intrinsics, sequence 3, noise on."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import oracle_py as O  # noqa: E402
import youth_pkg  # noqa: E402

W, H, N = 160, 120, 4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    pkg = youth_pkg.load()
    frames = pkg.synth_sequence(N, W, H, sequence=3, noise=1)
    cfg = O.default_config(width=W, height=H, fx=570.3 * W / 640, fy=570.3 * W / 640, cx=W / 2.0, cy=H / 2.0)
    poses, status, _ = O.track_sequence(cfg, frames)
    prev, cur = O.OFrame(cfg, frames[0]), O.OFrame(cfg, frames[1])
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    out = {"frames": frames, "poses": poses, "status": status}
    for level in range(cfg.levels):
        sums, corr = O.icp_sums(cfg, level, cur, prev, ident)
        out[f"sums_l{level}"] = sums
        out[f"corr_l{level}"] = corr
        out[f"mask_l{level}"] = cur.mask(level)
        out[f"depth_sha_l{level}"] = np.array(sha(cur.depth(level)))
        out[f"vmap_sha_l{level}"] = np.array(sha(cur.vmap(level)))
        out[f"nmap_sha_l{level}"] = np.array(sha(cur.nmap(level)))
    np.savez_compressed(os.path.join(HERE, "golden_160x120.npz"), **out)
    print("wrote golden_160x120.npz; final pose", poses[-1])


if __name__ == "__main__":
    main()
