#!/usr/bin/env python
"""Stress of the YD16 encoder next to a live tracker handle (the situation of bench.py's packed-input arm): encode the
300 frames of a sequence `--reps` times at the given size, after tracking them, and count encoder errors."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=960)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    torch.cuda.set_device(args.device)
    w, h, n = args.width, args.height, 300
    cfg = pkg.default_config(batch=n, device=args.device, traj_capacity=n, levels=args.levels, icp_ppt=128,
                             iters=[10, 5, 4, 4][:args.levels] + [0] * (4 - args.levels), width=w, height=h,
                             fx=570.3 * w / 640, fy=570.3 * w / 640, cx=w / 2.0, cy=h / 2.0)
    trk = B.Tracker(cfg)
    frames = pkg.synth_sequence(n, w, h)
    d = torch.from_numpy(frames.view(np.int16)).cuda()
    errors = 0
    for i in range(args.reps):
        trk.reset()
        trk.track_batch_ptrs([d.data_ptr()], n, B.MEM_DEVICE)
        if i % 2 == 0:
            trk.sync()  # every other repetition encodes while the tracker's kernels are still running
        cd = pkg.Codec(w, h, max_frames=n, device=args.device)
        try:
            packed, offs = cd.encode_ptr(d.data_ptr(), n, B.MEM_DEVICE)
        except Exception as e:
            errors += 1
            print(f"rep {i}: {e}", flush=True)
        cd.close()
    trk.sync()
    trk.close()
    print(f"codec_stress: {args.reps} encodes of {n} frames {w}x{h}, {errors} errors")


if __name__ == "__main__":
    main()
