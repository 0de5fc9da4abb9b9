#!/usr/bin/env python
"""Short, ncu-friendly driver: a few launch groups of the tracking path on device-resident
frames (no timing claims are made from runs under a profiler)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--groups", type=int, default=2)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--ppt", type=int, default=64)
    ap.add_argument("--model", action="store_true", help="frame-to-model tracking (TSDF fusion + ray cast)")
    args = ap.parse_args()
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    n = args.batch * args.groups
    cfg = pkg.default_config(batch=args.batch, n_streams=args.streams, width=args.width, height=args.height,
                             levels=args.levels, iters=[10, 5, 4, 4][:args.levels] + [0] * (4 - args.levels),
                             icp_ppt=args.ppt, traj_capacity=max(n, 1),
                             fx=570.3 * args.width / 640, fy=570.3 * args.width / 640, cx=args.width / 2,
                             cy=args.height / 2)
    trk = B.Tracker(cfg)
    if args.model:
        trk.enable_model(pkg.tsdf_config())
    seqs = [pkg.synth_sequence(n, args.width, args.height, sequence=s) for s in range(args.streams)]
    dev = [torch.from_numpy(s.view(np.int16)).cuda() for s in seqs]
    fb = args.width * args.height * 2
    for g in range(args.groups):
        trk.track_batch_ptrs([d.data_ptr() + g * args.batch * fb for d in dev], args.batch, B.MEM_DEVICE)
    trk.sync()
    poses, _, st = trk.trajectory()
    print(f"tracked {n} frames x {args.streams} streams, launches={trk.launch_count()}, lost={(st & 2 != 0).sum()}, "
          f"final t={poses[-1].reshape(3, 4)[:, 3]}")
    trk.close()


if __name__ == "__main__":
    main()
