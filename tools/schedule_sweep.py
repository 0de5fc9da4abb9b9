#!/usr/bin/env python
"""Sweep of the stage 3-5 launch schedule (youth_cuda_set_icp_schedule) on device-resident frames:
frames/s of a 300-frame group for (icp_ppt, pairs_per_group, queues), and a bit-equality check of the
trajectory against the ungrouped schedule of the same icp_ppt.  Prints one JSON line per point."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=300)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--ppt", default="64,32,16")
    ap.add_argument("--groups", default="0,4,8,12,16,24")
    ap.add_argument("--queues", default="1,2,3,4")
    args = ap.parse_args()
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    n = args.batch
    seq = pkg.synth_sequence(n)
    dev = torch.from_numpy(seq.view(np.int16)).cuda()
    for ppt in [int(x) for x in args.ppt.split(",")]:
        cfg = pkg.default_config(batch=n, icp_ppt=ppt, traj_capacity=n)
        trk = B.Tracker(cfg)
        ref = None
        for G in [int(x) for x in args.groups.split(",")]:
            for K in [int(x) for x in args.queues.split(",")]:
                if G == 0 and K > 1:
                    continue
                trk.set_icp_schedule(G, K)
                best = 1e30
                for r in range(args.reps + 2):
                    trk.reset()
                    trk.sync()
                    trk.timer_start()
                    trk.track_batch_ptrs([dev.data_ptr()], n, B.MEM_DEVICE)
                    ms = trk.timer_stop()
                    if r >= 2:
                        best = min(best, ms)
                poses, _, st = trk.trajectory()
                if ref is None:
                    ref = poses.copy()
                same = bool(np.array_equal(poses.view(np.uint32), ref.view(np.uint32)))
                print(json.dumps({"ppt": ppt, "group": G, "queues": K, "ms": round(best, 3),
                                  "fps": round(n / best * 1e3, 0), "bit_identical": same,
                                  "lost": int((st & 2 != 0).sum())}), flush=True)
        trk.close()


if __name__ == "__main__":
    main()
