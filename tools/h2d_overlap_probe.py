#!/usr/bin/env python
"""Does a pinned host-to-device copy keep its bandwidth while the tracker's kernels run?  Times a
184 MB H2D copy on its own stream, idle and concurrently with device-resident tracking groups."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    n = 300
    seq = pkg.synth_sequence(n)
    dev = torch.from_numpy(seq.view(np.int16)).cuda()
    trk = B.Tracker(pkg.default_config(batch=n, traj_capacity=n * 16))
    host = torch.empty(seq.size, dtype=torch.int16).pin_memory()
    dst = torch.empty_like(host, device="cuda")
    cs = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def copy_ms(busy):
        trk.reset()
        trk.sync()
        if busy:
            for _ in range(3):
                trk.track_batch_ptrs([dev.data_ptr()], n, B.MEM_DEVICE)
        with torch.cuda.stream(cs):
            e0.record()
            dst.copy_(host, non_blocking=True)
            e1.record()
        torch.cuda.synchronize()
        trk.sync()
        return e0.elapsed_time(e1)

    for busy in (False, True, False, True):
        ms = [copy_ms(busy) for _ in range(4)]
        print(f"tracker busy={busy}: H2D {host.numel() * 2 / 1e6:.0f} MB in {min(ms):.2f} ms "
              f"= {host.numel() * 2 / min(ms) / 1e6:.1f} GB/s (all: {[round(m, 2) for m in ms]})")
    trk.close()


if __name__ == "__main__":
    main()
