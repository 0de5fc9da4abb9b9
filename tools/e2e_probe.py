#!/usr/bin/env python
"""Where does an end-to-end step spend its time?  Host enqueue time vs completion time of one 300-frame
group fed from pinned host memory, raw and YD16-packed (no timing claims beyond this breakdown)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    n = 300
    frames = pkg.synth_sequence(n)
    trk = B.Tracker(pkg.default_config(batch=n, traj_capacity=n))
    pin = trk.lib.youth_cuda_host_alloc(frames.nbytes)
    C.memmove(pin, frames.ctypes.data, frames.nbytes)
    cd = pkg.Codec(640, 480, max_frames=n)
    packed, offs = cd.encode(frames)
    cd.close()
    ppin = trk.lib.youth_cuda_host_alloc(len(packed))
    C.memmove(ppin, packed.ctypes.data, len(packed))
    sp = (C.c_void_p * 1)(ppin)
    op = (C.c_void_p * 1)(offs.ctypes.data)
    out = np.empty((n, 12), dtype=np.float32)

    def raw():
        trk.reset()
        t0 = time.perf_counter()
        trk.track_batch_ptrs([pin], n, B.MEM_HOST_PINNED)
        t1 = time.perf_counter()
        trk.lib.youth_cuda_get_trajectory(trk.h, 0, 0, n, out.ctypes.data, None, None)
        return t1 - t0, time.perf_counter() - t0

    def pk():
        trk.reset()
        t0 = time.perf_counter()
        assert trk.lib.youth_cuda_track_batch_packed(trk.h, sp, op, n, B.MEM_HOST_PINNED, None, None)
        t1 = time.perf_counter()
        trk.lib.youth_cuda_get_trajectory(trk.h, 0, 0, n, out.ctypes.data, None, None)
        return t1 - t0, time.perf_counter() - t0

    for name, fn in (("raw", raw), ("packed", pk)):
        for _ in range(3):
            fn()
        r = np.array([fn() for _ in range(20)])
        print(f"{name}: host enqueue {r[:, 0].mean() * 1e3:.2f} ms, step {r[:, 1].mean() * 1e3:.2f} ms "
              f"(min {r[:, 1].min() * 1e3:.2f})")
    trk.close()


if __name__ == "__main__":
    main()
