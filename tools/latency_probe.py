#!/usr/bin/env python
"""Where does a single-frame (batch 1) call spend its time?  Device-resident vs pinned vs
pageable input, blocking vs enqueue-only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
import youth_pkg
pkg = youth_pkg.load()
from slam_rgbd_b200 import binding as B

frames = pkg.synth_sequence(128)
fb = frames[0].nbytes
dev = torch.from_numpy(frames.view(np.int16)).cuda()
for ppt in (64, 16):
    trk = B.Tracker(pkg.default_config(batch=1, traj_capacity=100000, icp_ppt=ppt))
    pin = trk.lib.youth_cuda_host_alloc(frames.nbytes)
    C.memmove(pin, frames.ctypes.data, frames.nbytes)
    pose = np.empty((1, 1, 12), dtype=np.float32)
    def run(kind, base, blocking, n=100):
        for i in range(8):
            trk.track_batch_ptrs([base + i * fb], 1, kind, None, pose if blocking else None)
        trk.sync()
        t0 = time.perf_counter()
        for i in range(n):
            trk.track_batch_ptrs([base + (8 + i) * fb], 1, kind, None, pose if blocking else None)
        trk.sync()
        return (time.perf_counter() - t0) / n * 1e6
    print(f"ppt={ppt:3d} device enqueue-only {run(B.MEM_DEVICE, dev.data_ptr(), False):7.1f} us/frame | device blocking "
          f"{run(B.MEM_DEVICE, dev.data_ptr(), True):7.1f} | pinned blocking {run(B.MEM_HOST_PINNED, pin, True):7.1f} | "
          f"pinned enqueue-only {run(B.MEM_HOST_PINNED, pin, False):7.1f} | pageable blocking "
          f"{run(B.MEM_HOST, frames.ctypes.data, True):7.1f}")
    trk.close()

# per-kernel-class device time for single-frame groups (CUDA events around every launch)
trk = B.Tracker(pkg.default_config(batch=1, traj_capacity=100000))
for i in range(8):
    trk.track_batch_ptrs([dev.data_ptr() + i * fb], 1, B.MEM_DEVICE)
trk.profile(True)
for i in range(50):
    trk.track_batch_ptrs([dev.data_ptr() + (8 + i) * fb], 1, B.MEM_DEVICE)
ms, n = trk.profile_read()
names = ["k_ingest", "k_normals", "k_icp_L0", "k_icp_L1", "k_icp_L2", "k_icp_L3", "-", "k_compose"]
print("batch-1 per-launch device time (us):", {names[i]: round(ms[i] / n[i] * 1e3, 1) for i in range(8) if n[i]})
trk.close()
