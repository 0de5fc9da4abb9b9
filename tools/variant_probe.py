#!/usr/bin/env python
"""Quick A/B of builds of libyouth_cuda.so (slam-rgbd_b200/lib/variants/, `make -C slam-rgbd_b200 variant NAME=.. EXTRA_NVFLAGS=..`) on a GPU box,
without torch: for the default library and each variant, in its own process, track the same 300-frame synthetic
sequence from page-locked host memory (one launch group, icp_ppt 128 = what bench.py runs), print the SHA-1 of the
trajectory (bit-exactness against the default library is the gate) and the per-kernel-class CUDA-event times of
profiled steps (youth_cuda_profile_*; A/B evidence, not a bench value).  One JSON line per library, flushed at once.
    python tools/variant_probe.py [--frames 300] [--reps 5] [names ...]        (default: the default library and every library under lib/variants/)
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHM = "/dev/shm/youth_variant_probe_frames.npy"


def child(name, frames_n, reps, w=640, h=480, levels=3, min_inliers=0):
    t0 = time.time()
    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    frames = np.load(SHM, mmap_mode="r")
    extra = {}
    if (w, h) != (640, 480):
        extra = dict(width=w, height=h, fx=570.3 * w / 640, fy=570.3 * w / 640, cx=w / 2.0, cy=h / 2.0)
    cfg = pkg.default_config(batch=frames_n, icp_ppt=128, traj_capacity=frames_n, levels=levels,
                             iters=[10, 5, 4, 4][:levels] + [0] * (4 - levels), **extra)
    if min_inliers:
        cfg.min_inliers = min_inliers  # timing experiments: no iteration updates the pose (identity association)
    trk = B.Tracker(cfg)
    lib = trk.lib
    nbytes = frames_n * h * w * 2
    pinned = lib.youth_cuda_host_alloc(nbytes)
    assert pinned
    C.memmove(pinned, frames.ctypes.data if hasattr(frames, "ctypes") else np.ascontiguousarray(frames).ctypes.data, nbytes)
    t_init = time.time() - t0

    def step():
        trk.reset()
        trk.track_batch_ptrs([pinned], frames_n, B.MEM_HOST_PINNED)
        trk.sync()

    for _ in range(2):
        step()
    poses, _, st = trk.trajectory()
    sha = hashlib.sha1(np.ascontiguousarray(poses).tobytes()).hexdigest()
    walls = []
    trk.profile(True)
    for _ in range(reps):
        t1 = time.time()
        step()
        walls.append((time.time() - t1) * 1e3)
    ms, n = trk.profile_read()
    trk.profile(False)
    names = {0: "k_ingest", 1: "k_normals", 2: "k_icp_L0", 3: "k_icp_L1", 4: "k_icp_L2", 5: "k_icp_L3", 7: "k_compose"}
    per = {names[k]: round(float(ms[k]) / reps, 4) for k in names if n[k]}
    print(json.dumps({"lib": name, "frames": frames_n, "trajectory_sha1": sha, "lost": int((st & 2 != 0).sum()),
                      "ms_per_step_by_kernel": per, "sum_ms": round(sum(per.values()), 4),
                      "wall_ms_per_profiled_step_from_host_frames": round(min(walls), 3), "init_s": round(t_init, 2)}), flush=True)
    lib.youth_cuda_host_free(pinned)
    trk.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--child", default=None)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--min-inliers", type=int, default=0)
    ap.add_argument("names", nargs="*", default=None)
    args = ap.parse_args()
    if not args.names:
        import glob

        found = sorted(glob.glob(os.path.join(ROOT, "slam-rgbd_b200", "lib", "variants", "libyouth_cuda_*.so")))
        args.names = ["default"] + [os.path.basename(f)[len("libyouth_cuda_"):-3] for f in found]
    if args.child:
        return child(args.child, args.frames, args.reps, args.width, args.height, args.levels, args.min_inliers)
    import youth_pkg

    pkg = youth_pkg.load()
    np.save(SHM, pkg.synth_sequence(args.frames, args.width, args.height))
    vdir = os.path.join(ROOT, "slam-rgbd_b200", "lib", "variants")
    for name in args.names:
        env = dict(os.environ)
        env.pop("YOUTH_CUDA_LIB", None)
        if name != "default":
            so = os.path.join(vdir, f"libyouth_cuda_{name}.so")
            if not os.path.exists(so):
                print(json.dumps({"lib": name, "error": "not built"}), flush=True)
                continue
            env["YOUTH_CUDA_LIB"] = so
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, "--frames", str(args.frames),
                              "--reps", str(args.reps), "--width", str(args.width), "--height", str(args.height),
                              "--levels", str(args.levels), "--min-inliers", str(args.min_inliers)], env=env, capture_output=True, text=True)
        sys.stdout.write(res.stdout)
        if res.returncode != 0:
            print(json.dumps({"lib": name, "error": res.stderr[-600:]}), flush=True)
        sys.stdout.flush()
    try:
        os.remove(SHM)
    except OSError:
        pass


if __name__ == "__main__":
    main()
