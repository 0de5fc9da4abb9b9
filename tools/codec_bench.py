#!/usr/bin/env python
"""YD16 codec micro-benchmark: device time of the encode / decode kernels over a batch of
synthetic 640x480 frames already resident in HBM (CUDA events inside the library), the
compression ratio, and the HBM roofline fraction of each direction (algorithmic bytes = raw
frame + packed stream, each moved once).  One JSON line on stdout."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--noise", type=int, default=0)
    args = ap.parse_args()
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    n, w, h = args.frames, args.width, args.height
    frames = pkg.synth_sequence(n, w, h, noise=args.noise)
    dev = torch.from_numpy(frames.view(np.int16)).cuda()
    cd = pkg.Codec(w, h, max_frames=n)
    out = np.empty(n * cd.max_bytes(), dtype=np.uint8)
    enc_ms, dec_ms = [], []
    packed = offs = None
    back = torch.empty_like(dev)
    for _ in range(args.reps + 2):
        packed, offs = cd.encode_ptr(dev.data_ptr(), n, B.MEM_DEVICE, out)
        enc_ms.append(cd.last_kernel_ms())
        cd.decode_to_device(packed, offs, back.data_ptr())
        dec_ms.append(cd.last_kernel_ms())
    torch.cuda.synchronize()
    ok = bool(torch.equal(back, dev))
    enc, dec = float(np.median(enc_ms[2:])), float(np.median(dec_ms[2:]))
    raw_b, pk_b = frames.nbytes, int(offs[-1])
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    line = {
        "what": f"YD16 codec, {n} synthetic {w}x{h} frames resident in HBM, noise={args.noise}",
        "round_trip_identical": ok, "ratio": raw_b / pk_b, "raw_bytes": raw_b, "packed_bytes": pk_b,
        "encode": {"ms": enc, "frames_per_sec": n / (enc * 1e-3), "gbs": (raw_b + pk_b) / (enc * 1e-3) / 1e9,
                   "frac_of_hbm_peak": (raw_b + pk_b) / (enc * 1e-3) / 1e9 / peak},
        "decode": {"ms": dec, "frames_per_sec": n / (dec * 1e-3), "gbs": (raw_b + pk_b) / (dec * 1e-3) / 1e9,
                   "frac_of_hbm_peak": (raw_b + pk_b) / (dec * 1e-3) / 1e9 / peak},
        "hbm_peak_gbs": peak,
    }
    print(json.dumps(line))
    cd.close()


if __name__ == "__main__":
    main()
