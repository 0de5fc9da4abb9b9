#!/usr/bin/env python
"""bench.py -- frames/s of the frame-to-frame depth tracking path on B200.

Workload (BASELINE.json configs[1]): synthetic 640x480 uint16 depth sequence, 300 frames,
3-level pyramid, 10/5/4 ICP iterations (fine->coarse listing), bilateral filter on.  One
"step" = one pass of the whole path over one 300-frame sequence per GPU.

  value      frames/s with the raw frames already resident in HBM (device-timed, CUDA events
             on the launching stream), whole job over all ranks
  e2e        the same metric through the C ABI with HOST (pinned) buffers: the H2D copy of
             every frame and the D2H read of the trajectory are inside the timed region; two
             steps in flight (submit step i, then collect step i-1); e2e_blocking = one
             blocking step at a time
  roofline   dominant kernel (k_icp at level 0): algorithmic 48 B/pixel/iteration x pixels
             per launch / average launch duration (CUDA events, measured live in a separate
             profiled step) vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port: the reference has no CPU implementation of this
             path) timed on a bounded sample of the same workload on the host cores

`--impl reference` times the CPU oracle with all host threads on the same config (see
DESIGN.md section 6: the reference repository contains no tracker to run).

Multi-GPU: one process per GPU (torchrun), one independent sequence per rank, no data-path
collective; NCCL all_gather of the per-sequence trajectories once per step.
"""
import argparse
import collections
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, FRAMES = 640, 480, 300
ALG_BYTES_PER_PX_ITER = 48  # stream cur V+N (24 B) + gather prev V+N (24 B)
HBM_FALLBACK_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_RESULT_FD = None


def claim_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.
    Keep the real stdout for the result and point fd 1 at stderr for everything else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.regions = []  # (label, wall-clock begin, end): the timed regions of the run

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_begin(self, label="device"):
        self._open = (label, time.time())

    def mark_end(self):
        self.regions.append((self._open[0], self._open[1], time.time()))

    @staticmethod
    def _stamp(text):
        # nvidia-smi prints its own sampling time ("2026/10/18 14:03:07.123", local time): samples are
        # assigned to regions by that stamp, so pipe buffering cannot move them
        import datetime

        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            ts = self._stamp(parts[0])
            try:
                rows.append((ts, float(parts[1]), float(parts[2]), [v.lower().startswith("active") for v in parts[4:8]]))
            except ValueError:
                continue

        def pick(labels):
            return [r for r in rows if r[0] is not None and
                    any(lb in labels and b - 0.02 <= r[0] <= e + 0.02 for lb, b, e in self.regions)]

        used = ["device"]
        got = pick(used)
        if len(got) < 3:  # a short device-resident region: take every timed region of the run (all under load)
            used = sorted({lb for lb, _, _ in self.regions})
            got = pick(used)
        reasons = sorted({nm for r in got for nm, on in zip(names, r[3]) if on})
        return {"sm_mhz": float(np.median([r[1] for r in got])) if got else None,
                "sm_max_mhz": max(r[2] for r in got) if got else None,
                "samples": len(got), "regions": used, "reasons": reasons}


# --------------------------------------------------------------------------- CPU oracle arm


def cpu_oracle_fps(frames, cfg_product, threads, frames_per_thread, fast=True, tsdf_cfg=None):
    """Track `threads` independent sub-sequences of `frames_per_thread` frames in parallel
    (same partitioning as the GPU's independent frame pairs).  Returns (fps, seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O

    ocfg = O.config_from(cfg_product)
    use_fast = False
    if tsdf_cfg is not None:  # frame-to-model statement: parity build only
        fast = False
        otcfg = O.tsdf_config_from(tsdf_cfg)
    if fast:
        try:
            O.lib(fast=True)
            use_fast = True
        except Exception:
            use_fast = False
    n = frames.shape[0]
    chunks = []
    for t in range(threads):
        a = (t * frames_per_thread) % max(1, n - frames_per_thread)
        chunks.append(np.ascontiguousarray(frames[a:a + frames_per_thread]))
    secs = [0.0] * threads

    def work(i):
        if tsdf_cfg is not None:
            O.track_sequence_model(ocfg, otcfg, chunks[i])
        else:
            _, _, s = O.track_sequence(ocfg, chunks[i], fast=use_fast)
            secs[i] = s

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    tracked = threads * frames_per_thread
    return tracked / wall, wall, ("-O3 -mavx2 -mfma build" if use_fast else "-O2 -ffp-contract=off parity build")


def bind_to_gpu_numa_node(local):
    """Run this rank (and so first-touch its pinned host buffers) on the CPUs of the NUMA node the GPU hangs
    off: with one rank per GPU the H2D streams of 8 GPUs otherwise cross the socket interconnect.
    Best effort: returns a description, never raises."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return f"gpu {bdf}: no NUMA node reported"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return f"gpu {bdf}: node {node} has no CPU in this process's affinity mask"
        os.sched_setaffinity(0, allowed)
        return f"gpu {bdf}: bound to NUMA node {node} ({len(allowed)} CPUs)"
    except Exception as e:  # no sysfs, old torch, restricted container ...
        return f"not bound ({type(e).__name__}: {e})"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- main arms


def base_config_dict(args, n_gpus):
    spg = getattr(args, "sequences_per_gpu", 1)
    w, h, lv = getattr(args, "width", W), getattr(args, "height", H), getattr(args, "levels", 3)
    if getattr(args, "mode", "frame") == "model":
        name = "frame-to-model variant of configs[1]/[3] (TSDF 256x128x256 @ 25 mm, one captured graph per frame)"
    elif (w, h, lv, spg) == (W, H, 3, 1):
        name = "configs[1]"
    elif (w, h, lv) == (W, H, 3):
        name = "configs[3]-style (several sequences per GPU)"
    else:
        name = "configs[4]-style (high resolution)"
    iters = "/".join(str(x) for x in [10, 5, 4, 4][:lv])
    raw_mb = spg * FRAMES * w * h * 2 / 1e6
    return {
        "workload": f"{name}: synthetic {w}x{h} uint16 depth sequence, 300 frames, {lv}-level pyramid, "
                    f"ICP iterations fine->coarse = {iters}, 7x7 bilateral on, {spg} sequence(s) per GPU",
        "frames_per_step_per_gpu": FRAMES * spg,
        "batch_frames_per_launch_group": args.batch,
        "icp_ppt": getattr(args, "ppt", 0) or 64,
        "sequences_per_gpu": spg,
        "partition": f"{n_gpus * spg} independent sequence(s), {spg} per GPU, no data-path collective",
        "host_placement": getattr(args, "numa_note", "n/a"),
        "l2": f"inputs ({raw_mb:.0f} MB raw depth per step) exceed the 126 MB L2 and are streamed once per step; "
              "no explicit flush",
    }


def run_reference(args):
    """CPU arm: the oracle port on all host threads (the reference has no tracker to run)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import youth_pkg

    pkg = youth_pkg.load()
    cfg = pkg.default_config(**({"icp_ppt": args.ppt} if args.ppt else {}))
    threads = max(1, min(host_threads(), 64))
    fpt = 11  # 10 frame pairs per thread per step: a bounded sample of the 300-frame workload
    frames = pkg.synth_sequence(FRAMES)
    fps_all = []
    kind = ""
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        fps, wall, kind = cpu_oracle_fps(frames, cfg, threads, fpt)
        if i >= args.warmup:
            fps_all.append((fps, wall))
        if time.perf_counter() - t_start > 240:
            break
    if not fps_all:
        fps_all.append((fps, wall))
    val = float(np.mean([f for f, _ in fps_all]))
    ms = float(np.mean([w for _, w in fps_all]) * 1e3)
    sample = f"{threads} threads x {fpt} consecutive frames each per step ({kind}); oracle port, no reference tracker exists"
    line = {
        "impl": "reference", "metric": "icp_tracked_frames_per_sec", "value": val, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(fps_all), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config_dict(args, args.gpus),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def run_ours(args):
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tracking path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else "single rank: not bound"
    log(f"[rank {rank}] {numa}")
    args.numa_note = numa
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    stream = torch.cuda.Stream()
    if args.mode == "model" and args.batch == 300:
        args.batch = 20  # frames per call; the chain runs frame by frame anyway and the ring only needs a few slots
    if args.mode == "model" and not args.ppt:
        # a chain of single frames is latency-bound: more, shorter ICP runs per frame (icp_ppt is part of the
        # configuration the CPU statement follows); measured best: 16 for one sequence, 32 for eight
        args.ppt = 16 if args.sequences_per_gpu <= 2 else 32
    extra = {"icp_ppt": args.ppt} if args.ppt else {}
    Wd, Hd, S = args.width, args.height, args.sequences_per_gpu
    if (Wd, Hd) != (W, H):
        extra.update(width=Wd, height=Hd, fx=570.3 * Wd / 640, fy=570.3 * Wd / 640, cx=Wd / 2.0, cy=Hd / 2.0)
    cfg = pkg.default_config(batch=args.batch, n_streams=S, device=local, traj_capacity=FRAMES, levels=args.levels,
                             stream=stream.cuda_stream, **extra)
    trk = B.Tracker(cfg)
    if args.mode == "model":  # frame-to-model tracking (include/youth_model.h): a sequence is a chain of frames
        trk.enable_model(pkg.tsdf_config())

    # independent sequences: rank r tracks sequences r*S .. r*S+S-1 (seed 20261018 + sequence)
    t0 = time.perf_counter()
    frames = np.stack([pkg.synth_sequence(FRAMES, Wd, Hd, sequence=rank * S + k) for k in range(S)])
    log(f"[rank {rank}] generated {S} x {FRAMES} frames in {time.perf_counter() - t0:.1f}s")
    frame_bytes = Wd * Hd * 2
    seq_bytes = FRAMES * frame_bytes
    d_frames = torch.from_numpy(frames.view(np.int16)).to(f"cuda:{local}")
    d_base = d_frames.data_ptr()
    # pinned host copy for the e2e arm
    pin_ptr = trk.lib.youth_cuda_host_alloc(frames.nbytes)
    if not pin_ptr:
        raise SystemExit("pinned allocation failed")
    C.memmove(pin_ptr, frames.ctypes.data, frames.nbytes)

    groups = [(a, min(args.batch, FRAMES - a)) for a in range(0, FRAMES, args.batch)]
    traj_view = None
    if dist is not None:
        class _Ptr:
            pass
        holder = _Ptr()
        holder.__cuda_array_interface__ = {
            "shape": (S, FRAMES, 12), "typestr": "<f4", "version": 3,  # traj_capacity == FRAMES: streams are contiguous
            "data": (trk.lib.youth_cuda_trajectory_device_ptr(trk.h, 0), False)}
        traj_view = torch.as_tensor(holder, device=f"cuda:{local}")
        gathered = torch.empty((world, S, FRAMES, 12), dtype=torch.float32, device=f"cuda:{local}")

    def step_device():
        trk.reset()
        for a, n in groups:
            trk.track_batch_ptrs([d_base + k * seq_bytes + a * frame_bytes for k in range(S)], n, B.MEM_DEVICE)
        if dist is not None:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, traj_view)

    host_traj = np.empty((S, FRAMES, 12), dtype=np.float32)

    def step_e2e():
        trk.reset()
        for a, n in groups:
            trk.track_batch_ptrs([pin_ptr + k * seq_bytes + a * frame_bytes for k in range(S)], n, B.MEM_HOST_PINNED)
        for k in range(S):
            got = trk.lib.youth_cuda_get_trajectory(trk.h, k, 0, FRAMES, host_traj[k].ctypes.data, None, None)
            assert got == FRAMES

    def barrier():
        trk.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident arm
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~1 s to start printing; only in-region samples are kept
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = trk.launch_count()
    sampler.mark_begin()
    trk.timer_start()
    for _ in range(args.steps):
        step_device()
    ms_total = trk.timer_stop()
    barrier()
    sampler.mark_end()
    launches = trk.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * S * FRAMES / (ms_per_step * 1e-3)

    # pose error vs synthetic ground truth (reported, not part of the timed region)
    poses, _, status = trk.trajectory()
    gt = pkg.synth_gt(FRAMES, Wd, Hd, sequence=rank * S)
    terr = np.linalg.norm(poses.reshape(-1, 3, 4)[:, :, 3] - gt.reshape(-1, 3, 4)[:, :, 3], axis=1)
    lost = int((status & B.STATUS_LOST != 0).sum())

    # ---- end-to-end arm (host buffers through the C ABI).  Every step copies its 300 frames per sequence from
    # pinned host memory and reads its trajectory back into pinned host memory, all inside the timed region.
    # Headline: two steps in flight -- submit step i (youth_cuda_track_batch + youth_cuda_read_trajectory_async),
    # then collect step i-1 (youth_cuda_wait_ticket) -- so that the H2D copy of a step runs on the copy stream
    # under the kernels of the step before.  Also reported: one blocking step at a time (e2e_blocking).
    res_pin = [trk.lib.youth_cuda_host_alloc(S * FRAMES * 48) for _ in range(2)]
    res_np = [np.ctypeslib.as_array((C.c_float * (S * FRAMES * 12)).from_address(p)).reshape(S, FRAMES, 12) for p in res_pin]
    pending = collections.deque()

    def submit(i):
        trk.reset()
        for a, n in groups:
            trk.track_batch_ptrs([pin_ptr + k * seq_bytes + a * frame_bytes for k in range(S)], n, B.MEM_HOST_PINNED)
        ticket = None
        for k in range(S):
            got, ticket = trk.read_trajectory_async(res_pin[i % 2] + k * FRAMES * 48, FRAMES, stream=k)
            assert got == FRAMES
        pending.append((i, ticket))

    def collect():
        i, ticket = pending.popleft()
        trk.wait_ticket(ticket)
        host_traj[...] = res_np[i % 2]  # the step's result, consumed on the host

    def run_pipelined(steps):
        for i in range(steps):
            submit(i)
            if len(pending) > 1:
                collect()
        while pending:
            collect()

    run_pipelined(args.warmup)
    barrier()
    sampler.mark_begin("e2e")
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    trk.sync()
    e2e_s = time.perf_counter() - t0
    sampler.mark_end()
    if dist is not None:
        torch.cuda.synchronize()
        t = torch.tensor([e2e_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * S * FRAMES * args.steps / e2e_s
    poses_e2e = host_traj[0].copy()
    e2e_matches = bool(np.array_equal(poses_e2e.view(np.uint32), poses.view(np.uint32)))

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    sampler.mark_begin("e2e_blocking")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    trk.sync()
    blk_s = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        torch.cuda.synchronize()
        t = torch.tensor([blk_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        blk_s = float(t.item())
    e2e_blocking = world * S * FRAMES * args.steps / blk_s
    e2e_matches = e2e_matches and bool(np.array_equal(host_traj[0].view(np.uint32), poses.view(np.uint32)))
    for p in res_pin:
        trk.lib.youth_cuda_host_free(p)

    # ---- packed-input arm: the same step fed from YD16 streams (include/youth_codec.h) in pinned host
    # memory; the packed bytes cross PCIe and are unpacked on the device.  Reported next to e2e.
    packed_info = None
    if not args.no_packed and args.mode != "model":
        cd = pkg.Codec(Wd, Hd, max_frames=FRAMES, device=local)
        enc = [cd.encode_ptr(d_base + k * seq_bytes, FRAMES, B.MEM_DEVICE) for k in range(S)]
        enc_ms = cd.last_kernel_ms()
        back = torch.empty(FRAMES * Wd * Hd, dtype=torch.int16, device=f"cuda:{local}")
        cd.decode_to_device(enc[0][0], enc[0][1], back.data_ptr())
        dec_ms = cd.last_kernel_ms()
        codec_ok = bool(torch.equal(back.view(FRAMES, Hd, Wd), d_frames[0]))
        del back
        cd.close()
        pk_bytes = [int(e[1][-1]) for e in enc]
        pk_pin = [trk.lib.youth_cuda_host_alloc(b) for b in pk_bytes]
        for ptr, e in zip(pk_pin, enc):
            C.memmove(ptr, e[0].ctypes.data, len(e[0]))
        pk_ptrs = (C.c_void_p * S)(*pk_pin)

        def step_packed():
            trk.reset()
            for a, n in groups:
                offs = [np.ascontiguousarray(e[1][a:a + n + 1]) for e in enc]
                op = (C.c_void_p * S)(*[o.ctypes.data for o in offs])
                ok = trk.lib.youth_cuda_track_batch_packed(trk.h, pk_ptrs, op, n, B.MEM_HOST_PINNED, None, None)
                assert ok, trk.lib.youth_cuda_last_error()
            for k in range(S):
                got = trk.lib.youth_cuda_get_trajectory(trk.h, k, 0, FRAMES, host_traj[k].ctypes.data, None, None)
                assert got == FRAMES

        for _ in range(args.warmup):
            step_packed()
        barrier()
        psteps = max(1, min(args.steps, 20))
        t0 = time.perf_counter()
        for _ in range(psteps):
            step_packed()
        trk.sync()
        pk_s = time.perf_counter() - t0
        if dist is not None:
            torch.cuda.synchronize()
            t = torch.tensor([pk_s], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            pk_s = float(t.item())
        raw_b = FRAMES * frame_bytes
        packed_info = {
            "e2e_value": world * S * FRAMES * psteps / pk_s, "unit": "frames/s", "steps": psteps,
            "h2d_bytes_per_step": int(sum(pk_bytes)), "compression_ratio": S * raw_b / float(sum(pk_bytes)),
            "bit_identical_to_device_arm": bool(np.array_equal(host_traj[0].view(np.uint32), poses.view(np.uint32))),
            "codec_round_trip_identical": codec_ok,
            "k_yd16_encode": {"ms_per_300_frames": enc_ms, "gbs": (raw_b + pk_bytes[-1]) / (enc_ms * 1e-3) / 1e9},
            "k_yd16_decode": {"ms_per_300_frames": dec_ms, "gbs": (raw_b + pk_bytes[0]) / (dec_ms * 1e-3) / 1e9},
            "what": "youth_cuda_track_batch_packed: YD16 streams in pinned host memory, unpacked on the device",
        }
        for ptr in pk_pin:
            trk.lib.youth_cuda_host_free(ptr)

    # ---- per-kernel timing (separate profiled step: events around every launch)
    trk.profile(True)
    step_device()
    trk.sync()
    prof_ms, prof_n = trk.profile_read()
    trk.profile(False)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    peak, peak_kind = measured_peak()
    icp0_ms = prof_ms[B.PROF_ICP0] / max(1, int(prof_n[B.PROF_ICP0]))
    # launches differ in pairs per launch (last group is shorter): use the average pairs per launch
    icp0_launches = int(prof_n[B.PROF_ICP0])
    pairs_total = S * FRAMES * cfg.iters[0]
    bytes_per_launch = ALG_BYTES_PER_PX_ITER * Wd * Hd * pairs_total / max(1, icp0_launches)
    achieved = bytes_per_launch / (icp0_ms * 1e-3) / 1e9 if icp0_ms > 0 else 0.0
    names = {B.PROF_INGEST: "k_ingest", B.PROF_NORMALS: "k_normals", B.PROF_ICP0: "k_icp_L0",
             B.PROF_ICP0 + 1: "k_icp_L1", B.PROF_ICP0 + 2: "k_icp_L2", B.PROF_ICP0 + 3: "k_icp_L3",
             B.PROF_SOLVE: "k_tsdf_integrate", B.PROF_MISC: "k_compose", B.PROF_RAYCAST: "k_tsdf_raycast"}
    step_prof = {names[i]: {"ms": round(float(prof_ms[i]), 4), "launches": int(prof_n[i])}
                 for i in range(B.PROF_CLASSES) if prof_n[i]}

    # second kernel by time: k_ingest.  Algorithmic bytes per frame (SURVEY 8(d) B_pre): raw depth in +
    # vertex/normal maps of all levels out (the level >= 1 normals are written by k_normals).
    npix_all = sum((Wd >> l) * (Hd >> l) for l in range(args.levels))
    b_pre = 2 * Wd * Hd + 24 * npix_all
    ing_ms = (prof_ms[B.PROF_INGEST] + prof_ms[B.PROF_NORMALS]) / max(1, int(prof_n[B.PROF_INGEST]))
    ing_frames = S * FRAMES / max(1, int(prof_n[B.PROF_INGEST]))
    ing_gbs = b_pre * ing_frames / (ing_ms * 1e-3) / 1e9 if ing_ms > 0 else 0.0
    ingest_roof = {"bound": "hbm", "kernel": "k_ingest + k_normals", "achieved": ing_gbs, "peak": peak, "unit": "GB/s",
                   "frac": ing_gbs / peak, "algorithmic_bytes_per_frame": b_pre,
                   "note": "bounded by shared-memory wavefronts of the 49-tap bilateral (ncu: 87 % of the LSU data-pipe peak, "
                           "78 % issue utilisation), not by HBM"}

    traffic = None  # ncu DRAM bytes per launch: only valid for the configuration it was captured on
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if ((Wd, Hd, S, args.levels) == (W, H, 1, 3) and tj.get("pairs_per_launch") == args.batch and args.mode == "frame"
                and tj.get("icp_ppt", 64) == (args.ppt or 64)):
            traffic = tj.get("k_icp_L0_dram_bytes_per_launch")
    except Exception:
        pass

    # ---- live mode: one frame per call (youth_cuda_track), pose read back every frame
    streaming = None
    if (Wd, Hd, S) == (W, H, 1) and args.mode != "model":
        live = B.Tracker(pkg.default_config(batch=1, device=local, traj_capacity=128, levels=args.levels,
                                            **({"icp_ppt": args.user_ppt} if args.user_ppt else {})))
        for i in range(8):
            live.track(frames[0][i], ts=33 * i)
        t0 = time.perf_counter()
        nlive = 100
        for i in range(nlive):
            live.track(frames[0][8 + i], ts=33 * (8 + i))
        dt = time.perf_counter() - t0
        live.close()
        streaming = {"frames_per_sec": nlive / dt, "us_per_frame": dt / nlive * 1e6,
                     "what": "youth_cuda_track: one pageable host frame in, blocking pose out, batch 1"}

    # ---- the reference-facing facade (include/SLAM.h): one processSlamFrame() call per frame from a producer
    # thread (synchronous copy into the pinned host ring, SLAM.cpp:133-134), worker thread with two runs of
    # 64 frames in flight, youthSlamDrain + trajectory read-back at the end of every pass
    facade = None
    if (Wd, Hd, S) == (W, H, 1) and args.mode != "model" and args.levels == 3:
        try:
            os.environ["YOUTH_SLAM_DEVICE"] = str(local)
            os.environ["YOUTH_SLAM_TRAJ_CAPACITY"] = str(FRAMES)
            if args.ppt:
                os.environ["YOUTH_SLAM_ICP_PPT"] = str(args.ppt)  # same reduction geometry as the device arm
            host = pkg.host_lib()
            host.youthSlamSetOptions(1, 64)  # lossless, 64 frames per launch group (the facade's maximum)
            host.initSlamModule(None, None)
            if host.isSlamModuleRunning() == 1:
                fposes = np.empty((FRAMES, 12), dtype=np.float32)
                ptrs = [frames[0][i].ctypes.data for i in range(FRAMES)]

                def facade_pass():
                    host.resetSlam()
                    for i in range(FRAMES):
                        assert host.processSlamFrame(ptrs[i], None, W, H, 33 * i) == 1
                    host.youthSlamDrain()
                    assert host.youthSlamGetTrajectory(fposes.ctypes.data, None, None, FRAMES) == FRAMES

                for _ in range(2):
                    facade_pass()
                fsteps = max(1, min(args.steps, 10))
                t0 = time.perf_counter()
                for _ in range(fsteps):
                    facade_pass()
                fdt = time.perf_counter() - t0
                facade = {"frames_per_sec": FRAMES * fsteps / fdt, "steps": fsteps,
                          "bit_identical_to_device_arm": bool(np.array_equal(fposes.view(np.uint32), poses.view(np.uint32))),
                          "what": "SLAM.h facade: processSlamFrame() per frame (pageable host frame copied into the pinned "
                                  "ring by the caller's thread), lossless, 64 frames per launch group, two groups in flight"}
                host.stopSlamModule()
        except Exception as e:  # the facade arm is informative only
            facade = {"error": str(e)}

    threads = max(1, min(host_threads(), 32))
    fpt = max(3, int(61 * (W * H) / (Wd * Hd)))  # 60 frame pairs per thread at 640x480: about 10-15 s of CPU work
    if args.mode == "model":  # the frame-to-model statement (fusion + ray cast on top of the same ICP), parity build
        fpt = 12
        mt = pkg.tsdf_config()
        cpu_fps, cpu_wall, cpu_kind = cpu_oracle_fps(frames[0], cfg, threads, fpt, tsdf_cfg=mt)
        cpu_kind = "frame-to-model statement, " + cpu_kind
        cpu1_fps, _, _ = cpu_oracle_fps(frames[0], cfg, 1, 4, tsdf_cfg=mt)
        cpu1p_fps = cpu1_fps
    else:
        cpu_fps, cpu_wall, cpu_kind = cpu_oracle_fps(frames[0], cfg, threads, fpt)
        cpu1_fps, _, _ = cpu_oracle_fps(frames[0], cfg, 1, max(3, fpt // 6))               # one thread, speed build
        cpu1p_fps, _, _ = cpu_oracle_fps(frames[0], cfg, 1, max(3, fpt // 6), fast=False)  # one thread, parity build

    line = {
        "metric": "icp_tracked_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config_dict(args, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": S * FRAMES * frame_bytes,
                "d2h_bytes_per_step": S * FRAMES * 48, "bit_identical_to_device_arm": e2e_matches,
                "steps_in_flight": 2,
                "how": "youth_cuda_track_batch (pinned host frames) + youth_cuda_read_trajectory_async per step, "
                       "youth_cuda_wait_ticket of the step before; H2D and D2H of every step inside the timed region"},
        "e2e_blocking": {"value": e2e_blocking, "unit": "frames/s",
                         "how": "one step at a time: youth_cuda_track_batch then a blocking youth_cuda_get_trajectory"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_icp (level 0)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                     "avg_launch_ms": icp0_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "note": "algorithmic 48 B/px/iter = the bytes k_icp requests (three float2 planes per frame, "
                             "24 B/px); each frame's second use in a launch hits L2, so DRAM traffic is about half "
                             "(traffic = ncu dram bytes per launch, profiles/traffic.json); a fraction above 1 is L2 reuse, "
                             "see DESIGN.md section 5; the kernel is latency-bound, not bandwidth-bound: its rate is resident "
                             "lanes x 2 pixels in flight / loaded memory latency (DESIGN.md section 12, item 0a)"},
        "roofline_k_ingest": ingest_roof,
        "per_kernel_ms_per_step": step_prof,
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{threads} threads x {fpt} consecutive frames ({cpu_kind}), {cpu_wall:.1f}s wall",
                         "single_thread_value": cpu1_fps, "single_thread_parity_build_value": cpu1p_fps},
        "pose_error_vs_ground_truth": {"max_translation_m": float(terr.max()), "final_translation_m": float(terr[-1]),
                                       "frames_flagged_lost": lost},
        "frames_per_sec_per_gpu": value / world,
        "streaming_single_frame": streaming,
        "facade": facade,
        "packed_input": packed_info,
    }
    emit(line)
    C.cast(pin_ptr, C.c_void_p)
    trk.lib.youth_cuda_host_free(pin_ptr)
    trk.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=300, help="frames per launch group (300 = the whole sequence)")
    ap.add_argument("--ppt", type=int, default=0, help="override icp_ppt (reduction geometry; 0 = library default)")
    ap.add_argument("--sequences-per-gpu", type=int, default=1, help="independent sequences tracked per GPU (configs[3] uses 8)")
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--no-packed", action="store_true", help="skip the YD16 packed-input arm")
    ap.add_argument("--no-numa", action="store_true", help="do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--mode", default="frame", choices=["frame", "model"],
                    help="frame = frame-to-frame (the headline workload), model = frame-to-model (TSDF fusion + ray cast)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.user_ppt = args.ppt
    if not args.ppt and args.mode == "frame" and args.batch * args.sequences_per_gpu >= 64:
        # launches of many pairs: longer ICP runs (75 runs = 19 CTAs per pair at every level instead of 150 / 38)
        # amortise the per-CTA prologue and tail; measured on B200: 5.89 vs 6.19 ms of k_icp per 300-frame step,
        # 256 is slower again (4.05 waves at level 0).  icp_ppt is part of the configuration the CPU statement
        # follows, so both arms use it.
        args.ppt = 128
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
