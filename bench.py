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
  roofline   dominant kernel (k_icp at level 0) against the measured HBM peak:
             frac = COMPULSORY bytes per launch (every frame a launch touches, moved once: 24 B/pixel)
             / average launch duration (CUDA events, measured live in a separate profiled step) / peak;
             the 48 B/pixel/iteration the kernel requests is printed beside it as
             requested_bytes_per_launch.  roofline_all lists every kernel class of the step the same way.
  cpu_baseline  the CPU oracle (a port: the reference has no CPU implementation of this
             path) timed on a bounded sample of the same workload on the host cores
  parity_vs_oracle  the device trajectory against the oracle's parity build over the first frames of the
             sequence (bit equality of the float poses)
  extra_configs  short arms for BASELINE.json configs[3] (8 sequences per GPU) and configs[4] (1280x960, 4 levels)

`--impl reference` times the CPU oracle with all host threads on the same config (see
DESIGN.md section 6: the reference repository contains no tracker to run); it loads none of the GPU libraries.

Multi-GPU: one process per GPU (torchrun), independent sequences per rank, no data-path
collective; NCCL all_gather of the per-sequence trajectories once per step.
"""
import argparse
import collections
import ctypes as C
import glob
import hashlib
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
REQ_BYTES_PER_PX_ITER = 48  # stream cur V+N (24 B) + gather prev V+N (24 B) = what k_icp requests
MAP_BYTES_PER_PX = 24       # vertex + normal of one frame pixel: what DRAM must deliver once per frame and launch
HBM_FALLBACK_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
CPU_SAMPLE_FRAMES_PER_THREAD = 31  # 30 frame pairs per host thread: the ONE sampler of cpu_baseline and --impl reference
ITERS = [10, 5, 4, 4]


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
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def kernels_sha1():
    """Identity of the device code: SHA-1 over the CUDA sources that hold kernels (everything under csrc/ except
    youth_cuda.cu, which is host code: the C ABI).  profiles/traffic.json carries the value its ncu capture was taken
    with; a capture of other kernels is not reported as this run's traffic."""
    h = hashlib.sha1()
    for p in sorted(glob.glob(os.path.join(ROOT, "slam-rgbd_b200", "csrc", "*.cu*"))):
        if os.path.basename(p) == "youth_cuda.cu":
            continue
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def load_traffic(geom_key):
    """ncu DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one --set full capture) per kernel
    class, or ({}, why-not).  Only for the geometry and the kernel sources the capture was taken with."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
    except Exception as e:
        return {}, f"no profiles/traffic.json ({type(e).__name__})"
    if tj.get("kernels_sha1") != kernels_sha1():
        return {}, "profiles/traffic.json was captured with other kernel sources (kernels_sha1 differs): not reported"
    entry = tj.get("captures", {}).get(geom_key)
    if not entry:
        return {}, f"profiles/traffic.json holds no capture for {geom_key}"
    return entry.get("dram_bytes_per_launch", {}), f"{entry.get('source', 'profiles/')} captured at {tj.get('captured_at_commit', '?')}"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed regions."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.regions = []  # (label, wall-clock begin, end): the timed regions of the run
        self.rows = None

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
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        self.rows = []
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            ts = self._stamp(parts[0])
            try:
                self.rows.append((ts, float(parts[1]), float(parts[2]), [v.lower().startswith("active") for v in parts[4:8]]))
            except ValueError:
                continue

    def summary(self, prefix=""):
        """Clocks over the regions whose label starts with `prefix` + 'device' (or, when that region is too short
        to hold three samples, over all regions with the prefix: all of them are under load)."""
        if self.rows is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def pick(labels):
            return [r for r in self.rows if r[0] is not None and
                    any(lb in labels and b - 0.02 <= r[0] <= e + 0.02 for lb, b, e in self.regions)]

        used = [prefix + "device"]
        got = pick(used)
        if len(got) < 3:
            used = sorted({lb for lb, _, _ in self.regions if lb.startswith(prefix)})
            got = pick(used)
        reasons = sorted({nm for r in got for nm, on in zip(names, r[3]) if on})
        return {"sm_mhz": float(np.median([r[1] for r in got])) if got else None,
                "sm_max_mhz": max(r[2] for r in got) if got else None,
                "samples": len(got), "regions": used, "reasons": reasons}


# --------------------------------------------------------------------------- CPU oracle arm


def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O

    return O


def cpu_sample(frames, ocfg, threads, frames_per_thread, fast=True, tsdf_cfg=None):
    """THE CPU sampler (cpu_baseline and --impl reference both call it with CPU_SAMPLE_FRAMES_PER_THREAD): track
    `threads` independent sub-sequences of `frames_per_thread` consecutive frames in parallel, one per host thread
    (the same partitioning as the GPU's independent frame pairs).  Returns (frames/s, seconds, build)."""
    O = oracle_module()
    use_fast = False
    if tsdf_cfg is not None:  # frame-to-model statement: parity build only
        fast = False
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

    def work(i):
        if tsdf_cfg is not None:
            O.track_sequence_model(ocfg, tsdf_cfg, chunks[i])
        else:
            O.track_sequence(ocfg, chunks[i], fast=use_fast)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    wall = time.perf_counter() - t0
    tracked = threads * frames_per_thread
    return tracked / wall, wall, ("-O3 -mavx2 -mfma build" if use_fast else "-O2 -ffp-contract=off parity build")


def cpu_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def cpu_sample_text(threads, fpt, kind):
    return (f"{threads} host threads x {fpt} consecutive frames each ({kind}); oracle port -- the reference has no "
            f"tracker to run")


def bind_to_gpu_cpus(local):
    """Run this rank (and so first-touch its pinned host buffers, which are allocated AFTER this call) on the CPUs
    next to its GPU: with one rank per GPU the H2D streams of 8 GPUs otherwise cross the socket interconnect.
    Sources, in order: the PCI device's numa_node, the PCI device's local_cpulist, the CPU-affinity column of
    `nvidia-smi topo -m`.  Best effort: returns a description of what was done, never raises."""
    def cpus_of(text):
        out = set()
        for part in text.strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            out.update(range(int(a), int(b or a) + 1))
        return out

    def apply(cpus, how):
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        if allowed == os.sched_getaffinity(0):
            return f"{how}: covers the whole affinity mask ({len(allowed)} CPUs), nothing to narrow"
        os.sched_setaffinity(0, allowed)
        return f"{how}: bound to {len(allowed)} CPUs"

    notes = []
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception as e:
        return f"not bound ({type(e).__name__}: {e})"
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                r = apply(cpus_of(f.read()), f"gpu {bdf} numa_node {node}")
            if r:
                return r
        notes.append("numa_node -1")
    except Exception as e:
        notes.append(f"numa_node: {type(e).__name__}")
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            r = apply(cpus_of(f.read()), f"gpu {bdf} local_cpulist")
        if r:
            return r
        notes.append("local_cpulist empty")
    except Exception as e:
        notes.append(f"local_cpulist: {type(e).__name__}")
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        import re

        clean = re.sub(r"\x1b\[[0-9;]*m", "", out)
        header = None
        for ln in clean.splitlines():
            cols = [c.strip() for c in ln.split("\t") if c.strip()]
            if cols and cols[0].startswith("GPU0") and "CPU Affinity" in ln:
                header = cols
            if header and cols and cols[0] == f"GPU{local}":
                idx = header.index("CPU Affinity") + 1  # the data rows start with their own label
                if idx < len(cols) and re.fullmatch(r"[0-9,\-]+", cols[idx]):
                    r = apply(cpus_of(cols[idx]), f"gpu {local} nvidia-smi topo CPU affinity {cols[idx]}")
                    if r:
                        return r
        notes.append("nvidia-smi topo: no CPU affinity column for this GPU")
    except Exception as e:
        notes.append(f"nvidia-smi topo: {type(e).__name__}")
    return f"gpu {bdf}: not bound ({'; '.join(notes)})"


# --------------------------------------------------------------------------- configuration text


def config_name(w, h, lv, spg, mode):
    if mode == "model":
        return "frame-to-model variant of configs[1]/[3] (TSDF 256x128x256 @ 25 mm, one captured graph per frame)"
    if (w, h, lv, spg) == (W, H, 3, 1):
        return "configs[1]"
    if (w, h, lv) == (W, H, 3):
        return "configs[3] shard (several sequences per GPU)"
    return "configs[4] (high resolution)"


def config_dict(w, h, lv, spg, batch, ppt, n_gpus, mode="frame"):
    """The workload description: the same dict for both arms (nothing arm-specific in it)."""
    iters = "/".join(str(x) for x in ITERS[:lv])
    raw_mb = spg * FRAMES * w * h * 2 / 1e6
    return {
        "workload": f"{config_name(w, h, lv, spg, mode)}: synthetic {w}x{h} uint16 depth sequence, 300 frames, {lv}-level "
                    f"pyramid, ICP iterations fine->coarse = {iters}, 7x7 bilateral on, {spg} sequence(s) per GPU",
        "frames_per_step_per_gpu": FRAMES * spg,
        "batch_frames_per_launch_group": batch,
        "icp_ppt": ppt or 64,
        "sequences_per_gpu": spg,
        "partition": f"{n_gpus * spg} independent sequence(s), {spg} per GPU, no data-path collective",
        "l2": f"inputs ({raw_mb:.0f} MB raw depth per step) exceed the 126 MB L2 and are streamed once per step; "
              "no explicit flush",
    }


def default_ppt(args, batch, spg):
    if args.user_ppt:
        return args.user_ppt
    if args.mode == "model":
        # a chain of single frames is latency-bound: more, shorter ICP runs per frame (icp_ppt is part of the
        # configuration the CPU statement follows); measured best: 16 for one sequence, 32 for eight
        return 16 if spg <= 2 else 32
    # launches of many pairs: longer ICP runs (75 runs = 19 CTAs per pair at every level instead of 150 / 38)
    # amortise the per-CTA prologue and tail; icp_ppt is part of the configuration the CPU statement follows, so
    # both arms use it
    return 128 if batch * spg >= 64 else 0


# --------------------------------------------------------------------------- reference arm


def run_reference(args):
    """CPU arm: the oracle port on all host threads (the reference has no tracker to run).  Loads the synthetic
    generator (libyouth_synth.so, plain C) and the oracle -- none of the GPU libraries."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import youth_pkg

    pkg = youth_pkg.load()
    O = oracle_module()
    ppt = default_ppt(args, args.batch, args.sequences_per_gpu)
    over = {"icp_ppt": ppt} if ppt else {}
    Wd, Hd = args.width, args.height
    if (Wd, Hd) != (W, H):
        over.update(width=Wd, height=Hd, fx=570.3 * Wd / 640, fy=570.3 * Wd / 640, cx=Wd / 2.0, cy=Hd / 2.0)
    ocfg = O.default_config(levels=args.levels, iters=ITERS[:args.levels] + [0] * (4 - args.levels), **over)
    threads = cpu_threads()
    fpt = CPU_SAMPLE_FRAMES_PER_THREAD
    frames = pkg.synth_sequence(FRAMES, Wd, Hd)
    samples = []
    kind = ""
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        fps, wall, kind = cpu_sample(frames, ocfg, threads, fpt)
        if i >= args.warmup:
            samples.append((fps, wall))
        if time.perf_counter() - t_start > 240:
            break
    if not samples:
        samples.append((fps, wall))
    val = float(np.mean([f for f, _ in samples]))
    ms = float(np.mean([w for _, w in samples]) * 1e3)
    line = {
        "impl": "reference", "metric": "icp_tracked_frames_per_sec", "value": val, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(samples), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(Wd, Hd, args.levels, args.sequences_per_gpu, args.batch, ppt, args.gpus, args.mode),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(threads, fpt, kind)},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------- our arm


class Ctx:
    pass


def measure(ctx, Wd, Hd, levels, S, batch, ppt, steps, warmup, label="", full=False):
    """One configuration on this rank's GPU: device-resident arm (CUDA events), end-to-end arm from pinned host
    frames (two steps in flight), a profiled step (per-kernel CUDA events) and the rooflines derived from it.
    full: also the blocking end-to-end arm and the YD16 packed-input arm.  Returns a dict (rank-local fields;
    the throughput values are whole-job: max over ranks of the time)."""
    pkg, B, torch, dist, args = ctx.pkg, ctx.B, ctx.torch, ctx.dist, ctx.args
    world, rank, local, stream, sampler = ctx.world, ctx.rank, ctx.local, ctx.stream, ctx.sampler
    mode = args.mode
    extra = {"icp_ppt": ppt} if ppt else {}
    if (Wd, Hd) != (W, H):
        extra.update(width=Wd, height=Hd, fx=570.3 * Wd / 640, fy=570.3 * Wd / 640, cx=Wd / 2.0, cy=Hd / 2.0)
    cfg = pkg.default_config(batch=batch, n_streams=S, device=local, traj_capacity=FRAMES, levels=levels,
                             iters=ITERS[:levels] + [0] * (4 - levels), stream=stream.cuda_stream, **extra)
    trk = B.Tracker(cfg)
    if mode == "model":  # frame-to-model tracking (include/youth_model.h): a sequence is a chain of frames
        trk.enable_model(pkg.tsdf_config())

    # independent sequences: rank r tracks sequences r*S .. r*S+S-1 (seed 20261018 + sequence)
    t0 = time.perf_counter()
    frames = np.stack([pkg.synth_sequence(FRAMES, Wd, Hd, sequence=rank * S + k) for k in range(S)])
    log(f"[rank {rank}] {label or 'main'}: generated {S} x {FRAMES} frames of {Wd}x{Hd} in {time.perf_counter() - t0:.1f}s")
    frame_bytes = Wd * Hd * 2
    seq_bytes = FRAMES * frame_bytes
    d_frames = torch.from_numpy(frames.view(np.int16)).to(f"cuda:{local}")
    d_base = d_frames.data_ptr()
    pin_ptr = trk.lib.youth_cuda_host_alloc(frames.nbytes)  # pinned host copy for the e2e arm
    if not pin_ptr:
        raise SystemExit("pinned allocation failed")
    C.memmove(pin_ptr, frames.ctypes.data, frames.nbytes)

    groups = [(a, min(batch, FRAMES - a)) for a in range(0, FRAMES, batch)]
    traj_view = gathered = None
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

    def barrier():
        trk.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        torch.cuda.synchronize()
        t = torch.tensor([x], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm
    for _ in range(warmup):
        step_device()
    barrier()
    launches0 = trk.launch_count()
    if sampler:
        sampler.mark_begin(label + "device")
    trk.timer_start()
    for _ in range(steps):
        step_device()
    ms_total = trk.timer_stop()
    barrier()
    if sampler:
        sampler.mark_end()
    launches = trk.launch_count() - launches0
    ms_per_step = max_over_ranks(ms_total) / steps
    value = world * S * FRAMES / (ms_per_step * 1e-3)

    # pose error vs synthetic ground truth (reported, not part of the timed region)
    poses, _, status = trk.trajectory()
    gt = pkg.synth_gt(FRAMES, Wd, Hd, sequence=rank * S)
    terr = np.linalg.norm(poses.reshape(-1, 3, 4)[:, :, 3] - gt.reshape(-1, 3, 4)[:, :, 3], axis=1)
    lost = int((status & B.STATUS_LOST != 0).sum())

    # ---- end-to-end arm (host buffers through the C ABI).  Every step copies its 300 frames per sequence from
    # pinned host memory and reads its trajectory back into pinned host memory, all inside the timed region.
    # Two steps in flight -- submit step i (youth_cuda_track_batch + youth_cuda_read_trajectory_async), then collect
    # step i-1 (youth_cuda_wait_ticket) -- so that the H2D copy of a step runs on the copy stream under the kernels
    # of the step before.
    host_traj = np.empty((S, FRAMES, 12), dtype=np.float32)
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

    def run_pipelined(n):
        for i in range(n):
            submit(i)
            if len(pending) > 1:
                collect()
        while pending:
            collect()

    run_pipelined(warmup)
    barrier()
    if sampler:
        sampler.mark_begin(label + "e2e")
    t0 = time.perf_counter()
    run_pipelined(steps)
    trk.sync()
    e2e_s = time.perf_counter() - t0
    if sampler:
        sampler.mark_end()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = world * S * FRAMES * steps / e2e_s
    e2e_matches = bool(np.array_equal(host_traj[0].view(np.uint32), poses.view(np.uint32)))

    out = {
        "value": value, "ms_per_step": ms_per_step, "launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": S * FRAMES * frame_bytes,
                "d2h_bytes_per_step": S * FRAMES * 48, "bit_identical_to_device_arm": e2e_matches,
                "steps_in_flight": 2,
                "how": "youth_cuda_track_batch (pinned host frames) + youth_cuda_read_trajectory_async per step, "
                       "youth_cuda_wait_ticket of the step before; H2D and D2H of every step inside the timed region"},
        "pose_error_vs_ground_truth": {"max_translation_m": float(terr.max()), "final_translation_m": float(terr[-1]),
                                       "frames_flagged_lost": lost},
        "poses": poses, "frames": frames, "cfg": cfg,
    }

    if full:
        def step_e2e():
            trk.reset()
            for a, n in groups:
                trk.track_batch_ptrs([pin_ptr + k * seq_bytes + a * frame_bytes for k in range(S)], n, B.MEM_HOST_PINNED)
            for k in range(S):
                got = trk.lib.youth_cuda_get_trajectory(trk.h, k, 0, FRAMES, host_traj[k].ctypes.data, None, None)
                assert got == FRAMES

        for _ in range(warmup):
            step_e2e()
        barrier()
        if sampler:
            sampler.mark_begin(label + "e2e_blocking")
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        trk.sync()
        blk_s = time.perf_counter() - t0
        if sampler:
            sampler.mark_end()
        blk_s = max_over_ranks(blk_s)
        out["e2e_blocking"] = {"value": world * S * FRAMES * steps / blk_s, "unit": "frames/s",
                               "how": "one step at a time: youth_cuda_track_batch then a blocking youth_cuda_get_trajectory"}
        out["e2e"]["bit_identical_to_device_arm"] = e2e_matches and bool(
            np.array_equal(host_traj[0].view(np.uint32), poses.view(np.uint32)))
    for p in res_pin:
        trk.lib.youth_cuda_host_free(p)

    # ---- packed-input arm: the same step fed from YD16 streams (include/youth_codec.h) in pinned host memory; the
    # packed bytes cross PCIe and are unpacked on the device.  Pipelined like e2e (two steps in flight): with
    # several GPUs behind one host memory system this is the feed that keeps the host side out of the way.
    want_packed = (full or (world > 1 and (Wd, Hd) == (W, H))) and not args.no_packed and mode != "model"
    enc = None
    if want_packed:
        # the set-up (encoding the step's frames once) is local work: if it fails on ANY rank every rank skips the
        # arm -- the timed part below contains collectives, and the headline arms above must not be lost to it
        why = ""
        try:
            cd = pkg.Codec(Wd, Hd, max_frames=FRAMES, device=local)
            enc = [cd.encode_ptr(d_base + k * seq_bytes, FRAMES, B.MEM_DEVICE) for k in range(S)]
            enc_ms = cd.last_kernel_ms()
            back = torch.empty(FRAMES * Wd * Hd, dtype=torch.int16, device=f"cuda:{local}")
            cd.decode_to_device(enc[0][0], enc[0][1], back.data_ptr())
            dec_ms = cd.last_kernel_ms()
            codec_ok = bool(torch.equal(back.view(FRAMES, Hd, Wd), d_frames[0]))
            del back
            cd.close()
        except Exception as e:
            enc, why = None, f"{type(e).__name__}: {e}"
        if max_over_ranks(0.0 if enc is not None else 1.0) > 0.0:
            out["packed_input"] = {"error": why or "set-up failed on another rank", "what": "YD16 packed-input arm skipped"}
            enc = None
    if enc is not None:
        pk_bytes = [int(e[1][-1]) for e in enc]
        pk_pin = [trk.lib.youth_cuda_host_alloc(b) for b in pk_bytes]
        for ptr, e in zip(pk_pin, enc):
            C.memmove(ptr, e[0].ctypes.data, len(e[0]))
        pk_ptrs = (C.c_void_p * S)(*pk_pin)
        offs_all = [[np.ascontiguousarray(e[1][a:a + n + 1]) for e in enc] for a, n in groups]
        pres_pin = [trk.lib.youth_cuda_host_alloc(S * FRAMES * 48) for _ in range(2)]
        pres_np = [np.ctypeslib.as_array((C.c_float * (S * FRAMES * 12)).from_address(p)).reshape(S, FRAMES, 12) for p in pres_pin]

        def submit_packed(i):
            trk.reset()
            for (a, n), offs in zip(groups, offs_all):
                op = (C.c_void_p * S)(*[o.ctypes.data for o in offs])
                ok = trk.lib.youth_cuda_track_batch_packed(trk.h, pk_ptrs, op, n, B.MEM_HOST_PINNED, None, None)
                assert ok, trk.lib.youth_cuda_last_error()
            ticket = None
            for k in range(S):
                got, ticket = trk.read_trajectory_async(pres_pin[i % 2] + k * FRAMES * 48, FRAMES, stream=k)
                assert got == FRAMES
            pending.append((i, ticket))

        def collect_packed():
            i, ticket = pending.popleft()
            trk.wait_ticket(ticket)
            host_traj[...] = pres_np[i % 2]

        def run_packed(n):
            for i in range(n):
                submit_packed(i)
                if len(pending) > 1:
                    collect_packed()
            while pending:
                collect_packed()

        run_packed(warmup)
        barrier()
        psteps = steps if world > 1 else max(2, min(steps, 20))
        t0 = time.perf_counter()
        run_packed(psteps)
        trk.sync()
        pk_s = max_over_ranks(time.perf_counter() - t0)
        raw_b = FRAMES * frame_bytes
        out["packed_input"] = {
            "e2e_value": world * S * FRAMES * psteps / pk_s, "unit": "frames/s", "steps": psteps, "steps_in_flight": 2,
            "h2d_bytes_per_step": int(sum(pk_bytes)), "d2h_bytes_per_step": S * FRAMES * 48,
            "compression_ratio": S * raw_b / float(sum(pk_bytes)),
            "bit_identical_to_device_arm": bool(np.array_equal(host_traj[0].view(np.uint32), poses.view(np.uint32))),
            "codec_round_trip_identical": codec_ok,
            "k_yd16_encode": {"ms_per_300_frames": enc_ms, "gbs": (raw_b + pk_bytes[-1]) / (enc_ms * 1e-3) / 1e9},
            "k_yd16_decode": {"ms_per_300_frames": dec_ms, "gbs": (raw_b + pk_bytes[0]) / (dec_ms * 1e-3) / 1e9},
            "what": "youth_cuda_track_batch_packed: YD16 streams in pinned host memory, unpacked on the device; "
                    "trajectory read back through the ticket API",
        }
        for ptr in pk_pin + pres_pin:
            trk.lib.youth_cuda_host_free(ptr)

    # ---- per-kernel timing (separate profiled step: events around every launch)
    trk.profile(True)
    step_device()
    trk.sync()
    prof_ms, prof_n = trk.profile_read()
    trk.profile(False)
    out.update(rooflines(B, prof_ms, prof_n, Wd, Hd, levels, S, groups, batch, ppt, mode))

    trk.lib.youth_cuda_host_free(pin_ptr)
    trk.close()
    del d_frames
    torch.cuda.empty_cache()
    return out


def rooflines(B, prof_ms, prof_n, Wd, Hd, levels, S, groups, batch, ppt, mode):
    """Per kernel class of one profiled step: average launch duration (CUDA events), the bytes DRAM must move per
    launch (compulsory), the bytes the kernel requests (k_icp: 48 B/px/iteration), and the
    fraction of the measured HBM peak the compulsory bytes run at.  No remembered numbers: everything here is
    computed from this run's event times, except `traffic` (ncu capture of the same kernels, see load_traffic)."""
    peak, peak_kind = measured_peak()
    names = {B.PROF_INGEST: "k_ingest", B.PROF_NORMALS: "k_normals", B.PROF_ICP0: "k_icp_L0",
             B.PROF_ICP0 + 1: "k_icp_L1", B.PROF_ICP0 + 2: "k_icp_L2", B.PROF_ICP0 + 3: "k_icp_L3",
             B.PROF_SOLVE: "k_tsdf_integrate", B.PROF_MISC: "k_compose", B.PROF_RAYCAST: "k_tsdf_raycast"}
    step_prof = {names[i]: {"ms": round(float(prof_ms[i]), 4), "launches": int(prof_n[i])}
                 for i in range(B.PROF_CLASSES) if prof_n[i]}
    geom_key = f"{Wd}x{Hd}_L{levels}_S{S}_B{batch}_ppt{ppt or 64}_{mode}"
    traffic, traffic_source = load_traffic(geom_key)
    # pairs / frames per launch, averaged over the launch groups of a step: group g of n frames per sequence holds
    # n pairs (n - 1 in the first group: frame 0 has no predecessor) and touches pairs + 1 frames per sequence
    pairs = [S * (n - (1 if a == 0 else 0)) for a, n in groups]
    avg_pairs = sum(pairs) / len(pairs)
    avg_frames = sum(p + S for p in pairs) / len(pairs)
    rows = []
    for l in range(levels):
        cls = B.PROF_ICP0 + l
        if not prof_n[cls]:
            continue
        npix = (Wd >> l) * (Hd >> l)
        ms = float(prof_ms[cls]) / int(prof_n[cls])
        comp = MAP_BYTES_PER_PX * npix * avg_frames
        req = REQ_BYTES_PER_PX_ITER * npix * avg_pairs
        gbs = comp / (ms * 1e-3) / 1e9
        rows.append({"bound": "hbm", "kernel": f"k_icp (level {l})", "achieved": gbs, "peak": peak, "unit": "GB/s",
                     "frac": gbs / peak, "traffic": traffic.get(f"k_icp_L{l}"), "avg_launch_ms": ms,
                     "launches_per_step": int(prof_n[cls]), "pairs_per_launch": avg_pairs,
                     "compulsory_bytes_per_launch": comp, "requested_bytes_per_launch": req,
                     "requested_gbs": req / (ms * 1e-3) / 1e9})
    # stages 1-2: raw depth in + vertex/normal maps of all levels out (SURVEY 8(d) B_pre), k_ingest + k_normals together
    npix_all = sum((Wd >> l) * (Hd >> l) for l in range(levels))
    b_pre = 2 * Wd * Hd + MAP_BYTES_PER_PX * npix_all
    n_ing = max(1, int(prof_n[B.PROF_INGEST]))
    ing_ms = float(prof_ms[B.PROF_INGEST] + prof_ms[B.PROF_NORMALS]) / n_ing
    ing_frames = S * FRAMES / n_ing
    ing_gbs = b_pre * ing_frames / (ing_ms * 1e-3) / 1e9 if ing_ms > 0 else 0.0
    tr_ing = None
    if traffic.get("k_ingest") is not None:
        tr_ing = traffic.get("k_ingest") + (traffic.get("k_normals") or 0)
    rows.append({"bound": "hbm", "kernel": "k_ingest + k_normals", "achieved": ing_gbs, "peak": peak, "unit": "GB/s",
                 "frac": ing_gbs / peak, "traffic": tr_ing, "avg_launch_ms": ing_ms, "launches_per_step": n_ing,
                 "frames_per_launch": ing_frames, "compulsory_bytes_per_launch": b_pre * ing_frames,
                 "algorithmic_bytes_per_frame": b_pre})
    top = dict(rows[0]) if rows else {}
    top.update({"peak_source": peak_kind, "traffic_source": traffic_source,
                "algorithmic_bytes_per_launch": top.get("compulsory_bytes_per_launch"),
                "note": "frac = compulsory DRAM bytes (24 B/px of every frame a launch touches, moved once) / CUDA-event "
                        "launch time / measured peak.  The kernel REQUESTS 48 B/px/iteration (requested_bytes_per_launch: "
                        "every frame is the current frame of one pair and the previous frame of the next; the second use "
                        "is served by L2).  traffic = ncu dram bytes per launch of the same kernel sources, or null."})
    return {"roofline": top, "roofline_all": rows, "per_kernel_ms_per_step": step_prof}


def parity_check(ctx, res, n_frames):
    """The device trajectory of the timed workload against the CPU oracle's PARITY build (-O2 -ffp-contract=off),
    pairs spread over the host threads: float poses must be bit-identical for the first n_frames frames."""
    O = oracle_module()
    ocfg = O.config_from(res["cfg"])
    frames = np.ascontiguousarray(res["frames"][0][:n_frames])
    ref, st, _, secs = O.track_sequence_parallel(ocfg, frames, threads=cpu_threads())
    got = res["poses"][:n_frames]
    same = np.all(got.view(np.uint32) == ref.view(np.uint32), axis=1)
    dt = np.linalg.norm(got.reshape(-1, 3, 4)[:, :, 3] - ref.reshape(-1, 3, 4)[:, :, 3], axis=1)
    return {"frames_compared": int(n_frames), "frames_bit_identical": int(same.sum()),
            "bit_identical": bool(same.all()), "max_translation_diff_m": float(dt.max()),
            "oracle_build": "-O2 -ffp-contract=off parity build, pairs over host threads", "seconds": round(secs, 2)}


def slim(res):
    return {k: v for k, v in res.items() if k not in ("poses", "frames", "cfg")}


def run_ours(args):
    import torch

    import youth_pkg

    pkg = youth_pkg.load()
    from slam_rgbd_b200 import binding as B

    ctx = Ctx()
    ctx.pkg, ctx.B, ctx.torch, ctx.args = pkg, B, torch, args
    world = ctx.world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = ctx.rank = int(os.environ.get("RANK", "0"))
    local = ctx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the tracking path has no CPU fallback")
    torch.cuda.set_device(local)
    # bind BEFORE any pinned allocation: first touch places the pinned frames next to the GPU
    placement = bind_to_gpu_cpus(local) if world > 1 and not args.no_numa else "single rank: not bound"
    log(f"[rank {rank}] {placement}")
    ctx.dist = None
    if world > 1:
        import torch.distributed as dist_mod

        ctx.dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        ctx.dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx.stream = torch.cuda.Stream()
    ctx.sampler = ClockSampler(local) if rank == 0 else None
    if ctx.sampler:
        ctx.sampler.start()  # nvidia-smi needs ~1 s to start printing; only in-region samples are kept

    if args.mode == "model" and args.batch == 300:
        args.batch = 20  # frames per call; the chain runs frame by frame anyway and the ring only needs a few slots
    Wd, Hd, S = args.width, args.height, args.sequences_per_gpu
    ppt = default_ppt(args, args.batch, S)
    main = measure(ctx, Wd, Hd, args.levels, S, args.batch, ppt, args.steps, args.warmup, full=True)

    # ---- the other BASELINE.json configurations, short arms at every N (nested under extra_configs)
    extras = {}
    if not args.no_extra and args.mode == "frame" and (Wd, Hd, S, args.levels) == (W, H, 1, 3):
        xs = max(3, min(args.steps, 8))
        r3 = measure(ctx, W, H, 3, 8, 300, 128, xs, 3, label="configs3.")
        extras["configs3"] = dict(slim(r3), config=config_dict(W, H, 3, 8, 300, 128, world), steps=xs, warmup=3, clocks=None)
        pk3 = r3.get("packed_input") or {}
        if world > 1 and "e2e_value" in pk3 and pk3.get("bit_identical_to_device_arm"):
            extras["configs3"]["e2e_raw_frames"] = extras["configs3"]["e2e"]
            extras["configs3"]["e2e"] = {"value": pk3["e2e_value"], "unit": "frames/s", "h2d_bytes_per_step": pk3["h2d_bytes_per_step"],
                                         "d2h_bytes_per_step": pk3["d2h_bytes_per_step"], "bit_identical_to_device_arm": True,
                                         "input": "YD16-packed host frames, unpacked on the device"}
        del r3
        r4 = measure(ctx, 1280, 960, 4, 1, 300, 128, xs, 3, label="configs4.")
        extras["configs4"] = dict(slim(r4), config=config_dict(1280, 960, 4, 1, 300, 128, world), steps=xs, warmup=3,
                                  vga_equivalent_frames_per_sec=r4["value"] * 4.0, clocks=None)
        del r4

    if ctx.sampler:
        ctx.sampler.stop()
        for k, pre in (("configs3", "configs3."), ("configs4", "configs4.")):
            if k in extras:
                extras[k]["clocks"] = ctx.sampler.summary(pre)
    if rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return 0
    clocks = ctx.sampler.summary("")
    frames, cfg, poses = main["frames"], main["cfg"], main["poses"]

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

    # ---- the reference-facing facade (include/SLAM.h)
    facade = facade_arm(pkg, args, local, frames, poses, ppt) if (Wd, Hd, S) == (W, H, 1) and args.mode != "model" and args.levels == 3 else None

    # ---- CPU baseline (the one sampler) and parity of the timed trajectory against the oracle's parity build
    O = oracle_module()
    ocfg = O.config_from(cfg)
    threads = cpu_threads()
    fpt = CPU_SAMPLE_FRAMES_PER_THREAD
    parity = None
    if args.mode == "model":  # the frame-to-model statement (fusion + ray cast on top of the same ICP), parity build
        mt = O.tsdf_config_from(pkg.tsdf_config())
        fpt = 12
        cpu_fps, cpu_wall, cpu_kind = cpu_sample(frames[0], ocfg, threads, fpt, tsdf_cfg=mt)
        cpu_kind = "frame-to-model statement, " + cpu_kind
        cpu1_fps, _, _ = cpu_sample(frames[0], ocfg, 1, 4, tsdf_cfg=mt)
        cpu1p_fps = cpu1_fps
    else:
        cpu_sample(frames[0], ocfg, threads, 3)  # untimed: first touch of the oracle's buffers, library load
        cpu_fps, cpu_wall, cpu_kind = cpu_sample(frames[0], ocfg, threads, fpt)
        cpu1_fps, _, _ = cpu_sample(frames[0], ocfg, 1, 6)               # one thread, speed build
        cpu1p_fps, _, _ = cpu_sample(frames[0], ocfg, 1, 6, fast=False)  # one thread, parity build
        if not args.no_parity:
            parity = parity_check(ctx, main, min(FRAMES, args.parity_frames))

    # Several GPUs behind one host memory system: raw frames cost 184 MB of H2D per GPU and step, and eight ranks
    # copying at once are bound by the host side (round 1: 0.963 scaling efficiency end to end).  With more than
    # one rank the end-to-end feed is therefore the YD16-packed host input (a third of the bytes, unpacked on the
    # device inside the timed region, bit-identical frames and trajectory); the raw-frame number stays in the line
    # as e2e_raw_frames.  One rank: raw frames (packed_input reports the packed feed).
    e2e_line = main["e2e"]
    pk = main.get("packed_input") or {}
    if world > 1 and "e2e_value" in pk and pk.get("bit_identical_to_device_arm"):
        e2e_line = {"value": pk["e2e_value"], "unit": "frames/s", "h2d_bytes_per_step": pk["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": pk["d2h_bytes_per_step"], "bit_identical_to_device_arm": True,
                    "steps_in_flight": 2, "steps": pk["steps"],
                    "input": "YD16-packed host frames (lossless, include/youth_codec.h), unpacked on the device",
                    "how": "youth_cuda_track_batch_packed (pinned host streams) + youth_cuda_read_trajectory_async per step, "
                           "youth_cuda_wait_ticket of the step before; H2D, unpacking and D2H of every step inside the timed region"}
    line = {
        "metric": "icp_tracked_frames_per_sec", "value": main["value"], "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(Wd, Hd, args.levels, S, args.batch, ppt, world, args.mode),
        "host_placement": placement,
        "clocks": clocks,
        "e2e": e2e_line,
        "e2e_raw_frames": main["e2e"] if e2e_line is not main["e2e"] else None,
        "e2e_blocking": main.get("e2e_blocking"),
        "gpu_launches": main["launches"],
        "roofline": main["roofline"],
        "roofline_all": main["roofline_all"],
        "per_kernel_ms_per_step": main["per_kernel_ms_per_step"],
        "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(threads, fpt, cpu_kind) + f", {cpu_wall:.1f}s wall",
                         "single_thread_value": cpu1_fps, "single_thread_parity_build_value": cpu1p_fps},
        "parity_vs_oracle": parity,
        "pose_error_vs_ground_truth": main["pose_error_vs_ground_truth"],
        "frames_per_sec_per_gpu": main["value"] / world,
        "streaming_single_frame": streaming,
        "facade": facade,
        "facade_zero_copy": (facade or {}).get("pinned_runs"),
        "packed_input": main.get("packed_input"),
        "extra_configs": extras or None,
    }
    emit(line)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()
    return 0


def facade_arm(pkg, args, local, frames, poses, ppt):
    """One processSlamFrame() call per frame from a producer thread (synchronous copy into the pinned host ring,
    SLAM.cpp:133-134), worker thread with two runs in flight, youthSlamDrain + trajectory read-back at the end of
    every pass; and the zero-copy producer calls (youthSlamAcquireSlot / youthSlamCommitSlot) next to it."""
    out = {}
    try:
        os.environ["YOUTH_SLAM_DEVICE"] = str(local)
        os.environ["YOUTH_SLAM_TRAJ_CAPACITY"] = str(FRAMES)
        if ppt:
            os.environ["YOUTH_SLAM_ICP_PPT"] = str(ppt)  # same reduction geometry as the device arm
        host = pkg.host_lib()
        group = int(os.environ.get("YOUTH_BENCH_FACADE_GROUP", "150"))
        host.youthSlamSetOptions(1, group)  # lossless, frames per launch group
        host.initSlamModule(None, None)
        if host.isSlamModuleRunning() != 1:
            return {"error": "initSlamModule failed"}
        fposes = np.empty((FRAMES, 12), dtype=np.float32)
        ptrs = [frames[0][i].ctypes.data for i in range(FRAMES)]
        fb = W * H * 2

        def pass_copy():
            host.resetSlam()
            for i in range(FRAMES):
                assert host.processSlamFrame(ptrs[i], None, W, H, 33 * i) == 1
            host.youthSlamDrain()
            assert host.youthSlamGetTrajectory(fposes.ctypes.data, None, None, FRAMES) == FRAMES

        def pass_zero_copy():
            host.resetSlam()
            for i in range(FRAMES):
                slot = host.youthSlamAcquireSlot(W, H)
                assert slot
                C.memmove(slot, ptrs[i], fb)  # stands for the producer writing the frame where it is born (chunk reassembly)
                assert host.youthSlamCommitSlot(33 * i) == 1
            host.youthSlamDrain()
            assert host.youthSlamGetTrajectory(fposes.ctypes.data, None, None, FRAMES) == FRAMES

        # frames that already live in page-locked memory: no CPU copy at all (youthSlamProcessPinnedFrames)
        nb = FRAMES * fb
        pin = pkg.cuda_lib().youth_cuda_host_alloc(nb)
        assert pin
        C.memmove(pin, frames[0].ctypes.data, nb)
        ts = np.arange(FRAMES, dtype=np.uint32) * 33

        def pass_pinned():
            host.resetSlam()
            assert host.youthSlamProcessPinnedFrames(pin, FRAMES, W, H, ts.ctypes.data) == 1
            assert host.youthSlamGetTrajectory(fposes.ctypes.data, None, None, FRAMES) == FRAMES

        fsteps = max(1, min(args.steps, 10))
        for name, fn in (("processSlamFrame", pass_copy), ("zero_copy_slot", pass_zero_copy), ("pinned_runs", pass_pinned)):
            for _ in range(2):
                fn()
            t0 = time.perf_counter()
            for _ in range(fsteps):
                fn()
            fdt = time.perf_counter() - t0
            out[name] = {"frames_per_sec": FRAMES * fsteps / fdt, "steps": fsteps,
                         "bit_identical_to_device_arm": bool(np.array_equal(fposes.view(np.uint32), poses.view(np.uint32)))}
        pkg.cuda_lib().youth_cuda_host_free(pin)
        out["frames_per_sec"] = out["processSlamFrame"]["frames_per_sec"]
        out["bit_identical_to_device_arm"] = out["processSlamFrame"]["bit_identical_to_device_arm"]
        out["copy_threads"] = int(os.environ.get("YOUTH_SLAM_COPY_THREADS", "4"))
        out["what"] = (f"SLAM.h facade, lossless, {group} frames per launch group, two groups in flight.  processSlamFrame: "
                       "one call per frame, the pageable host frame is copied into the pinned ring before the call returns "
                       "(SLAM.cpp:133-134), the copy shared by copy_threads host threads.  zero_copy_slot: "
                       "youthSlamAcquireSlot / youthSlamCommitSlot, the producer fills the ring slot itself (here: one "
                       "single-threaded memmove per frame standing for chunk reassembly).  pinned_runs: "
                       "youthSlamProcessPinnedFrames, the frames already live in page-locked memory and are read in place")
        host.stopSlamModule()
    except Exception as e:  # the facade arm is informative only
        out["error"] = f"{type(e).__name__}: {e}"
    return out


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
    ap.add_argument("--no-numa", action="store_true", help="do not bind ranks to the CPUs next to their GPU")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3] / configs[4] arms")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check of the timed trajectory against the oracle")
    ap.add_argument("--parity-frames", type=int, default=FRAMES, help="frames of the timed sequence compared with the oracle")
    ap.add_argument("--mode", default="frame", choices=["frame", "model"],
                    help="frame = frame-to-frame (the headline workload), model = frame-to-model (TSDF fusion + ray cast)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.user_ppt = args.ppt
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
