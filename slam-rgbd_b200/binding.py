"""ctypes view of include/youth_cuda.h and include/youth_host.h (no arithmetic here)."""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 4
SUM_SLOTS = 32

MEM_HOST, MEM_DEVICE, MEM_HOST_PINNED = 0, 1, 2
DBG_DEPTH, DBG_VERTEX, DBG_NORMAL, DBG_MASK, DBG_PYRCNT = 1, 2, 3, 4, 5
STATUS_FIRST, STATUS_LOST = 1, 2
PROF_INGEST, PROF_NORMALS, PROF_ICP0, PROF_SOLVE, PROF_MISC, PROF_RAYCAST, PROF_CLASSES = 0, 1, 2, 6, 7, 8, 9
PROF_INTEGRATE = PROF_SOLVE


class CudaLibraryMissing(RuntimeError):
    pass


class YouthConfig(C.Structure):
    """Mirror of ``youth_cuda_config`` (include/youth_cuda.h)."""

    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("depth_factor", C.c_float),
        ("levels", C.c_int32),
        ("iters", C.c_int32 * MAX_LEVELS),
        ("depth_min_mm", C.c_int32), ("depth_max_mm", C.c_int32),
        ("bilateral", C.c_int32),
        ("sigma_space_px", C.c_float), ("sigma_range_mm", C.c_float),
        ("dist_thresh_m", C.c_float), ("cos_thresh", C.c_float),
        ("min_inliers", C.c_int32),
        ("icp_ppt", C.c_int32),
        ("n_streams", C.c_int32),
        ("batch", C.c_int32),
        ("traj_capacity", C.c_int32),
        ("device", C.c_int32),
        ("stream", C.c_void_p),
    ]


class TsdfConfig(C.Structure):
    """Mirror of ``youth_tsdf_config`` (include/youth_model.h)."""

    _fields_ = [("dim", C.c_int32 * 3), ("voxel_m", C.c_float), ("origin", C.c_float * 3), ("trunc_m", C.c_float),
                ("max_weight", C.c_int32), ("near_m", C.c_float), ("far_m", C.c_float)]


class SynthConfig(C.Structure):
    """Mirror of ``youth_synth_config`` (include/youth_host.h)."""

    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("seed", C.c_uint32),
        ("period", C.c_int32),
        ("phase", C.c_double),
        ("dropout", C.c_double),
        ("noise", C.c_int32),
        ("dmin_mm", C.c_int32), ("dmax_mm", C.c_int32),
    ]


def lib_paths():
    return {
        # YOUTH_CUDA_LIB: A/B builds of the same source (tools/ only); never a non-CUDA substitute
        "cuda": os.environ.get("YOUTH_CUDA_LIB") or os.path.join(PKG_DIR, "lib", "libyouth_cuda.so"),
        "host": os.path.join(PKG_DIR, "lib", "libAlgorithmModule.so"),
        "synth": os.path.join(PKG_DIR, "lib", "libyouth_synth.so"),
        "harness": os.path.join(PKG_DIR, "bin", "youth_harness"),
    }


_cuda = None
_host = None
_synth = None


def synth_lib():
    """The synthetic sequence generator (host/youth_synth.c) as a library of its own: plain C, links no CUDA, so
    callers that must not map the GPU code (bench.py --impl reference) can still generate the workload."""
    global _synth
    if _synth is not None:
        return _synth
    path = lib_paths()["synth"]
    if not os.path.exists(path):
        raise CudaLibraryMissing(f"{path} is missing: build the package first")
    L = C.CDLL(path)
    SC = C.POINTER(SynthConfig)
    for name, (res, args) in {
        "youth_synth_default": (None, [SC, C.c_int, C.c_int, C.c_int]),
        "youth_synth_gt": (None, [SC, C.c_int, C.c_void_p]),
        "youth_synth_sequence": (None, [SC, C.c_int, C.c_int, C.c_void_p]),
    }.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _synth = L
    return L


def cuda_lib():
    """Load libyouth_cuda.so; raises loudly when the extension has not been built."""
    global _cuda
    if _cuda is not None:
        return _cuda
    path = lib_paths()["cuda"]
    if not os.path.exists(path):
        raise CudaLibraryMissing(
            f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the tracking path.")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    H = C.c_void_p
    u16pp = C.POINTER(C.c_void_p)
    sig = {
        "youth_cuda_default_config": (C.c_int, [C.POINTER(YouthConfig)]),
        "youth_cuda_init": (C.c_int, [C.POINTER(YouthConfig), C.POINTER(H)]),
        "youth_cuda_destroy": (None, [H]),
        "youth_cuda_track": (C.c_int, [H, C.c_void_p, C.c_uint32, C.c_void_p]),
        "youth_cuda_track_batch": (C.c_int, [H, u16pp, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
        "youth_cuda_set_icp_schedule": (C.c_int, [H, C.c_int, C.c_int]),
        "youth_cuda_sync": (C.c_int, [H]),
        "youth_cuda_reset": (C.c_int, [H, C.c_int]),
        "youth_cuda_frame_count": (C.c_int, [H, C.c_int]),
        "youth_cuda_get_trajectory": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
        "youth_cuda_read_trajectory_async": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
        "youth_cuda_wait_ticket": (C.c_int, [H, C.c_uint64]),
        "youth_cuda_read_last_inliers_async": (C.c_int, [H, C.c_int, C.c_void_p]),
        "youth_cuda_last_inliers": (C.c_int, [H, C.c_int]),
        "youth_cuda_trajectory_device_ptr": (C.c_void_p, [H, C.c_int]),
        "youth_cuda_device_count": (C.c_int, []),
        "youth_cuda_stream": (C.c_void_p, [H]),
        "youth_cuda_set_device": (C.c_int, [C.c_int]),
        "youth_cuda_device_alloc": (C.c_void_p, [C.c_size_t]),
        "youth_cuda_device_free": (None, [C.c_void_p]),
        "youth_cuda_copy_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "youth_cuda_copy_to_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "youth_cuda_device_sync": (C.c_int, []),
        "youth_cuda_host_alloc": (C.c_void_p, [C.c_size_t]),
        "youth_cuda_host_free": (None, [C.c_void_p]),
        "youth_cuda_debug_enable_maps": (C.c_int, [H]),
        "youth_cuda_debug_read": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
        "youth_cuda_debug_icp": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
        "youth_cuda_debug_rcp_check": (C.c_longlong, [H, C.c_uint32, C.c_uint32]),
        "youth_cuda_timer_start": (C.c_int, [H]),
        "youth_cuda_timer_stop": (C.c_int, [H, C.POINTER(C.c_float)]),
        "youth_cuda_launch_count": (C.c_uint64, [H]),
        "youth_cuda_profile_enable": (C.c_int, [H, C.c_int]),
        "youth_cuda_profile_read": (C.c_int, [H, C.c_void_p, C.c_void_p]),
        "youth_cuda_last_error": (C.c_char_p, []),
        "youth_cuda_abi_version": (C.c_int, []),
        # include/youth_model.h
        "youth_tsdf_default_config": (C.c_int, [C.POINTER(TsdfConfig)]),
        "youth_cuda_enable_model": (C.c_int, [H, C.POINTER(TsdfConfig)]),
        "youth_cuda_model_enabled": (C.c_int, [H]),
        "youth_cuda_model_surface_voxels": (C.c_longlong, [H, C.c_int]),
        "youth_cuda_debug_read_volume": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t]),
        "youth_cuda_debug_read_model": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
        "youth_cuda_debug_integrate": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p]),
        "youth_cuda_debug_raycast": (C.c_int, [H, C.c_int, C.c_void_p, C.c_int]),
        # include/youth_codec.h
        "youth_codec_max_bytes": (C.c_size_t, [C.c_int, C.c_int]),
        "youth_codec_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(H)]),
        "youth_codec_destroy": (None, [H]),
        "youth_codec_encode": (C.c_int, [H, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
        "youth_codec_decode": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
        "youth_codec_last_kernel_ms": (C.c_float, [H]),
        "youth_codec_launch_count": (C.c_uint64, [H]),
        "youth_cuda_track_batch_packed": (C.c_int, [H, u16pp, u16pp, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError = header/library drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _cuda = L
    return L


def host_lib():
    """Load libAlgorithmModule.so (C facade + host helpers)."""
    global _host
    if _host is not None:
        return _host
    cuda_lib()
    path = lib_paths()["host"]
    if not os.path.exists(path):
        raise CudaLibraryMissing(f"{path} is missing: build the package first")
    L = C.CDLL(path)
    SC = C.POINTER(SynthConfig)
    sig = {
        "youth_synth_default": (None, [SC, C.c_int, C.c_int, C.c_int]),
        "youth_synth_pose": (None, [SC, C.c_int, C.c_void_p]),
        "youth_synth_gt": (None, [SC, C.c_int, C.c_void_p]),
        "youth_synth_frame": (None, [SC, C.c_int, C.c_void_p]),
        "youth_synth_sequence": (None, [SC, C.c_int, C.c_int, C.c_void_p]),
        "youth_config_from_yaml": (C.c_int, [C.c_char_p, C.POINTER(YouthConfig)]),
        "youth_pose_to_quat": (None, [C.c_void_p, C.c_void_p]),
        "youth_tum_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]),
        "youth_chunk_count": (C.c_int, [C.c_size_t]),
        "youth_chunk_build": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                           C.c_size_t, C.c_int]),
        "youth_pose_msg_build": (C.c_size_t, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]),
        "youth_pose_msg_parse": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_uint32),
                                           C.c_void_p]),
        "youth_reasm_create": (C.c_void_p, []),
        "youth_reasm_destroy": (None, [C.c_void_p]),
        "youth_reasm_feed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "youth_reasm_depth": (C.c_void_p, [C.c_void_p]),
        "youth_reasm_color": (C.c_void_p, [C.c_void_p]),
        "youth_reasm_info": (None, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_uint32)]),
        "youth_mq_consume": (C.c_long, [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.c_int]),
        # facade (include/SLAM.h, algorithmModule.h)
        "initSlamModule": (None, [C.c_char_p, C.c_char_p]),
        "stopSlamModule": (None, []),
        "processSlamFrame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32]),
        "saveSlamMap": (C.c_int, [C.c_char_p]),
        "isSlamModuleRunning": (C.c_int, []),
        "getSlamMapPoints": (C.c_int, []),
        "resetSlam": (None, []),
        "algorithmModule": (C.c_void_p, [C.c_void_p]),
        "youthSlamSetOptions": (None, [C.c_int, C.c_int]),
        "youthSlamDrain": (None, []),
        "youthSlamGetTrajectory": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
        "youthSlamStats": (None, [C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]),
        "youthSlamAcquireSlot": (C.c_void_p, [C.c_int, C.c_int]),
        "youthSlamCommitSlot": (C.c_int, [C.c_uint32]),
        "youthSlamAbortSlot": (None, []),
        "youthSlamProcessPinnedFrames": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _host = L
    return L


def default_config(**overrides) -> YouthConfig:
    cfg = YouthConfig()
    if not cuda_lib().youth_cuda_default_config(C.byref(cfg)):
        raise RuntimeError("youth_cuda_default_config failed")
    for k, v in overrides.items():
        if k == "iters":
            for i, it in enumerate(v):
                cfg.iters[i] = int(it)
        else:
            setattr(cfg, k, v)
    return cfg


def tsdf_config(**overrides) -> TsdfConfig:
    t = TsdfConfig()
    if not cuda_lib().youth_tsdf_default_config(C.byref(t)):
        raise RuntimeError("youth_tsdf_default_config failed")
    for k, v in overrides.items():
        if k in ("dim", "origin"):
            for i, x in enumerate(v):
                getattr(t, k)[i] = x
        else:
            setattr(t, k, v)
    return t


def synth_config(width=640, height=480, sequence=0, noise=0) -> SynthConfig:
    sc = SynthConfig()
    synth_lib().youth_synth_default(C.byref(sc), width, height, sequence)
    sc.noise = noise
    return sc


def synth_sequence(n, width=640, height=480, sequence=0, noise=0, first=0) -> np.ndarray:
    """uint16 [n][height][width] synthetic depth (mm) from the C generator."""
    sc = synth_config(width, height, sequence, noise)
    out = np.empty((n, height, width), dtype=np.uint16)
    synth_lib().youth_synth_sequence(C.byref(sc), first, n, out.ctypes.data)
    return out


def synth_gt(n, width=640, height=480, sequence=0) -> np.ndarray:
    """float64 [n][12] ground-truth camera poses relative to frame 0."""
    sc = synth_config(width, height, sequence)
    out = np.empty((n, 12), dtype=np.float64)
    for i in range(n):
        synth_lib().youth_synth_gt(C.byref(sc), i, out[i].ctypes.data)
    return out


class Tracker:
    """Thin object view of a ``youth_cuda_handle``."""

    def __init__(self, cfg: YouthConfig):
        self.lib = cuda_lib()
        self.cfg = cfg
        self.h = C.c_void_p()
        if not self.lib.youth_cuda_init(C.byref(cfg), C.byref(self.h)):
            raise RuntimeError("youth_cuda_init failed: " + self.lib.youth_cuda_last_error().decode())

    def close(self):
        if self.h:
            self.lib.youth_cuda_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, ok, what):
        if not ok:
            raise RuntimeError(f"{what} failed: " + self.lib.youth_cuda_last_error().decode())

    def level_shape(self, level):
        return self.cfg.height >> level, self.cfg.width >> level

    def track(self, frame: np.ndarray, ts=0, want_pose=True):
        assert frame.dtype == np.uint16 and frame.flags.c_contiguous
        pose = np.empty(12, dtype=np.float32) if want_pose else None
        self._check(self.lib.youth_cuda_track(self.h, frame.ctypes.data, ts, pose.ctypes.data if want_pose else None),
                    "youth_cuda_track")
        return pose

    def set_icp_schedule(self, pairs_per_group, queues=1):
        self._check(self.lib.youth_cuda_set_icp_schedule(self.h, pairs_per_group, queues), "youth_cuda_set_icp_schedule")

    def track_batch_ptrs(self, ptrs, n, mem_kind, ts=None, poses_out=None):
        arr = (C.c_void_p * len(ptrs))(*ptrs)
        self._check(self.lib.youth_cuda_track_batch(
            self.h, arr, n, mem_kind,
            ts.ctypes.data if ts is not None else None,
            poses_out.ctypes.data if poses_out is not None else None), "youth_cuda_track_batch")

    def track_batch(self, frames, want_poses=True):
        """frames: list (one per stream) of uint16 [n][H][W] host arrays."""
        n = frames[0].shape[0]
        for f in frames:
            assert f.dtype == np.uint16 and f.flags.c_contiguous and f.shape[0] == n
        poses = np.empty((len(frames), n, 12), dtype=np.float32) if want_poses else None
        self.track_batch_ptrs([f.ctypes.data for f in frames], n, MEM_HOST, None, poses)
        return poses

    def track_batch_packed(self, streams, offsets, n, mem_kind=MEM_HOST, ts=None, want_poses=True):
        """streams / offsets: one uint8 array and one uint64 [n+1] array per sequence (YD16, back to back)."""
        sp = (C.c_void_p * len(streams))(*[a.ctypes.data for a in streams])
        op = (C.c_void_p * len(offsets))(*[a.ctypes.data for a in offsets])
        poses = np.empty((len(streams), n, 12), dtype=np.float32) if want_poses else None
        self._check(self.lib.youth_cuda_track_batch_packed(
            self.h, sp, op, n, mem_kind, ts.ctypes.data if ts is not None else None,
            poses.ctypes.data if want_poses else None), "youth_cuda_track_batch_packed")
        return poses

    def sync(self):
        self._check(self.lib.youth_cuda_sync(self.h), "youth_cuda_sync")

    def reset(self, stream=-1):
        self._check(self.lib.youth_cuda_reset(self.h, stream), "youth_cuda_reset")

    def frame_count(self, stream=0):
        return self.lib.youth_cuda_frame_count(self.h, stream)

    def trajectory(self, stream=0):
        n = self.frame_count(stream)
        poses = np.empty((n, 12), dtype=np.float32)
        ts = np.empty(n, dtype=np.uint32)
        st = np.empty(n, dtype=np.uint32)
        got = self.lib.youth_cuda_get_trajectory(self.h, stream, 0, n, poses.ctypes.data, ts.ctypes.data, st.ctypes.data)
        if got < 0:
            raise RuntimeError("youth_cuda_get_trajectory failed: " + self.lib.youth_cuda_last_error().decode())
        return poses[:got], ts[:got], st[:got]

    def read_trajectory_async(self, poses_ptr, n, stream=0, first=0, status_ptr=None):
        """Stream-ordered read-back into (pinned) host memory; returns (frames, ticket)."""
        t = C.c_uint64()
        got = self.lib.youth_cuda_read_trajectory_async(self.h, stream, first, n, poses_ptr, status_ptr, C.byref(t))
        self._check(got >= 0, "youth_cuda_read_trajectory_async")
        return got, t.value

    def wait_ticket(self, ticket):
        self._check(self.lib.youth_cuda_wait_ticket(self.h, ticket), "youth_cuda_wait_ticket")

    def last_inliers(self, stream=0):
        return self.lib.youth_cuda_last_inliers(self.h, stream)

    def enable_debug_maps(self):
        """store the float depth pyramid / pyramid sample counts too (DBG_DEPTH, DBG_PYRCNT); call before tracking"""
        self._check(self.lib.youth_cuda_debug_enable_maps(self.h), "youth_cuda_debug_enable_maps")

    def debug_read(self, what, frame, level, stream=0):
        h, w = self.level_shape(level)
        if what == DBG_DEPTH:
            out = np.empty((h, w), dtype=np.float32)
        elif what in (DBG_VERTEX, DBG_NORMAL):
            out = np.empty((h, w, 4), dtype=np.float32)
        else:
            out = np.empty((h, w), dtype=np.uint8)
        self._check(self.lib.youth_cuda_debug_read(self.h, what, stream, frame, level, out.ctypes.data, out.nbytes),
                    "youth_cuda_debug_read")
        return out

    def debug_icp(self, frame, level, pose, stream=0, want_corr=True):
        h, w = self.level_shape(level)
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        sums = np.empty(SUM_SLOTS, dtype=np.float64)
        corr = np.empty((h, w), dtype=np.int32) if want_corr else None
        self._check(self.lib.youth_cuda_debug_icp(self.h, stream, frame, level, pose.ctypes.data, sums.ctypes.data,
                                                  corr.ctypes.data if want_corr else None), "youth_cuda_debug_icp")
        return sums, corr

    # ---- frame-to-model tracking (include/youth_model.h)
    def enable_model(self, tcfg: TsdfConfig):
        self.tcfg = tcfg
        self._check(self.lib.youth_cuda_enable_model(self.h, C.byref(tcfg)), "youth_cuda_enable_model")

    def surface_voxels(self, stream=0):
        n = self.lib.youth_cuda_model_surface_voxels(self.h, stream)
        if n < 0:
            raise RuntimeError("youth_cuda_model_surface_voxels failed: " + self.lib.youth_cuda_last_error().decode())
        return int(n)

    def read_volume(self, stream=0):
        t = self.tcfg
        vol = np.empty((t.dim[2], t.dim[1], t.dim[0], 2), dtype=np.int16)
        self._check(self.lib.youth_cuda_debug_read_volume(self.h, stream, vol.ctypes.data, vol.nbytes),
                    "youth_cuda_debug_read_volume")
        return vol

    def read_model(self, what, level, stream=0):
        h, w = self.level_shape(level)
        out = np.empty((h, w, 4), dtype=np.float32)
        self._check(self.lib.youth_cuda_debug_read_model(self.h, what, stream, level, out.ctypes.data, out.nbytes),
                    "youth_cuda_debug_read_model")
        return out

    def debug_integrate(self, frame, pose, stream=0):
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        self._check(self.lib.youth_cuda_debug_integrate(self.h, stream, frame, pose.ctypes.data), "youth_cuda_debug_integrate")

    def debug_raycast(self, pose, stream=0, hint_frame=-1):
        pose = np.ascontiguousarray(pose, dtype=np.float32)
        self._check(self.lib.youth_cuda_debug_raycast(self.h, stream, pose.ctypes.data, hint_frame),
                    "youth_cuda_debug_raycast")

    def debug_rcp_check(self, lo_bits, hi_bits):
        n = self.lib.youth_cuda_debug_rcp_check(self.h, lo_bits, hi_bits)
        self._check(n >= 0, "youth_cuda_debug_rcp_check")
        return n

    def timer_start(self):
        self._check(self.lib.youth_cuda_timer_start(self.h), "youth_cuda_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.lib.youth_cuda_timer_stop(self.h, C.byref(ms)), "youth_cuda_timer_stop")
        return ms.value

    def profile(self, on):
        self._check(self.lib.youth_cuda_profile_enable(self.h, 1 if on else 0), "youth_cuda_profile_enable")

    def profile_read(self):
        ms = np.zeros(PROF_CLASSES, dtype=np.float64)
        n = np.zeros(PROF_CLASSES, dtype=np.uint64)
        self._check(self.lib.youth_cuda_profile_read(self.h, ms.ctypes.data, n.ctypes.data), "youth_cuda_profile_read")
        return ms, n

    def launch_count(self):
        return int(self.lib.youth_cuda_launch_count(self.h))


class Codec:
    """Thin object view of a ``youth_codec`` (include/youth_codec.h): the YD16 lossless depth codec."""

    def __init__(self, width, height, max_frames=1, device=0):
        self.lib = cuda_lib()
        self.w, self.h_, self.max_frames = width, height, max_frames
        self.h = C.c_void_p()
        if not self.lib.youth_codec_create(width, height, max_frames, device, C.byref(self.h)):
            raise RuntimeError("youth_codec_create failed: " + self.lib.youth_cuda_last_error().decode())

    def close(self):
        if self.h:
            self.lib.youth_codec_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def max_bytes(self):
        return self.lib.youth_codec_max_bytes(self.w, self.h_)

    def encode_ptr(self, ptr, n, mem_kind, out=None):
        """-> (packed uint8 array holding the n streams back to back, uint64 offsets [n+1])"""
        if out is None:
            out = np.empty(n * self.max_bytes(), dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.uint64)
        if not self.lib.youth_codec_encode(self.h, ptr, mem_kind, n, out.ctypes.data, out.nbytes, offs.ctypes.data):
            raise RuntimeError("youth_codec_encode failed: " + self.lib.youth_cuda_last_error().decode())
        return out[:int(offs[n])], offs

    def encode(self, frames):
        assert frames.dtype == np.uint16 and frames.flags.c_contiguous and frames.shape[1:] == (self.h_, self.w)
        return self.encode_ptr(frames.ctypes.data, frames.shape[0], MEM_HOST)

    def decode(self, packed, offs):
        n = len(offs) - 1
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        out = np.empty((n, self.h_, self.w), dtype=np.uint16)
        if not self.lib.youth_codec_decode(self.h, packed.ctypes.data, offs.ctypes.data, n, out.ctypes.data, MEM_HOST):
            raise RuntimeError("youth_codec_decode failed: " + self.lib.youth_cuda_last_error().decode())
        return out

    def decode_to_device(self, packed, offs, dev_ptr):
        n = len(offs) - 1
        if not self.lib.youth_codec_decode(self.h, packed.ctypes.data, offs.ctypes.data, n, dev_ptr, MEM_DEVICE):
            raise RuntimeError("youth_codec_decode failed: " + self.lib.youth_cuda_last_error().decode())

    def last_kernel_ms(self):
        return float(self.lib.youth_codec_last_kernel_ms(self.h))
