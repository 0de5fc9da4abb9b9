"""ctypes view of the CPU oracle (oracle/youth_oracle.h).  TEST INFRASTRUCTURE ONLY: may be
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by anything under slam-rgbd_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 4
SUM_SLOTS = 32


class OracleConfig(C.Structure):
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
    ]


class TsdfConfig(C.Structure):
    """Mirror of yo_tsdf_config (oracle/youth_tsdf_oracle.c) == youth_tsdf_config (include/youth_model.h)."""

    _fields_ = [("dim", C.c_int32 * 3), ("voxel_m", C.c_float), ("origin", C.c_float * 3), ("trunc_m", C.c_float),
                ("max_weight", C.c_int32), ("near_m", C.c_float), ("far_m", C.c_float)]


class Level(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float)]


class Frame(C.Structure):
    _fields_ = [("depth", C.POINTER(C.c_float) * MAX_LEVELS), ("pyrcnt", C.POINTER(C.c_uint8) * MAX_LEVELS),
                ("vmap", C.POINTER(C.c_float) * MAX_LEVELS), ("nmap", C.POINTER(C.c_float) * MAX_LEVELS)]


def build(force=False):
    so = os.path.join(HERE, "_build", "libyouth_oracle_hwfma.so")
    srcs = [os.path.join(HERE, f) for f in ("youth_oracle.c", "youth_codec_oracle.c", "youth_tsdf_oracle.c", "youth_oracle.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)
    return so


_libs = {}


def fast_supported():
    """The speed build uses AVX2+FMA; only load it on hosts that have both."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return "avx2" in flags and "fma" in flags
    except OSError:
        pass
    return False


def lib(fast=False):
    if fast and not fast_supported():
        raise RuntimeError("host CPU lacks AVX2/FMA: speed build of the oracle not usable")
    # parity build: with hardware FMA when the host has it (fmaf() inlined), else the portable
    # build that calls libm's fmaf() -- both are single-rounding, results are identical
    key = "fast" if fast else ("hwfma" if fast_supported() else "parity")
    if key in _libs:
        return _libs[key]
    build()
    name = {"fast": "libyouth_oracle_fast.so", "hwfma": "libyouth_oracle_hwfma.so", "parity": "libyouth_oracle.so"}[key]
    path = os.path.join(HERE, "_build", name)
    L = C.CDLL(path)
    CP = C.POINTER(OracleConfig)
    FP = C.POINTER(Frame)
    sig = {
        "yo_default_config": (None, [CP]),
        "yo_level_geometry": (C.c_int, [CP, C.c_int, C.POINTER(Level)]),
        "yo_frame_alloc": (FP, [CP]),
        "yo_frame_free": (None, [FP]),
        "yo_bilateral": (None, [CP, C.c_void_p, C.c_void_p]),
        "yo_pyrdown": (None, [CP, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
        "yo_vertex_normal": (None, [CP, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
        "yo_preprocess": (None, [CP, C.c_void_p, FP]),
        "yo_icp_sums": (None, [CP, C.c_int, FP, FP, C.c_void_p, C.c_void_p, C.c_void_p]),
        "yo_solve_update": (C.c_int, [CP, C.c_void_p, C.c_void_p, C.c_void_p]),
        "yo_track_pair": (C.c_uint32, [CP, FP, FP, C.c_void_p, C.POINTER(C.c_int32)]),
        "yo_compose": (None, [C.c_void_p, C.c_void_p, C.c_void_p]),
        "yo_track_sequence": (C.c_double, [CP, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    }
    if not fast:  # the codec statement lives in the parity builds only
        sig.update({
            "yc_max_bytes": (C.c_size_t, [C.c_int, C.c_int]),
            "yc_encode": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
            "yc_decode": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]),
            # frame-to-model statement (youth_tsdf_oracle.c)
            "yo_tsdf_default_config": (None, [C.POINTER(TsdfConfig)]),
            "yo_tsdf_clear": (None, [C.POINTER(TsdfConfig), C.c_void_p]),
            "yo_tsdf_integrate": (None, [CP, C.POINTER(TsdfConfig), C.c_void_p, C.c_void_p, C.c_void_p]),
            "yo_tsdf_raycast": (None, [CP, C.POINTER(TsdfConfig), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
            "yo_tsdf_surface_voxels": (C.c_int64, [C.POINTER(TsdfConfig), C.c_void_p]),
            "yo_track_sequence_model": (None, [CP, C.POINTER(TsdfConfig), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
        })
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _libs[key] = L
    return L


def default_config(**overrides) -> OracleConfig:
    cfg = OracleConfig()
    lib().yo_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if k == "iters":
            for i, it in enumerate(v):
                cfg.iters[i] = int(it)
        else:
            setattr(cfg, k, v)
    return cfg


def config_from(product_cfg) -> OracleConfig:
    """Copy the fields the oracle shares with youth_cuda_config (same names)."""
    cfg = OracleConfig()
    for name, _ in OracleConfig._fields_:
        if name == "iters":
            for i in range(MAX_LEVELS):
                cfg.iters[i] = product_cfg.iters[i]
        else:
            setattr(cfg, name, getattr(product_cfg, name))
    return cfg


def level_geometry(cfg, level) -> Level:
    g = Level()
    assert lib().yo_level_geometry(C.byref(cfg), level, C.byref(g))
    return g


class OFrame:
    """One preprocessed frame held by the oracle; numpy views into its buffers."""

    def __init__(self, cfg, raw=None, fast=False):
        self.cfg = cfg
        self.L = lib(fast)
        self.ptr = self.L.yo_frame_alloc(C.byref(cfg))
        if raw is not None:
            self.preprocess(raw)

    def preprocess(self, raw):
        assert raw.dtype == np.uint16 and raw.flags.c_contiguous
        self.L.yo_preprocess(C.byref(self.cfg), raw.ctypes.data, self.ptr)

    def _view(self, field, level, ctype, comps):
        h, w = self.cfg.height >> level, self.cfg.width >> level
        p = getattr(self.ptr.contents, field)[level]
        buf = (ctype * (h * w * comps)).from_address(C.addressof(p.contents))
        buf._owner = self  # the view keeps this frame (and so the C buffers) alive
        arr = np.ctypeslib.as_array(buf)
        return arr.reshape((h, w, comps) if comps > 1 else (h, w))

    def depth(self, level):
        return self._view("depth", level, C.c_float, 1)

    def pyrcnt(self, level):
        return self._view("pyrcnt", level, C.c_uint8, 1)

    def vmap(self, level):
        return self._view("vmap", level, C.c_float, 4)

    def nmap(self, level):
        return self._view("nmap", level, C.c_float, 4)

    def mask(self, level):
        return ((self.vmap(level)[..., 3] != 0).astype(np.uint8) | ((self.nmap(level)[..., 3] != 0).astype(np.uint8) << 1))

    def __del__(self):
        try:
            self.L.yo_frame_free(self.ptr)
        except Exception:
            pass


def icp_sums(cfg, level, cur: OFrame, prev: OFrame, pose, want_corr=True):
    pose = np.ascontiguousarray(pose, dtype=np.float32)
    sums = np.zeros(SUM_SLOTS, dtype=np.float64)
    h, w = cfg.height >> level, cfg.width >> level
    corr = np.empty((h, w), dtype=np.int32) if want_corr else None
    lib().yo_icp_sums(C.byref(cfg), level, cur.ptr, prev.ptr, pose.ctypes.data, sums.ctypes.data,
                      corr.ctypes.data if want_corr else None)
    return sums, corr


def track_pair(cfg, cur: OFrame, prev: OFrame):
    rel = np.empty(12, dtype=np.float64)
    inl = C.c_int32()
    st = lib().yo_track_pair(C.byref(cfg), cur.ptr, prev.ptr, rel.ctypes.data, C.byref(inl))
    return rel, int(st), inl.value


def track_sequence(cfg, frames, fast=False):
    """frames uint16 [n][H][W] -> (poses float32 [n][12], status uint32 [n], seconds)."""
    assert frames.dtype == np.uint16 and frames.flags.c_contiguous
    n = frames.shape[0]
    poses = np.empty((n, 12), dtype=np.float32)
    status = np.empty(n, dtype=np.uint32)
    secs = lib(fast).yo_track_sequence(C.byref(cfg), frames.ctypes.data, n, poses.ctypes.data, status.ctypes.data)
    return poses, status, secs


def track_sequence_parallel(cfg, frames, threads=None, fast=False):
    """yo_track_sequence with the independent frame pairs spread over host threads: every frame is preprocessed
    once (yo_preprocess), every pair tracked from the identity (yo_track_pair), and the pose chain composed in
    frame order (yo_compose) -- the same calls in the same per-pair order as yo_tracker_track, so the result is
    bit-identical to track_sequence (tests/test_oracle.py checks it).  Returns (poses f32 [n][12], status u32 [n],
    inliers i32 [n], seconds)."""
    import concurrent.futures
    import time

    assert frames.dtype == np.uint16 and frames.flags.c_contiguous
    n = frames.shape[0]
    L = lib(fast)
    threads = max(1, min(threads or (os.cpu_count() or 1), n))
    t0 = time.perf_counter()
    ofr = [None] * n

    def prep(i):
        ofr[i] = OFrame(cfg, frames[i], fast=fast)

    rel = np.zeros((n, 12), dtype=np.float64)
    st = np.zeros(n, dtype=np.uint32)
    inl = np.zeros(n, dtype=np.int32)

    def pair(i):
        v = C.c_int32()
        st[i] = L.yo_track_pair(C.byref(cfg), ofr[i].ptr, ofr[i - 1].ptr, rel[i].ctypes.data, C.byref(v))
        inl[i] = v.value

    with concurrent.futures.ThreadPoolExecutor(threads) as ex:
        list(ex.map(prep, range(n)))
        list(ex.map(pair, range(1, n)))
    world = np.zeros(12, dtype=np.float64)
    world[[0, 5, 10]] = 1.0
    poses = np.empty((n, 12), dtype=np.float32)
    st[0] = 1  # YO_STATUS_FIRST
    for i in range(n):
        if i:
            L.yo_compose(world.ctypes.data, rel[i].ctypes.data, world.ctypes.data)
        poses[i] = world.astype(np.float32)
    return poses, st, inl, time.perf_counter() - t0


def codec_encode(frame):
    """uint16 [H][W] -> packed bytes (numpy uint8) with the CPU statement of the YD16 codec."""
    assert frame.dtype == np.uint16 and frame.flags.c_contiguous
    h, w = frame.shape
    out = np.zeros(lib().yc_max_bytes(w, h), dtype=np.uint8)
    n = lib().yc_encode(frame.ctypes.data, w, h, out.ctypes.data)
    return out[:n].copy()


def codec_decode(stream, w, h):
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    out = np.empty((h, w), dtype=np.uint16)
    ok = lib().yc_decode(stream.ctypes.data, stream.size, w, h, out.ctypes.data)
    return out if ok else None


# --------------------------------------------------------------------------- frame-to-model


def tsdf_config(**overrides) -> TsdfConfig:
    t = TsdfConfig()
    lib().yo_tsdf_default_config(C.byref(t))
    for k, v in overrides.items():
        if k in ("dim", "origin"):
            for i, x in enumerate(v):
                getattr(t, k)[i] = x
        else:
            setattr(t, k, v)
    return t


def tsdf_config_from(product_tcfg) -> TsdfConfig:
    t = TsdfConfig()
    for name, _ in TsdfConfig._fields_:
        if name in ("dim", "origin"):
            for i in range(3):
                getattr(t, name)[i] = getattr(product_tcfg, name)[i]
        else:
            setattr(t, name, getattr(product_tcfg, name))
    return t


def tsdf_new(tcfg):
    """fresh volume: int16 [dz][dy][dx][2] = (32767, 0)"""
    vol = np.empty((tcfg.dim[2], tcfg.dim[1], tcfg.dim[0], 2), dtype=np.int16)
    lib().yo_tsdf_clear(C.byref(tcfg), vol.ctypes.data)
    return vol


def tsdf_integrate(cfg, tcfg, vol, depth0, pose):
    depth0 = np.ascontiguousarray(depth0, dtype=np.float32)
    pose = np.ascontiguousarray(pose, dtype=np.float32)
    lib().yo_tsdf_integrate(C.byref(cfg), C.byref(tcfg), vol.ctypes.data, depth0.ctypes.data, pose.ctypes.data)


def tsdf_raycast(cfg, tcfg, vol, pose, level, hint=None):
    """hint: optional float32 depth map of this level (raw units) seen from `pose` (march start, see the C file)"""
    h, w = cfg.height >> level, cfg.width >> level
    pose = np.ascontiguousarray(pose, dtype=np.float32)
    vmap = np.empty((h, w, 4), dtype=np.float32)
    nmap = np.empty((h, w, 4), dtype=np.float32)
    if hint is not None:
        hint = np.ascontiguousarray(hint, dtype=np.float32)
        assert hint.shape == (h, w)
    lib().yo_tsdf_raycast(C.byref(cfg), C.byref(tcfg), vol.ctypes.data, pose.ctypes.data, level,
                          hint.ctypes.data if hint is not None else None, vmap.ctypes.data, nmap.ctypes.data)
    return vmap, nmap


def track_sequence_model(cfg, tcfg, frames):
    assert frames.dtype == np.uint16 and frames.flags.c_contiguous
    n = frames.shape[0]
    poses = np.empty((n, 12), dtype=np.float32)
    status = np.empty(n, dtype=np.uint32)
    lib().yo_track_sequence_model(C.byref(cfg), C.byref(tcfg), frames.ctypes.data, n, poses.ctypes.data, status.ctypes.data)
    return poses, status


def tsdf_surface_voxels(tcfg, vol):
    return int(lib().yo_tsdf_surface_voxels(C.byref(tcfg), vol.ctypes.data))
