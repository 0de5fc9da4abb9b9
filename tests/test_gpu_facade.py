"""The reference-facing entry points on a GPU: the SLAM.h facade, the algorithmModule()
thread entry replaying a .bin recording, the harness binary, the golden fixture through
the C ABI, and a full-size property run."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def make_tracker(pkg, **kw):
    from slam_rgbd_b200.binding import Tracker

    return Tracker(pkg.default_config(**kw))


def test_facade_push_frames_and_save(pkg, small_seq, tmp_path):
    frames, gt = small_seq
    host = pkg.host_lib()
    host.youthSlamSetOptions(1, 4)  # lossless, 4 frames per launch group
    host.initSlamModule(None, b"ORBvoc.txt")  # vocabulary accepted and ignored
    assert host.isSlamModuleRunning() == 1
    assert host.processSlamFrame(frames[0].ctypes.data, None, 320, 240, 0) == 0  # wrong size is refused
    for i in range(6):
        scratch = frames[i].copy()
        assert host.processSlamFrame(scratch.ctypes.data, None, 640, 480, 33 * i) == 1
        scratch[:] = 0  # the callee copied synchronously (SLAM.cpp:133-134): clobbering is harmless
    host.youthSlamDrain()
    assert host.getSlamMapPoints() > 100000
    poses = np.empty((6, 12), dtype=np.float32)
    ts = np.empty(6, dtype=np.uint32)
    st = np.empty(6, dtype=np.uint32)
    assert host.youthSlamGetTrajectory(poses.ctypes.data, ts.ctypes.data, st.ctypes.data, 6) == 6
    assert list(ts) == [33 * i for i in range(6)] and list(st) == [1, 0, 0, 0, 0, 0]
    ref = make_tracker(pkg, batch=6)
    want = ref.track_batch([frames])[0]
    ref.close()
    assert np.array_equal(poses.view(np.uint32), want.view(np.uint32))  # grouping does not change results
    prefix = str(tmp_path / "map")
    assert host.saveSlamMap(prefix.encode()) == 1
    rows = np.loadtxt(prefix + "_trajectory.txt")
    assert rows.shape == (6, 8) and np.allclose(rows[:, 1:4], poses[:, [3, 7, 11]], atol=1e-6)
    assert np.loadtxt(prefix + "_keyframes.txt", ndmin=2).shape[0] >= 1
    host.resetSlam()
    assert host.getSlamMapPoints() == 0
    assert host.processSlamFrame(frames[0].ctypes.data, None, 640, 480, 0) == 1
    host.youthSlamDrain()
    assert host.youthSlamGetTrajectory(poses.ctypes.data, ts.ctypes.data, st.ctypes.data, 6) == 1 and st[0] == 1
    host.stopSlamModule()
    assert host.isSlamModuleRunning() == 0
    assert host.processSlamFrame(frames[0].ctypes.data, None, 640, 480, 0) == 0


def test_facade_lossy_backpressure(pkg, small_seq):
    """reference policy (SLAM.cpp:163-167): more than 10 queued -> drop oldest down to 5;
    nothing is ever tracked twice and accepted = tracked + dropped."""
    frames, _ = small_seq
    host = pkg.host_lib()
    host.youthSlamSetOptions(0, 1)
    host.initSlamModule(None, None)
    assert host.isSlamModuleRunning() == 1
    for i in range(200):
        assert host.processSlamFrame(frames[i % 6].ctypes.data, None, 640, 480, i) == 1
    host.youthSlamDrain()
    a, d, t = C.c_long(), C.c_long(), C.c_long()
    host.youthSlamStats(C.byref(a), C.byref(d), C.byref(t))
    assert a.value == 200 and a.value == d.value + t.value and t.value >= 1
    host.stopSlamModule()


def test_algorithm_module_replays_bin(pkg, small_seq, tmp_path):
    frames, _ = small_seq
    paths = pkg.lib_paths()
    rec = str(tmp_path / "rec.bin")
    out = subprocess.run([paths["harness"], "gen", rec, "6"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert os.path.getsize(rec) == 6 * 1536028 + 28
    host = pkg.host_lib()
    prefix = str(tmp_path / "run")
    os.environ["YOUTH_SLAM_OUT"] = prefix
    host.youthSlamSetOptions(1, 4)
    th = threading.Thread(target=lambda: host.algorithmModule(C.c_char_p(rec.encode())))
    th.start()
    th.join(timeout=120)
    assert not th.is_alive()
    del os.environ["YOUTH_SLAM_OUT"]
    rows = np.loadtxt(prefix + "_trajectory.txt")
    assert rows.shape == (6, 8)
    ref = make_tracker(pkg, batch=6)
    want = ref.track_batch([frames])[0]  # harness gen writes sequence 0 = the same frames
    ref.close()
    assert np.allclose(rows[:, 1:4], want[:, [3, 7, 11]], atol=1e-6)
    gt_rows = np.loadtxt(rec + ".gt.txt")
    assert np.abs(rows[:, 1:4] - gt_rows[:, 1:4]).max() < 2e-3


def test_packed_recording_replays_to_the_same_trajectory(pkg, oracle, tmp_path):
    """raw .bin -> `youth_harness pack` (GPU YD16 encoder) -> replay through algorithmModule(): the packed
    records are unpacked on the device and give the trajectory of the raw recording, text-identical."""
    paths = pkg.lib_paths()
    rec, pk = str(tmp_path / "rec.bin"), str(tmp_path / "rec_packed.bin")
    assert subprocess.run([paths["harness"], "gen", rec, "70"], capture_output=True).returncode == 0
    out = subprocess.run([paths["harness"], "pack", rec, pk], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert os.path.getsize(pk) < os.path.getsize(rec) - 70 * 350000  # depth shrank by more than half
    # the packed file is a valid record stream: same headers except type/size, payload decodes to the raw depth
    host = pkg.host_lib()
    io = C.CDLL(paths["host"])  # private handle: argtypes set here do not leak into other tests
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    io.youth_bin_read_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    fa, fb = libc.fopen(rec.encode(), b"rb"), libc.fopen(pk.encode(), b"rb")
    ha, hb = (C.c_uint8 * 28)(), (C.c_uint8 * 28)()
    da, db = np.empty(640 * 480, dtype=np.uint16), np.empty(1 << 20, dtype=np.uint8)
    for i in range(3):
        assert io.youth_bin_read_frame(fa, ha, da.ctypes.data, da.nbytes, None, 0) == 1
        assert io.youth_bin_read_frame(fb, hb, db.ctypes.data, db.nbytes, None, 0) == 1
        a, b = np.frombuffer(bytes(ha), dtype=np.uint8), np.frombuffer(bytes(hb), dtype=np.uint8)
        assert np.array_equal(a[:8], b[:8]) and np.array_equal(a[10:14], b[10:14]) and b[8] == 2  # FRAME_TYPE_DEPTH_PACKED
        nbytes = int(np.frombuffer(bytes(hb), dtype=np.uint32)[4])
        assert np.array_equal(oracle.codec_decode(db[:nbytes], 640, 480).ravel(), da)
    libc.fclose(fa)
    libc.fclose(fb)
    rows = []
    for path, tag in ((rec, "raw"), (pk, "packed")):
        prefix = str(tmp_path / tag)
        os.environ["YOUTH_SLAM_OUT"] = prefix
        host.youthSlamSetOptions(1, 32)
        th = threading.Thread(target=lambda: host.algorithmModule(C.c_char_p(path.encode())))
        th.start()
        th.join(timeout=120)
        assert not th.is_alive()
        del os.environ["YOUTH_SLAM_OUT"]
        rows.append(open(prefix + "_trajectory.txt").read())
    assert len(rows[0].splitlines()) == 70 and rows[0] == rows[1]


def test_facade_in_model_mode(pkg, small_seq):
    """YOUTH_SLAM_MODE=model: the same facade calls, tracking against the fused TSDF model."""
    from slam_rgbd_b200 import binding as B

    frames, gt = small_seq
    host = pkg.host_lib()
    os.environ["YOUTH_SLAM_MODE"] = "model"
    try:
        host.youthSlamSetOptions(1, 4)
        host.initSlamModule(None, None)
        assert host.isSlamModuleRunning() == 1
        for i in range(6):
            assert host.processSlamFrame(frames[i].ctypes.data, None, 640, 480, 33 * i) == 1
        host.youthSlamDrain()
        poses = np.empty((6, 12), dtype=np.float32)
        assert host.youthSlamGetTrajectory(poses.ctypes.data, None, None, 6) == 6
        assert host.getSlamMapPoints() > 20000  # the size of the map: surface voxels of the fused volume
        host.stopSlamModule()
    finally:
        del os.environ["YOUTH_SLAM_MODE"]
    ref = B.Tracker(pkg.default_config(batch=6))
    ref.enable_model(pkg.tsdf_config())
    want = ref.track_batch([frames])[0]
    ref.close()
    assert np.array_equal(poses.view(np.uint32), want.view(np.uint32))
    f2f = make_tracker(pkg, batch=6)
    other = f2f.track_batch([frames])[0]
    f2f.close()
    assert not np.array_equal(poses, other)  # it really is the other tracker
    assert np.abs(poses[:, [3, 7, 11]] - gt[:, [3, 7, 11]]).max() < 2e-3


def test_live_path_reference_chunks_to_tracker(pkg, small_seq):
    """(f)1 live path: frames leave through the REFERENCE's own sendMetadata/sendDataInChunks
    (oracle/_ref, compiled from loggingModule.c) over a real POSIX mq, are reassembled by
    youth_reasm_feed (the hook INTEGRATION.md section 3 describes) and pushed into
    processSlamFrame at the reference's frame-complete point; the trajectory must equal
    tracking the same frames directly."""
    ref_so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_logging.so")
    if not os.path.exists(ref_so):
        pytest.skip("oracle/_ref/libref_logging.so not built")
    ref = C.CDLL(ref_so)
    ref.ref_chunk_stream.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
    frames, _ = small_seq
    host = pkg.host_lib()
    host.youthSlamSetOptions(1, 2)
    host.initSlamModule(None, None)
    assert host.isSlamModuleRunning() == 1
    reasm = host.youth_reasm_create()
    color = np.full((480, 640, 3), 128, dtype=np.uint8)
    msgs = np.zeros((200, 8192), dtype=np.uint8)
    lens = np.zeros(200, dtype=np.int32)
    for i in range(4):
        n = ref.ref_chunk_stream(b"/youth_live_mq_%d" % os.getpid(), i, 33 * i, 640, 480, frames[i].ctypes.data,
                                 color.ctypes.data, msgs.ctypes.data, lens.ctypes.data, 200)
        if n < 0:
            host.stopSlamModule()
            pytest.skip("POSIX message queues unavailable")
        assert n == 196
        completed = 0
        for k in range(n):
            if host.youth_reasm_feed(reasm, msgs[k].ctypes.data, int(lens[k])) == 1:  # loggingModule.c:354
                completed += 1
                w, h, fid, ts = C.c_int(), C.c_int(), C.c_int(), C.c_uint32()
                host.youth_reasm_info(reasm, C.byref(w), C.byref(h), C.byref(fid), C.byref(ts))
                assert (w.value, h.value, fid.value, ts.value) == (640, 480, i, 33 * i)
                assert host.processSlamFrame(host.youth_reasm_depth(reasm), host.youth_reasm_color(reasm), w.value,
                                             h.value, ts.value) == 1
        assert completed == 1
    host.youthSlamDrain()
    poses = np.empty((4, 12), dtype=np.float32)
    assert host.youthSlamGetTrajectory(poses.ctypes.data, None, None, 4) == 4
    host.youth_reasm_destroy(reasm)
    host.stopSlamModule()
    direct = make_tracker(pkg, batch=4)
    want = direct.track_batch([frames[:4]])[0]
    direct.close()
    assert np.array_equal(poses.view(np.uint32), want.view(np.uint32))


def test_pose_egress_to_viewer_queue(pkg, small_seq):
    """(f)2: with YOUTH_SLAM_POSE_MQ set the facade publishes one MSG_TYPE_POSE message per tracked
    frame on that POSIX queue; the messages carry the trajectory."""
    class MqAttr(C.Structure):
        _fields_ = [("mq_flags", C.c_long), ("mq_maxmsg", C.c_long), ("mq_msgsize", C.c_long), ("mq_curmsgs", C.c_long),
                    ("pad", C.c_long * 4)]

    rt = C.CDLL(None, use_errno=True)
    rt.mq_open.restype = C.c_int
    rt.mq_open.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.POINTER(MqAttr)]
    rt.mq_receive.restype = C.c_ssize_t
    rt.mq_receive.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    name = b"/youth_pose_mq_%d" % os.getpid()
    attr = MqAttr(0, 10, 8192, 0)
    rt.mq_unlink(name)
    mq = rt.mq_open(name, os.O_RDONLY | os.O_CREAT | os.O_NONBLOCK, 0o644, C.byref(attr))
    if mq < 0:
        pytest.skip("POSIX message queues unavailable")
    frames, _ = small_seq
    host = pkg.host_lib()
    os.environ["YOUTH_SLAM_POSE_MQ"] = name.decode()
    try:
        host.youthSlamSetOptions(1, 3)
        host.initSlamModule(None, None)
        assert host.isSlamModuleRunning() == 1
        for i in range(6):
            assert host.processSlamFrame(frames[i].ctypes.data, None, 640, 480, 33 * i) == 1
        host.youthSlamDrain()
        poses = np.empty((6, 12), dtype=np.float32)
        assert host.youthSlamGetTrajectory(poses.ctypes.data, None, None, 6) == 6
        host.stopSlamModule()
    finally:
        del os.environ["YOUTH_SLAM_POSE_MQ"]
    buf = np.zeros(8192, dtype=np.uint8)
    got = []
    while True:
        n = rt.mq_receive(mq, buf.ctypes.data, 8192, None)
        if n < 0:
            break
        out = np.zeros(14, dtype=np.uint32)
        fid, ts = C.c_int(), C.c_uint32()
        assert host.youth_pose_msg_parse(buf.ctypes.data, n, C.byref(fid), C.byref(ts), out.ctypes.data) == 1
        got.append((fid.value, ts.value, out[:12].view(np.float32).copy(), int(out[12]), int(out[13])))
    rt.mq_close(mq)
    rt.mq_unlink(name)
    assert [g[0] for g in got] == list(range(6)) and [g[1] for g in got] == [33 * i for i in range(6)]
    assert all(np.array_equal(g[2], poses[i]) for i, g in enumerate(got))
    # status = the frame's own YOUTH_STATUS_* word (the first frame of a sequence is FIRST, tracked frames 0),
    # inliers = the inlier count of the launch group's last frame (> 0 for tracked groups)
    assert [g[3] for g in got] == [1, 0, 0, 0, 0, 0]
    assert all(g[4] > 100000 for g in got)


def test_golden_fixture_through_cabi(pkg):
    from slam_rgbd_b200 import binding as B

    g = np.load(os.path.join(HERE, "golden", "golden_160x120.npz"))
    frames = g["frames"]
    W, H = 160, 120
    trk = make_tracker(pkg, width=W, height=H, fx=570.3 * W / 640, fy=570.3 * W / 640, cx=W / 2.0, cy=H / 2.0, batch=4)
    poses = trk.track_batch([frames])[0]
    assert np.array_equal(poses.view(np.uint32), g["poses"].view(np.uint32))
    _, _, st = trk.trajectory()
    assert np.array_equal(st, g["status"])
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    for level in range(3):
        sums, corr = trk.debug_icp(1, level, ident)
        assert np.array_equal(corr, g[f"corr_l{level}"])
        assert np.array_equal(sums.view(np.uint64), g[f"sums_l{level}"].view(np.uint64))
        assert np.array_equal(trk.debug_read(B.DBG_MASK, 1, level), g[f"mask_l{level}"])
    trk.close()


def test_four_levels_high_res_parity(pkg, oracle):
    """configs[4] shape: 1280x960, 4-level pyramid (iterations 10/5/4/4), two frames."""
    W, H = 1280, 960
    frames = pkg.synth_sequence(2, W, H, sequence=1)
    kw = dict(width=W, height=H, fx=1140.6, fy=1140.6, cx=640.0, cy=480.0, levels=4, batch=2)
    trk = make_tracker(pkg, **kw)
    poses = trk.track_batch([frames])[0]
    want, st, _ = oracle.track_sequence(oracle.config_from(trk.cfg), frames)
    assert np.array_equal(poses.view(np.uint32), want.view(np.uint32))
    assert list(st) == [1, 0]
    trk.close()


def test_full_size_properties(pkg):
    """BASELINE-size run (300 frames, 640x480): size-independent properties -- batched ==
    differently batched (bitwise), rotations stay orthonormal, nothing is flagged lost, the
    closed synthetic loop returns near its start."""
    frames = pkg.synth_sequence(300)
    a = make_tracker(pkg, batch=32, traj_capacity=300)
    for s in range(0, 300, 32):
        a.track_batch([frames[s:s + 32]], want_poses=False)
    pa, _, sa = a.trajectory()
    b = make_tracker(pkg, batch=7, traj_capacity=300)
    for s in range(0, 300, 7):
        b.track_batch([frames[s:s + 7]], want_poses=False)
    pb, _, sb = b.trajectory()
    assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32)) and np.array_equal(sa, sb)
    assert sa[0] == 1 and not (sa[1:] & 2).any()
    R = pa.reshape(300, 3, 4)[:, :, :3].astype(np.float64)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-5
    gt = pkg.synth_gt(300)
    err = np.linalg.norm(pa.reshape(300, 3, 4)[:, :, 3] - gt.reshape(300, 3, 4)[:, :, 3], axis=1)
    assert err.max() < 0.05  # frame-to-frame drift over 300 frames stays below 5 cm
    a.close()
    b.close()
