"""Host logic of the AlgorithmModule facade (slam-rgbd_b200/host/slam_facade.c, algorithm_module.c) on a machine
without a GPU: the facade sources are linked with a TEST DOUBLE of the inner C ABI (tests/stub/youth_cuda_stub.c:
no tracking arithmetic, a "pose" is a marker of the frame it was handed) into tests/_build/libfacade_under_test.so.
What is checked is what the reference's SLAM.cpp does around its tracker: synchronous copy-in, the bounded queue
with drop-oldest back-pressure (SLAM.cpp:158-169), the lossless mode of this build, groups in flight, drain / stop /
reset, TUM egress (SLAM.cpp:187-188), the replay and queue-consumer modes of algorithmModule().  Parity of the
tracker itself is the business of the -m gpu tests; the product library never contains the stub."""
import ctypes as C
import os
import subprocess
import threading
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 64, 48
HOST_SRC = ["slam_facade.c", "algorithm_module.c", "youth_frameio.c", "youth_config.c", "youth_synth.c"]


@pytest.fixture(scope="module")
def fut(tmp_path_factory):
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libfacade_under_test.so")
    src = [os.path.join(ROOT, "tests", "stub", "youth_cuda_stub.c")] + \
          [os.path.join(ROOT, "slam-rgbd_b200", "host", f) for f in HOST_SRC]
    cmd = ["gcc", "-O2", "-std=gnu11", "-fPIC", "-Wall", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
           "-shared", "-Wl,-Bsymbolic", "-o", so] + src + ["-lpthread", "-lrt", "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    L = C.CDLL(so)
    L.initSlamModule.argtypes = [C.c_char_p, C.c_char_p]
    L.processSlamFrame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32]
    L.saveSlamMap.argtypes = [C.c_char_p]
    L.youthSlamSetOptions.argtypes = [C.c_int, C.c_int]
    L.youthSlamGetTrajectory.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.youthSlamStats.argtypes = [C.POINTER(C.c_long)] * 3
    L.algorithmModule.argtypes = [C.c_char_p]
    L.algorithmModule.restype = C.c_void_p
    L.youth_bin_write_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.youth_bin_write_eof.argtypes = [C.c_void_p]
    L.stub_torn_frames.restype = C.c_long
    L.youthSlamAcquireSlot.restype = C.c_void_p
    L.youthSlamAcquireSlot.argtypes = [C.c_int, C.c_int]
    L.youthSlamCommitSlot.argtypes = [C.c_uint32]
    L.youthSlamProcessPinnedFrames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    cfg = tmp_path_factory.mktemp("cfg") / "cam.yaml"
    cfg.write_text(f"%YAML:1.0\nCamera.width: {W}\nCamera.height: {H}\nCamera.fx: 57.03\nCamera.fy: 57.03\n"
                   "Camera.cx: 32.0\nCamera.cy: 24.0\nDepthMapFactor: 1000.0\n")
    L.cfg_path = str(cfg).encode()
    yield L
    L.stopSlamModule()


def frame(marker):
    f = np.full((H, W), 1500, dtype=np.uint16)
    f[0, 0] = f[-1, -1] = marker  # the test double checks that both stay put while it holds the frame
    return f


def stats(L):
    a, d, t = C.c_long(), C.c_long(), C.c_long()
    L.youthSlamStats(C.byref(a), C.byref(d), C.byref(t))
    return a.value, d.value, t.value


def trajectory(L, n):
    poses = np.zeros((n, 12), dtype=np.float32)
    ts = np.zeros(n, dtype=np.uint32)
    got = L.youthSlamGetTrajectory(poses.ctypes.data, ts.ctypes.data, None, n)
    return poses[:got], ts[:got]


def test_the_product_libraries_do_not_contain_the_test_double(pkg):
    for key in ("cuda", "host"):
        lib = C.CDLL(pkg.lib_paths()[key])
        assert not hasattr(lib, "stub_track_calls")
    needed = subprocess.run(["readelf", "-d", pkg.lib_paths()["host"]], capture_output=True, text=True).stdout
    assert "libyouth_cuda.so" in needed  # the facade's tracker is the CUDA library, by link


def test_entry_points_refuse_work_when_not_running(fut):
    assert fut.isSlamModuleRunning() == 0
    assert fut.processSlamFrame(frame(1).ctypes.data, None, W, H, 0) == 0  # SLAM.h:21: 0 = failure
    assert fut.saveSlamMap(b"/tmp/youth_should_not_exist") == 0
    assert fut.getSlamMapPoints() == 0
    fut.resetSlam()
    fut.stopSlamModule()  # no-ops, like SLAM.cpp:97-99
    fut.initSlamModule(b"/nonexistent/config.yaml", None)  # SLAM.cpp:91-94: a failed init leaves the module stopped
    assert fut.isSlamModuleRunning() == 0


def test_lossless_mode_tracks_every_frame_once_and_in_order(fut, tmp_path):
    fut.youthSlamSetOptions(1, 4)
    fut.initSlamModule(fut.cfg_path, b"ignored_vocabulary.txt")
    assert fut.isSlamModuleRunning() == 1
    fut.initSlamModule(fut.cfg_path, None)  # second init: refused, the running module is untouched
    assert fut.isSlamModuleRunning() == 1
    n = 150
    buf = frame(0)
    for i in range(1, n + 1):
        buf[0, 0] = buf[-1, -1] = i  # the caller's buffer is reused at once: the callee copies before returning (SLAM.cpp:133-134)
        assert fut.processSlamFrame(buf.ctypes.data, None, W, H, 33 * i) == 1
    assert fut.processSlamFrame(buf.ctypes.data, None, W + 8, H, 0) == 0  # not the configured size
    assert fut.processSlamFrame(None, None, W, H, 0) == 0
    fut.youthSlamDrain()
    assert stats(fut) == (n, 0, n)
    poses, ts = trajectory(fut, n + 10)
    assert len(poses) == n
    assert list(poses[:, 3]) == list(range(1, n + 1)) and list(ts) == [33 * i for i in range(1, n + 1)]
    assert poses[:, 11].max() <= 4 and np.all(poses[:, 7] < poses[:, 11])  # groups of at most `batch` frames
    assert fut.getSlamMapPoints() == 1000 + n  # the stub's inlier count of the last group
    prefix = str(tmp_path / "map")
    assert fut.saveSlamMap(prefix.encode()) == 1  # SLAM.cpp:187-188: TUM trajectory + key frames
    rows = np.loadtxt(prefix + "_trajectory.txt")
    assert rows.shape == (n, 8) and np.allclose(rows[:, 0], ts / 1000.0) and np.allclose(rows[:, 1], poses[:, 3])
    assert np.allclose(rows[:, 4:], [0, 0, 0, 1])
    keys = np.loadtxt(prefix + "_keyframes.txt").reshape(-1, 8)
    assert len(keys) == n  # every marker frame "moved" 1 m: all of them are key frames
    fut.resetSlam()  # SLAM.cpp:220-228
    assert len(trajectory(fut, 10)[0]) == 0
    assert fut.processSlamFrame(frame(7).ctypes.data, None, W, H, 5) == 1
    fut.youthSlamDrain()
    assert list(trajectory(fut, 10)[0][:, 3]) == [7.0]
    fut.stopSlamModule()
    assert fut.isSlamModuleRunning() == 0


def test_zero_copy_producer_fills_ring_slots_in_place(fut, monkeypatch):
    """youthSlamAcquireSlot / youthSlamCommitSlot (include/youth_slam_ext.h): the producer writes the frame into the
    page-locked ring itself; mixed freely with processSlamFrame(), every frame is tracked once and in order; large
    launch groups (the 64-frame cap is gone)."""
    monkeypatch.setenv("YOUTH_STUB_DELAY_US", "300")  # a tracker slower than the producer: frames queue up into groups
    fut.youthSlamSetOptions(1, 96)
    fut.initSlamModule(fut.cfg_path, None)
    assert fut.youthSlamCommitSlot(0) == 0  # nothing acquired
    assert fut.youthSlamAcquireSlot(W, H)
    fut.youthSlamAbortSlot()  # given back unpublished: nothing is tracked, the next acquire gets the same slot
    assert fut.youthSlamCommitSlot(0) == 0
    assert not fut.youthSlamAcquireSlot(W + 8, H)  # not the configured size
    n = 400
    for i in range(1, n + 1):
        if i % 3 == 0:
            f = frame(i)
            assert fut.processSlamFrame(f.ctypes.data, None, W, H, 7 * i) == 1
        else:
            slot = fut.youthSlamAcquireSlot(W, H)
            assert slot
            view = np.ctypeslib.as_array((C.c_uint16 * (W * H)).from_address(slot)).reshape(H, W)
            view[...] = frame(i)
            assert fut.youthSlamCommitSlot(7 * i) == 1
    fut.youthSlamDrain()
    assert stats(fut) == (n, 0, n)
    poses, ts = trajectory(fut, n)
    assert list(poses[:, 3]) == list(range(1, n + 1)) and list(ts) == [7 * i for i in range(1, n + 1)]
    assert poses[:, 11].max() > 64  # launch groups larger than the old cap did form
    assert fut.stub_torn_frames() == 0
    fut.stopSlamModule()
    assert not fut.youthSlamAcquireSlot(W, H)  # stopped


def test_parallel_copy_in_and_pinned_runs(fut, monkeypatch):
    """processSlamFrame with the copy shared by helper threads (forced for the small test frames), and
    youthSlamProcessPinnedFrames (no copy: the tracker reads the caller's page-locked frames in place), interleaved:
    every frame once, in order, whole (first and last pixel carry the marker)."""
    monkeypatch.setenv("YOUTH_SLAM_COPY_MIN_BYTES", "256")
    monkeypatch.setenv("YOUTH_SLAM_COPY_THREADS", "4")
    fut.youthSlamSetOptions(1, 16)
    fut.initSlamModule(fut.cfg_path, None)
    marks, i = [], 1
    for rnd in range(6):
        for _ in range(25):
            f = frame(i)
            assert fut.processSlamFrame(f.ctypes.data, None, W, H, i) == 1
            marks.append(i)
            i += 1
        run = np.stack([frame(i + k) for k in range(40)])
        ts = np.arange(i, i + 40, dtype=np.uint32)
        assert fut.youthSlamProcessPinnedFrames(run.ctypes.data, 40, W, H, ts.ctypes.data) == 1
        marks += list(range(i, i + 40))
        i += 40
    fut.youthSlamDrain()
    n = len(marks)
    assert stats(fut) == (n, 0, n)
    poses, ts = trajectory(fut, n)
    assert list(poses[:, 3]) == marks and list(ts) == marks
    assert fut.stub_torn_frames() == 0
    assert fut.youthSlamProcessPinnedFrames(run.ctypes.data, 40, W + 8, H, None) == 0
    fut.stopSlamModule()
    assert fut.youthSlamProcessPinnedFrames(run.ctypes.data, 40, W, H, None) == 0


def test_tracker_failure_is_sticky_until_reset(fut, monkeypatch):
    """once the tracker refuses a run (here: trajectory capacity exhausted) processSlamFrame reports failure
    instead of accepting frames that can never be tracked; resetSlam() clears it"""
    monkeypatch.setenv("YOUTH_SLAM_TRAJ_CAPACITY", "10")
    fut.youthSlamSetOptions(1, 4)
    fut.initSlamModule(fut.cfg_path, None)
    results = []
    for i in range(1, 60):
        f = frame(i)  # keep the array alive across the call
        results.append(fut.processSlamFrame(f.ctypes.data, None, W, H, i))
    fut.youthSlamDrain()
    assert results[0] == 1 and results[-1] == 0  # the failure reached the producer
    f1 = frame(1)
    assert fut.processSlamFrame(f1.ctypes.data, None, W, H, 0) == 0
    a, d, t = stats(fut)
    assert t <= 10
    fut.resetSlam()
    assert fut.processSlamFrame(f1.ctypes.data, None, W, H, 0) == 1
    fut.youthSlamDrain()
    assert len(trajectory(fut, 20)[0]) == 1
    fut.stopSlamModule()


def test_default_back_pressure_drops_the_oldest_frames(fut, monkeypatch):
    """SLAM.cpp:158-169: the producer never blocks; more than 10 waiting -> the oldest are dropped down to 5."""
    monkeypatch.setenv("YOUTH_STUB_DELAY_US", "3000")  # a tracker much slower than the producer
    fut.youthSlamSetOptions(0, 2)
    fut.initSlamModule(fut.cfg_path, None)
    n = 120
    t0 = time.time()
    for i in range(1, n + 1):
        f = frame(i)  # keep the array alive across the call
        assert fut.processSlamFrame(f.ctypes.data, None, W, H, i) == 1
    produced_in = time.time() - t0
    fut.youthSlamDrain()
    accepted, dropped, tracked = stats(fut)
    assert accepted == n and dropped > 0 and tracked == n - dropped
    assert produced_in < 0.003 * n * 0.5  # the producer was not held back by the tracker
    poses, ts = trajectory(fut, n)
    marks = list(poses[:, 3])
    assert len(marks) == tracked and marks == sorted(marks) and len(set(marks)) == len(marks)
    assert marks[-1] == n  # the newest frame always survives
    assert list(ts) == [int(m) for m in marks]
    # no slot was written while the tracker held it (the write position must not run round into the frames in
    # flight when frames are dropped faster than they are tracked)
    assert fut.stub_torn_frames() == 0
    fut.stopSlamModule()


def test_stop_while_a_producer_is_pushing(fut, monkeypatch):
    monkeypatch.setenv("YOUTH_STUB_DELAY_US", "200")
    for lossless in (1, 0):
        fut.youthSlamSetOptions(lossless, 4)
        fut.initSlamModule(fut.cfg_path, None)
        assert fut.isSlamModuleRunning() == 1
        pushed = []

        def producer():
            f, i = frame(1), 0
            while fut.processSlamFrame(f.ctypes.data, None, W, H, i) == 1:
                i += 1
            pushed.append(i)

        th = threading.Thread(target=producer)
        th.start()
        time.sleep(0.05)
        fut.stopSlamModule()  # drains what was accepted, then releases the ring; the producer gets 0 from then on
        th.join(timeout=10)
        assert not th.is_alive() and pushed and pushed[0] > 0
        assert fut.isSlamModuleRunning() == 0
        assert fut.stub_torn_frames() == 0


def write_recording(fut, path, n, t0=1000):
    libc = C.CDLL(None)
    libc.fopen.restype, libc.fopen.argtypes = C.c_void_p, [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    fp = libc.fopen(path, b"wb")
    for i in range(n):
        assert fut.youth_bin_write_frame(fp, i, t0 + 33 * i, W, H, frame(i + 1).ctypes.data, None) == 1
    assert fut.youth_bin_write_eof(fp) == 1
    libc.fclose(fp)


def test_algorithm_module_replays_a_recording(fut, tmp_path, monkeypatch):
    rec, prefix = str(tmp_path / "rec.bin").encode(), str(tmp_path / "out")
    write_recording(fut, rec, 40)
    monkeypatch.setenv("YOUTH_SLAM_CONFIG", fut.cfg_path.decode())
    monkeypatch.setenv("YOUTH_SLAM_OUT", prefix)
    fut.youthSlamSetOptions(0, 4)  # algorithmModule switches a replay to lossless by itself
    fut.algorithmModule(rec)  # returns when the recording has been tracked and the module stopped
    assert fut.isSlamModuleRunning() == 0
    rows = np.loadtxt(prefix + "_trajectory.txt")
    assert rows.shape == (40, 8)
    assert np.allclose(rows[:, 1], np.arange(1, 41)) and np.allclose(rows[:, 0], (1000 + 33 * np.arange(40)) / 1000.0)
    fut.algorithmModule(b"/nonexistent/recording.bin")  # reported on stderr, module stopped again
    assert fut.isSlamModuleRunning() == 0
    # a record whose depth payload is not a whole frame ends the replay (nothing is read past the payload)
    import struct

    bad = str(tmp_path / "bad.bin").encode()
    write_recording(fut, bad, 3)
    blob = open(bad, "rb").read()[:-28]  # without the end-of-file record
    blob += struct.pack("<IIHHHxxIII", 3, 1099, 1, W, H, 100, 0, 0) + bytes(100)
    blob += struct.pack("<IIHHHxxIII", 0, 0, 0xFF, 0, 0, 0, 0, 0)
    open(bad, "wb").write(blob)
    fut.algorithmModule(bad)
    assert np.loadtxt(prefix + "_trajectory.txt").shape == (3, 8)


def test_algorithm_module_in_the_viewers_seat_follows_reference_playback(fut, tmp_path, monkeypatch):
    """Whole host chain without a GPU: a recording -> the reference's OWN startPlayback / playbackThread
    (loggingModule.c:505-611, 710-720) -> /logger_viewer_queue -> algorithmModule("mq:...") -> facade -> tracker
    (the test double): every frame arrives once, in order, with the recording's timestamps."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_logging_hooked.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_logging_hooked.so not built (reference tree not mounted)")
    if not os.path.isdir("/dev/mqueue") and not os.path.exists("/proc/sys/fs/mqueue"):
        pytest.skip("POSIX message queues unavailable")
    ref = C.CDLL(so)
    ref.ref_start_playback.argtypes = [C.c_char_p]
    ref.ref_hook_install.argtypes = [C.c_void_p, C.c_void_p]
    ref.ref_hook_install(None, None)
    rec, prefix = str(tmp_path / "rec.bin").encode(), str(tmp_path / "out")
    n = 12
    write_recording(fut, rec, n, t0=5000)
    monkeypatch.setenv("YOUTH_SLAM_CONFIG", fut.cfg_path.decode())
    monkeypatch.setenv("YOUTH_SLAM_OUT", prefix)
    monkeypatch.setenv("YOUTH_SLAM_MQ_IDLE_MS", "1500")
    assert ref.ref_pipeline_start() == 1
    try:
        ref.ref_viewer_stop()
        th = threading.Thread(target=lambda: fut.algorithmModule(b"mq:/logger_viewer_queue"))
        th.start()
        time.sleep(0.2)
        assert ref.ref_start_playback(rec) == 1
        th.join(timeout=30)  # 12 frames at the reference's 30 fps pacing + the idle timeout
        assert not th.is_alive()
    finally:
        ref.ref_pipeline_stop()
    rows = np.loadtxt(prefix + "_trajectory.txt")
    assert rows.shape == (n, 8)
    assert np.allclose(rows[:, 1], np.arange(1, n + 1)) and np.allclose(rows[:, 0], (5000 + 33 * np.arange(n)) / 1000.0)


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_facade_under_sanitizers(fut, tmp_path, sanitizer):
    """Race / memory check of the facade's locking (the reference's plain-bool flags and
    unguarded queue are racy, SLAM.cpp:17,29): tests/stub/facade_race_driver.c pushes frames from one thread, polls
    status from another and drains / resets / saves / stops from a third, in both queue modes, built with
    -fsanitize=thread and -fsanitize=address,undefined.  Any report fails the test."""
    probe = tmp_path / "probe.c"
    probe.write_text("int main(void){return 0;}\n")
    flags = ["-fsanitize=" + sanitizer, "-g", "-O1", "-std=gnu11"]
    ok = subprocess.run(["gcc", *flags, "-o", str(tmp_path / "probe"), str(probe)], capture_output=True).returncode == 0
    if not ok or subprocess.run([str(tmp_path / "probe")], capture_output=True).returncode != 0:
        pytest.skip(f"-fsanitize={sanitizer} does not run here")
    src = [os.path.join(ROOT, "tests", "stub", f) for f in ("facade_race_driver.c", "youth_cuda_stub.c")] + \
          [os.path.join(ROOT, "slam-rgbd_b200", "host", f) for f in HOST_SRC if f != "algorithm_module.c"]
    exe = str(tmp_path / "driver")
    res = subprocess.run(["gcc", *flags, "-I" + os.path.join(ROOT, "include"), "-o", exe, *src, "-lpthread", "-lrt", "-lm"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe, fut.cfg_path.decode(), str(tmp_path / "out")], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.strip() == "ok", run.stderr[-3000:]
    assert "Sanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-3000:]
