"""The reference's OWN LoggingModule thread with the one-line hook of INTEGRATION.md section 3 applied
(oracle/_ref/libref_logging_hooked.so: loggingModule.c piped through sed by oracle/Makefile, nothing
copied into the repository).  A fake sensor pushes frames into /sensor_logger_queue with the reference's
sendMetadata / sendDataInChunks, the reference loggerThread reassembles them and reaches the hook, a fake
viewer drains the pass-through queue.  CPU: the hook receives exactly the frames that were sent.
GPU (-m gpu): the hook is the real facade (processSlamFrame) and the trajectory equals direct tracking."""
import ctypes as C
import os
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libref_logging_hooked.so")
RUNNING = C.CFUNCTYPE(C.c_int)
PROCESS = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32)


@pytest.fixture()
def ref():
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libref_logging_hooked.so not built (reference tree not mounted)")
    try:
        import posix_ipc  # noqa: F401
    except Exception:
        pass
    if not os.path.isdir("/dev/mqueue") and not os.path.exists("/proc/sys/fs/mqueue"):
        pytest.skip("POSIX message queues unavailable")
    L = C.CDLL(SO)
    L.ref_hook_install.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_fake_sensor_send.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_hook_calls.restype = C.c_long
    L.ref_viewer_messages.restype = C.c_long
    L.ref_start_playback.argtypes = [C.c_char_p]
    return L


def wait_for(cond, timeout=20.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if cond():
            return True
        time.sleep(0.01)
    return False


def test_reference_logger_thread_reaches_the_hook_with_whole_frames(ref):
    rng = np.random.default_rng(3)
    sent, got = [], []

    def process(d, c, w, h, ts):
        depth = np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint16)), shape=(h, w)).copy()
        color = np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_uint8)), shape=(h, w, 3)).copy()
        got.append((depth, color, w, h, ts))
        return 1

    cb_run, cb_proc = RUNNING(lambda: 1), PROCESS(process)
    ref.ref_hook_install(C.cast(cb_run, C.c_void_p), C.cast(cb_proc, C.c_void_p))
    assert ref.ref_pipeline_start() == 1
    try:
        for i, (w, h) in enumerate([(64, 48), (64, 48), (640, 480), (64, 48)]):  # the size may change between frames
            d = rng.integers(0, 9000, size=(h, w), dtype=np.uint16)
            c = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
            sent.append((d, c, w, h, 33 * i))
            assert ref.ref_fake_sensor_send(i, 33 * i, w, h, d.ctypes.data, c.ctypes.data) == 1
        assert wait_for(lambda: len(got) == 4)
    finally:
        ref.ref_pipeline_stop()
    assert ref.ref_hook_calls() == 4  # once per frame, at the frame-complete test (loggingModule.c:354)
    for (d, c, w, h, ts), (gd, gc, gw, gh, gts) in zip(sent, got):
        assert (w, h, ts) == (gw, gh, gts) and np.array_equal(d, gd) and np.array_equal(c, gc)
    # pass-through to the viewer queue kept running: 1 metadata + depth + colour chunks per frame
    assert ref.ref_viewer_messages() >= 3 * (1 + 1 + 2) + (1 + 78 + 117)


def test_reference_playback_thread_feeds_the_queue_consumer(pkg, ref, tmp_path):
    """Playback direction: a recording (written by this repository's record writer) is replayed by the reference's
    OWN playbackThread (loggingModule.c:505-611, started by its own startPlayback) into /logger_viewer_queue, where
    youth_mq_consume sits in the viewer's seat (viewerModule.c:160-250) and hands whole frames to a
    processSlamFrame-shaped sink: every frame arrives once, in order, bit-identical, sizes changing in between."""
    import threading

    host = pkg.host_lib()
    rng = np.random.default_rng(5)
    path = str(tmp_path / "rec.bin").encode()
    sent = []
    f = C.CDLL(None).fopen
    f.restype, f.argtypes = C.c_void_p, [C.c_char_p, C.c_char_p]
    fclose = C.CDLL(None).fclose
    fclose.argtypes = [C.c_void_p]
    host.youth_bin_write_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    host.youth_bin_write_eof.argtypes = [C.c_void_p]
    fp = f(path, b"wb")
    for i, (w, h) in enumerate([(64, 48), (640, 480), (640, 480), (32, 24), (64, 48)]):  # 640x480 is the reader's 1 MiB limit
        d = rng.integers(0, 9000, size=(h, w), dtype=np.uint16)
        c = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        sent.append((d, c, w, h, 1000 + 33 * i))
        assert host.youth_bin_write_frame(fp, i, 1000 + 33 * i, w, h, d.ctypes.data, c.ctypes.data) == 1
    assert host.youth_bin_write_eof(fp) == 1
    fclose(fp)

    got = []

    def sink(d, c, w, h, ts):
        depth = np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint16)), shape=(h, w)).copy()
        color = np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_uint8)), shape=(h, w, 3)).copy()
        got.append((depth, color, w, h, ts))
        return 1

    cb_sink = PROCESS(sink)
    cb_run, cb_proc = RUNNING(lambda: 0), PROCESS(lambda *a: 1)
    ref.ref_hook_install(C.cast(cb_run, C.c_void_p), C.cast(cb_proc, C.c_void_p))
    assert ref.ref_pipeline_start() == 1
    stop = C.c_int(0)
    result = {}
    try:
        ref.ref_viewer_stop()  # the consumer under test reads the viewer queue instead of the fake viewer
        t = threading.Thread(
            target=lambda: result.update(n=host.youth_mq_consume(b"/logger_viewer_queue", C.cast(cb_sink, C.c_void_p),
                                                                 C.byref(stop), 0)))
        t.start()
        assert ref.ref_start_playback(path) == 1
        assert wait_for(lambda: len(got) == len(sent))  # paced at 30 fps by the reference (loggingModule.c:593)
        assert wait_for(lambda: ref.ref_is_playing_back() == 0)  # the EOF record ends the playback
        stop.value = 1
        t.join(timeout=10)
        assert not t.is_alive()
    finally:
        stop.value = 1
        ref.ref_pipeline_stop()
    assert result["n"] == len(sent) == len(got)
    for (d, c, w, h, ts), (gd, gc, gw, gh, gts) in zip(sent, got):
        assert (w, h, ts) == (gw, gh, gts) and np.array_equal(d, gd) and np.array_equal(c, gc)
    assert host.youth_mq_consume(b"/no_such_queue_here", C.cast(cb_sink, C.c_void_p), None, 100) == -1


@pytest.mark.gpu
def test_sensor_to_logging_to_algorithm_drop_in(pkg, small_seq, ref):
    """Sensor (fake) -> Logging (the reference's own thread, hooked) -> Algorithm (this repository) -> poses"""
    frames, _ = small_seq
    host = pkg.host_lib()
    host.youthSlamSetOptions(1, 4)
    host.initSlamModule(None, None)
    assert host.isSlamModuleRunning() == 1
    ref.ref_hook_install(C.cast(host.isSlamModuleRunning, C.c_void_p), C.cast(host.processSlamFrame, C.c_void_p))
    color = np.full((480, 640, 3), 128, dtype=np.uint8)
    assert ref.ref_pipeline_start() == 1
    try:
        for i in range(6):
            assert ref.ref_fake_sensor_send(i, 33 * i, 640, 480, frames[i].ctypes.data, color.ctypes.data) == 1
        assert wait_for(lambda: ref.ref_hook_calls() == 6)
    finally:
        ref.ref_pipeline_stop()
    host.youthSlamDrain()
    poses = np.empty((6, 12), dtype=np.float32)
    ts = np.empty(6, dtype=np.uint32)
    assert host.youthSlamGetTrajectory(poses.ctypes.data, ts.ctypes.data, None, 6) == 6
    host.stopSlamModule()
    assert list(ts) == [33 * i for i in range(6)]
    from slam_rgbd_b200.binding import Tracker

    direct = Tracker(pkg.default_config(batch=6))
    want = direct.track_batch([frames])[0]
    direct.close()
    assert np.array_equal(poses.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_playback_to_algorithm_module_on_the_viewer_queue(pkg, small_seq, ref, tmp_path):
    """Recording -> the reference's own startPlayback / playbackThread -> /logger_viewer_queue ->
    algorithmModule("mq:/logger_viewer_queue") -> tracker: the TUM trajectory equals direct tracking of the frames."""
    import subprocess
    import threading

    frames, _ = small_seq
    paths = pkg.lib_paths()
    rec = str(tmp_path / "rec.bin")
    assert subprocess.run([paths["harness"], "gen", rec, "6"], capture_output=True).returncode == 0  # sequence 0 = small_seq
    host = pkg.host_lib()
    prefix = str(tmp_path / "run")
    os.environ["YOUTH_SLAM_OUT"] = prefix
    os.environ["YOUTH_SLAM_MQ_IDLE_MS"] = "3000"
    cb_run, cb_proc = RUNNING(lambda: 0), PROCESS(lambda *a: 1)
    ref.ref_hook_install(C.cast(cb_run, C.c_void_p), C.cast(cb_proc, C.c_void_p))
    assert ref.ref_pipeline_start() == 1
    try:
        ref.ref_viewer_stop()
        host.youthSlamSetOptions(1, 4)
        th = threading.Thread(target=lambda: host.algorithmModule(C.c_char_p(b"mq:/logger_viewer_queue")))
        th.start()
        assert wait_for(lambda: host.isSlamModuleRunning() == 1, timeout=120)  # first CUDA init on a fresh box is slow
        assert ref.ref_start_playback(rec.encode()) == 1
        th.join(timeout=120)
        assert not th.is_alive()
    finally:
        ref.ref_pipeline_stop()
        del os.environ["YOUTH_SLAM_OUT"], os.environ["YOUTH_SLAM_MQ_IDLE_MS"]
    rows = np.loadtxt(prefix + "_trajectory.txt")
    assert rows.shape == (6, 8)
    from slam_rgbd_b200.binding import Tracker

    direct = Tracker(pkg.default_config(batch=6))
    want = direct.track_batch([frames])[0]
    direct.close()
    assert np.allclose(rows[:, 1:4], want[:, [3, 7, 11]], atol=1e-6)
