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
