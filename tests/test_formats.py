"""Record (.bin) and mq-chunk formats: sizes pinned to the reference header
(Youth.Source/frameDefinitions.h:11-20,45-56,64), our writer/reader/reassembler checked
against the reference's OWN LoggingModule code compiled from its sources (oracle/_ref,
built by oracle/Makefile where /root/reference is mounted; those tests skip elsewhere)."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_logging.so")
libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p
libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
libc.fclose.argtypes = [C.c_void_p]


class FrameHeader(C.Structure):  # mirror of include/frameDefinitions.h
    _fields_ = [("frameId", C.c_uint32), ("timestamp", C.c_uint32), ("frameType", C.c_uint16),
                ("width", C.c_uint16), ("height", C.c_uint16), ("depthDataSize", C.c_uint32),
                ("colorDataSize", C.c_uint32), ("reserved", C.c_uint32)]


class MessageHeader(C.Structure):
    _fields_ = [("msgType", C.c_int), ("width", C.c_int), ("height", C.c_int), ("chunkIndex", C.c_int),
                ("totalChunks", C.c_int), ("dataSize", C.c_int), ("frameId", C.c_int),
                ("timestamp", C.c_uint32), ("ctrlCommand", C.c_int), ("filename", C.c_char * 256)]


@pytest.fixture(scope="module")
def host(pkg):
    L = pkg.host_lib()
    L.youth_bin_write_frame.restype = C.c_int
    L.youth_bin_write_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.youth_bin_write_eof.argtypes = [C.c_void_p]
    L.youth_bin_read_frame.restype = C.c_int
    L.youth_bin_read_frame.argtypes = [C.c_void_p, C.POINTER(FrameHeader), C.c_void_p, C.c_size_t, C.c_void_p,
                                       C.c_size_t]
    return L


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_logging.so not built (reference tree not mounted)")
    L = C.CDLL(REF_SO)
    L.ref_save_frames.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_read_frames.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_size_t]
    L.ref_chunk_stream.argtypes = [C.c_char_p, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int]
    return L


def rand_frames(n, w, h, seed=7):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 9000, size=(n, h, w), dtype=np.uint16),
            rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8))


def write_ours(host, path, depth, color, eof=True):
    f = libc.fopen(path.encode(), b"wb")
    for i in range(depth.shape[0]):
        assert host.youth_bin_write_frame(f, i, 33 * i, depth.shape[2], depth.shape[1], depth[i].ctypes.data,
                                          color[i].ctypes.data if color is not None else None)
    if eof:
        host.youth_bin_write_eof(f)
    libc.fclose(f)


def read_ours(host, path, w, h, max_frames=100, cap=None):
    f = libc.fopen(path.encode(), b"rb")
    out = []
    d = np.empty((h, w), dtype=np.uint16)
    c = np.empty((h, w, 3), dtype=np.uint8)
    hdr = FrameHeader()
    while len(out) < max_frames and host.youth_bin_read_frame(f, C.byref(hdr), d.ctypes.data, cap or d.nbytes,
                                                              c.ctypes.data, c.nbytes):
        out.append((hdr.frameId, hdr.timestamp, hdr.width, hdr.height, d.copy(), c.copy()))
    libc.fclose(f)
    return out


def test_struct_layout():
    assert C.sizeof(FrameHeader) == 28
    assert [getattr(FrameHeader, n).offset for n, _ in FrameHeader._fields_] == [0, 4, 8, 10, 12, 16, 20, 24]
    assert C.sizeof(MessageHeader) == 292
    assert 8192 - C.sizeof(MessageHeader) == 7900


def test_message_arithmetic(host):
    # 78 depth + 117 colour chunks at 640x480, 312 + 467 at 1280x960
    assert host.youth_chunk_count(640 * 480 * 2) == 78
    assert host.youth_chunk_count(640 * 480 * 3) == 117
    assert host.youth_chunk_count(1280 * 960 * 2) == 312
    assert host.youth_chunk_count(1280 * 960 * 3) == 467
    assert host.youth_chunk_count(0) == 0


def test_bin_roundtrip_and_record_size(host, tmp_path):
    depth, color = rand_frames(3, 640, 480)
    p = str(tmp_path / "a.bin")
    write_ours(host, p, depth, color)
    assert os.path.getsize(p) == 3 * 1536028 + 28  # 28 + 614400 + 921600 per frame, + EOF marker
    got = read_ours(host, p, 640, 480)
    assert len(got) == 3  # stops at the EOF marker
    for i, (fid, ts, w, h, d, c) in enumerate(got):
        assert (fid, ts, w, h) == (i, 33 * i, 640, 480)
        assert np.array_equal(d, depth[i]) and np.array_equal(c, color[i])


def test_bin_edge_cases(host, tmp_path):
    depth, color = rand_frames(2, 64, 48)
    p = str(tmp_path / "b.bin")
    write_ours(host, p, depth, color, eof=False)  # no marker: plain end of file also terminates
    assert len(read_ours(host, p, 64, 48)) == 2
    assert read_ours(host, p, 64, 48, cap=100) == []  # payload larger than the caller's buffer is refused
    empty = str(tmp_path / "empty.bin")
    open(empty, "wb").close()
    assert read_ours(host, empty, 64, 48) == []
    with open(p, "r+b") as f:  # truncated payload
        f.truncate(28 + 100)
    assert read_ours(host, p, 64, 48) == []
    # colour omitted -> constant 128 plane, record still format-valid
    q = str(tmp_path / "c.bin")
    write_ours(host, q, depth, None)
    got = read_ours(host, q, 64, 48)
    assert len(got) == 2 and (got[0][5] == 128).all()


def test_packed_depth_records(host, oracle, tmp_path):
    """FRAME_TYPE_DEPTH_PACKED (2): the same record with a YD16 depth payload (include/youth_codec.h);
    the reader is unchanged -- frameType and depthDataSize tell the consumer what it got."""
    host.youth_bin_write_packed_frame.restype = C.c_int
    host.youth_bin_write_packed_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p,
                                                  C.c_uint32, C.c_void_p]
    depth, color = rand_frames(2, 64, 48)
    depth[1] = (1200 + np.arange(64 * 48) // 7).astype(np.uint16).reshape(48, 64)
    p = str(tmp_path / "p.bin")
    f = libc.fopen(p.encode(), b"wb")
    streams = [oracle.codec_encode(d) for d in depth]
    for i, s in enumerate(streams):
        assert host.youth_bin_write_packed_frame(f, i, 33 * i, 64, 48, s.ctypes.data, len(s), color[i].ctypes.data)
    assert host.youth_bin_write_packed_frame(f, 9, 0, 64, 48, streams[0].ctypes.data, 8, None) == 0  # not even a header
    assert host.youth_bin_write_frame(f, 2, 66, 64, 48, depth[0].ctypes.data, None)  # raw and packed records may mix
    host.youth_bin_write_eof(f)
    libc.fclose(f)
    assert os.path.getsize(p) == sum(28 + len(s) + 64 * 48 * 3 for s in streams) + (28 + 64 * 48 * 5) + 28
    f = libc.fopen(p.encode(), b"rb")
    buf = np.empty(1 << 16, dtype=np.uint8)
    c = np.empty((48, 64, 3), dtype=np.uint8)
    hdr = FrameHeader()
    for i, s in enumerate(streams):
        assert host.youth_bin_read_frame(f, C.byref(hdr), buf.ctypes.data, buf.nbytes, c.ctypes.data, c.nbytes) == 1
        assert (hdr.frameId, hdr.timestamp, hdr.frameType, hdr.width, hdr.height) == (i, 33 * i, 2, 64, 48)
        assert hdr.depthDataSize == len(s) and np.array_equal(buf[:len(s)], s) and np.array_equal(c, color[i])
        assert np.array_equal(oracle.codec_decode(buf[:hdr.depthDataSize], 64, 48), depth[i])
    assert host.youth_bin_read_frame(f, C.byref(hdr), buf.ctypes.data, buf.nbytes, None, 0) == 1
    assert hdr.frameType == 1 and hdr.depthDataSize == 64 * 48 * 2
    assert host.youth_bin_read_frame(f, C.byref(hdr), buf.ctypes.data, buf.nbytes, None, 0) == 0
    libc.fclose(f)


def test_bin_large_frame_not_capped(host, tmp_path):
    """the reference playback path caps payloads at 1 MiB (loggingModule.c:528,424-427);
    1280x960 depth is 2.4 MB and must replay here."""
    depth, _ = rand_frames(1, 1280, 960)
    p = str(tmp_path / "big.bin")
    write_ours(host, p, depth, None)
    got = read_ours(host, p, 1280, 960)
    assert len(got) == 1 and np.array_equal(got[0][4], depth[0])


def test_bin_against_reference_reader_and_writer(host, ref, tmp_path):
    depth, color = rand_frames(3, 320, 240)
    ts = (np.arange(3) * 33).astype(np.uint32)
    ours, theirs = str(tmp_path / "ours.bin"), str(tmp_path / "ref.bin")
    write_ours(host, ours, depth, color, eof=False)
    assert ref.ref_save_frames(theirs.encode(), 3, 320, 240, depth.ctypes.data, color.ctypes.data, ts.ctypes.data)
    a, b = np.fromfile(ours, dtype=np.uint8), np.fromfile(theirs, dtype=np.uint8)
    assert a.size == b.size
    rec = 28 + 320 * 240 * 5
    keep = np.ones(a.size, dtype=bool)
    for i in range(3):  # bytes 14..15 of each header are struct padding the reference never initialises
        keep[i * rec + 14:i * rec + 16] = False
    assert np.array_equal(a[keep], b[keep])
    # reference reader on our file (with EOF marker appended)
    write_ours(host, ours, depth, color, eof=True)
    hdrs = (FrameHeader * 8)()
    d = np.zeros((8, 240, 320), dtype=np.uint16)
    c = np.zeros((8, 240, 320, 3), dtype=np.uint8)
    n = ref.ref_read_frames(ours.encode(), 8, 1 << 20, hdrs, d.ctypes.data, c.ctypes.data, d[0].nbytes, c[0].nbytes)
    assert n == 3
    assert np.array_equal(d[:3], depth) and np.array_equal(c[:3], color)
    assert [(h.frameId, h.timestamp, h.width, h.height) for h in hdrs[:3]] == [(i, 33 * i, 320, 240) for i in range(3)]
    # our reader on the reference's file
    got = read_ours(host, theirs, 320, 240)
    assert len(got) == 3 and all(np.array_equal(g[4], depth[i]) for i, g in enumerate(got))


def feed_all(host, msgs, lens):
    r = host.youth_reasm_create()
    done = []
    for k, ln in enumerate(lens):
        done.append(host.youth_reasm_feed(r, msgs[k].ctypes.data, int(ln)))
    return r, done


def test_chunk_roundtrip(host):
    depth, color = rand_frames(1, 640, 480)
    msgs = np.zeros((196, 8192), dtype=np.uint8)
    lens = []
    k = 0
    lens.append(host.youth_chunk_build(msgs[k].ctypes.data, 1, 5, 99, 640, 480, None, 0, 0)); k += 1
    for c in range(78):
        lens.append(host.youth_chunk_build(msgs[k].ctypes.data, 2, 5, 99, 640, 480, depth.ctypes.data, depth.nbytes, c)); k += 1
    for c in range(117):
        lens.append(host.youth_chunk_build(msgs[k].ctypes.data, 3, 5, 99, 640, 480, color.ctypes.data, color.nbytes, c)); k += 1
    assert lens[0] == 292 and lens[1] == 8192 and lens[78] == 292 + 6100 and lens[-1] == 292 + 5200
    r, done = feed_all(host, msgs, lens)
    assert done[-1] == 1 and sum(done) == 1  # complete exactly at the last colour chunk
    d = np.ctypeslib.as_array(C.cast(host.youth_reasm_depth(r), C.POINTER(C.c_uint16)), shape=(480, 640))
    c = np.ctypeslib.as_array(C.cast(host.youth_reasm_color(r), C.POINTER(C.c_uint8)), shape=(480, 640, 3))
    assert np.array_equal(d, depth[0]) and np.array_equal(c, color[0])
    assert host.youth_reasm_feed(r, msgs[0].ctypes.data, 10) == -1  # truncated message is rejected
    host.youth_reasm_destroy(r)


def test_chunks_against_reference_sender(host, ref):
    """the reference's sendMetadata + sendDataInChunks push a frame through a real POSIX mq;
    every message must equal ours (ignoring the ctrlCommand/filename bytes the reference
    leaves uninitialised) and reassemble to the original frame."""
    depth, color = rand_frames(1, 640, 480, seed=11)
    msgs = np.zeros((200, 8192), dtype=np.uint8)
    lens = np.zeros(200, dtype=np.int32)
    n = ref.ref_chunk_stream(b"/youth_test_mq_%d" % os.getpid(), 7, 1234, 640, 480, depth.ctypes.data,
                             color.ctypes.data, msgs.ctypes.data, lens.ctypes.data, 200)
    if n < 0:
        pytest.skip("POSIX message queues unavailable in this container")
    assert n == 196
    mine = np.zeros(8192, dtype=np.uint8)
    for k in range(196):
        if k == 0:
            ln = host.youth_chunk_build(mine.ctypes.data, 1, 7, 1234, 640, 480, None, 0, 0)
        elif k <= 78:
            ln = host.youth_chunk_build(mine.ctypes.data, 2, 7, 1234, 640, 480, depth.ctypes.data, depth.nbytes, k - 1)
        else:
            ln = host.youth_chunk_build(mine.ctypes.data, 3, 7, 1234, 640, 480, color.ctypes.data, color.nbytes, k - 79)
        assert ln == lens[k]
        assert np.array_equal(mine[:32], msgs[k][:32])        # the eight int/uint32 header fields
        assert np.array_equal(mine[292:ln], msgs[k][292:ln])  # payload
    r, done = feed_all(host, msgs, lens[:196])
    assert done[-1] == 1
    d = np.ctypeslib.as_array(C.cast(host.youth_reasm_depth(r), C.POINTER(C.c_uint16)), shape=(480, 640))
    assert np.array_equal(d, depth[0])
    host.youth_reasm_destroy(r)


def test_pose_message_roundtrip(host):
    """pose egress to the viewer queue: MSG_TYPE_POSE = 5, 292 + 56 bytes;
    the frame reassembler ignores it like any unknown type."""
    pose = np.arange(12, dtype=np.float32) * 0.25
    msg = np.zeros(8192, dtype=np.uint8)
    n = host.youth_pose_msg_build(msg.ctypes.data, 41, 1353, pose.ctypes.data, 2, 266000)
    assert n == 292 + 56
    hdr = MessageHeader.from_buffer_copy(msg[:292].tobytes())
    assert (hdr.msgType, hdr.frameId, hdr.timestamp, hdr.dataSize, hdr.totalChunks) == (5, 41, 1353, 56, 1)
    out = np.zeros(14, dtype=np.uint32)
    fid, ts = C.c_int(), C.c_uint32()
    assert host.youth_pose_msg_parse(msg.ctypes.data, n, C.byref(fid), C.byref(ts), out.ctypes.data) == 1
    assert (fid.value, ts.value) == (41, 1353)
    assert np.array_equal(out[:12].view(np.float32), pose) and out[12] == 2 and out[13] == 266000
    assert host.youth_pose_msg_parse(msg.ctypes.data, n - 1, None, None, out.ctypes.data) == 0  # truncated
    msg[0] = 2  # a depth chunk is not a pose message
    assert host.youth_pose_msg_parse(msg.ctypes.data, n, None, None, out.ctypes.data) == 0
    msg[0] = 5
    r = host.youth_reasm_create()
    assert host.youth_reasm_feed(r, msg.ctypes.data, n) == 0
    host.youth_reasm_destroy(r)


def test_host_parsers_survive_fuzzed_input_under_sanitizers(tmp_path):
    """tests/stub/host_fuzz_driver.c: mutated / truncated mq messages, recordings and camera YAML files against
    youth_reasm_feed, youth_pose_msg_parse, youth_bin_read_frame and youth_config_from_yaml, built with
    -fsanitize=address,undefined (+ float-cast-overflow).  They may reject, never misbehave."""
    import subprocess

    flags = ["-fsanitize=address,undefined", "-fsanitize=float-cast-overflow", "-fno-sanitize-recover=undefined", "-g", "-O1",
             "-std=gnu11"]
    probe = tmp_path / "probe.c"
    probe.write_text("int main(void){return 0;}\n")
    ok = subprocess.run(["gcc", *flags, "-o", str(tmp_path / "probe"), str(probe)], capture_output=True).returncode == 0
    if not ok or subprocess.run([str(tmp_path / "probe")], capture_output=True).returncode != 0:
        pytest.skip("sanitizer builds do not run here")
    src = [os.path.join(ROOT, "tests", "stub", f) for f in ("host_fuzz_driver.c", "youth_cuda_stub.c")] + \
          [os.path.join(ROOT, "slam-rgbd_b200", "host", f) for f in ("youth_frameio.c", "youth_config.c")]
    exe = str(tmp_path / "fuzz")
    res = subprocess.run(["gcc", *flags, "-I" + os.path.join(ROOT, "include"), "-o", exe, *src, "-lrt", "-lm"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    scratch = tmp_path / "scratch"
    scratch.mkdir()
    run = subprocess.run([exe, str(scratch)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.startswith("ok "), (run.stdout + run.stderr)[-3000:]
    assert "runtime error" not in run.stderr and "Sanitizer" not in run.stderr, run.stderr[-3000:]
