"""YD16 lossless depth codec (include/youth_codec.h).

CPU part: the C statement of the format (oracle/youth_codec_oracle.c) against hand-computed
known-answer streams, an independent pure-Python/numpy decoder written from the header's layout
description, round trips over edge cases, and rejection of malformed streams.
GPU part (-m gpu): the device encoder must emit byte-identical streams, both decoders must return
the original frames, the packed tracker input must give bit-identical poses to raw input.
The reference stores depth raw (loggingModule.c:118-127), so raw frames are the ground truth."""
import ctypes as C
import struct

import numpy as np
import pytest


def py_decode(stream, w, h):
    """independent decoder, straight from the layout comment in include/youth_codec.h"""
    b = bytes(stream)
    magic, ww, hh, nb, pay = struct.unpack_from("<IHHII", b, 0)
    assert magic == 0x36314459 and (ww, hh) == (w, h) and nb == (w * h + 31) // 32
    assert len(b) == 16 + nb + pay
    sizes = b[16:16 + nb]
    pos = 16 + nb
    out = np.zeros(nb * 32, dtype=np.uint16)
    for k in range(nb):
        blk = b[pos:pos + sizes[k]]
        pos += sizes[k]
        (mask,) = struct.unpack_from("<I", blk, 0)
        if not mask:
            assert len(blk) == 4
            continue
        first, bits = struct.unpack_from("<HB", blk, 4)
        code = int.from_bytes(blk[7:], "little")
        cur, seen = first, 0
        for i in range(32):
            if (mask >> i) & 1:
                if seen:
                    z = (code >> ((seen - 1) * bits)) & ((1 << bits) - 1)
                    d = (z >> 1) ^ -(z & 1)
                    cur = (cur + d) & 0xFFFF
                seen += 1
                out[k * 32 + i] = cur
        assert len(blk) == 7 + ((seen - 1) * bits + 7) // 8
    assert pos == len(b)
    return out[:w * h].reshape(h, w)


def edge_frames(w, h, rng):
    n = w * h
    ramp = (np.arange(n) % 65536).astype(np.uint16).reshape(h, w)
    alt = np.where(np.arange(n) % 2 == 0, 1, 65535).astype(np.uint16).reshape(h, w)  # 16-bit deltas
    sparse = np.zeros(n, dtype=np.uint16)
    sparse[rng.integers(0, n, size=max(1, n // 50))] = 1234
    single = np.zeros(n, dtype=np.uint16)
    single[n - 1] = 65535
    return {
        "zeros": np.zeros((h, w), dtype=np.uint16),
        "const": np.full((h, w), 1000, dtype=np.uint16),
        "max": np.full((h, w), 65535, dtype=np.uint16),
        "ramp": ramp,
        "alternating_extremes": alt,
        "sparse": sparse.reshape(h, w),
        "single_last_pixel": single.reshape(h, w),
        "random": rng.integers(0, 65536, size=(h, w), dtype=np.uint16),
        "random_small_steps": (2000 + np.cumsum(rng.integers(-3, 4, size=n))).astype(np.uint16).reshape(h, w),
    }


# --------------------------------------------------------------------------- CPU


def test_known_answer_streams(oracle):
    # one block, 32 pixels: zeros except three readings -> mask, first, 3-bit zig-zag deltas
    f = np.zeros((1, 32), dtype=np.uint16)
    f[0, 2], f[0, 5], f[0, 31] = 1000, 1003, 1001        # deltas +3 -> zz 6, -2 -> zz 3; 3 bits
    s = bytes(oracle.codec_encode(f))
    mask = (1 << 2) | (1 << 5) | (1 << 31)
    want = struct.pack("<IHHII", 0x36314459, 32, 1, 1, 8) + bytes([8]) + struct.pack("<IHB", mask, 1000, 3) + bytes([6 | (3 << 3)])
    assert s == want
    # an all-zero block costs its mask only; a block with a single reading has no delta bytes
    f = np.zeros((2, 32), dtype=np.uint16)
    f[1, 7] = 42
    s = bytes(oracle.codec_encode(f))
    want = (struct.pack("<IHHII", 0x36314459, 32, 2, 2, 11) + bytes([4, 7]) + struct.pack("<I", 0)
            + struct.pack("<IHB", 1 << 7, 42, 0))
    assert s == want
    # constant run: zero-width deltas
    f = np.full((1, 32), 777, dtype=np.uint16)
    s = bytes(oracle.codec_encode(f))
    assert s == struct.pack("<IHHII", 0x36314459, 32, 1, 1, 7) + bytes([7]) + struct.pack("<IHB", 0xFFFFFFFF, 777, 0)
    # wrap-around delta: 65535 -> 1 is +2 modulo 2^16 (zz 4, 3 bits), not -65534
    f = np.zeros((1, 32), dtype=np.uint16)
    f[0, 0], f[0, 1] = 65535, 1
    s = bytes(oracle.codec_encode(f))
    assert s[-4:] == struct.pack("<HB", 65535, 3) + bytes([4])


@pytest.mark.parametrize("shape", [(32, 1), (33, 1), (31, 3), (160, 120), (100, 7), (640, 480)])
def test_cpu_round_trip_and_independent_decoder(oracle, shape):
    w, h = shape
    rng = np.random.default_rng(w * 1000 + h)
    for name, f in edge_frames(w, h, rng).items():
        s = oracle.codec_encode(f)
        assert len(s) <= oracle.lib().yc_max_bytes(w, h), name
        d = oracle.codec_decode(s, w, h)
        assert d is not None and np.array_equal(d, f), name
        if w * h <= 160 * 120:
            assert np.array_equal(py_decode(s, w, h), f), name


def test_worst_case_bound_is_reached(oracle):
    w, h = 64, 4
    f = np.where(np.arange(w * h) % 2 == 0, 1, 32769).astype(np.uint16).reshape(h, w)  # every delta needs 16 bits
    s = oracle.codec_encode(f)
    assert len(s) == oracle.lib().yc_max_bytes(w, h) == 16 + 8 + 8 * 69


def test_synthetic_depth_compresses(pkg, oracle):
    f = pkg.synth_sequence(2)
    s = oracle.codec_encode(f[1])
    assert len(s) < f[1].nbytes / 2.5  # about 3x on the Astra-shaped frames of the benchmark
    assert np.array_equal(oracle.codec_decode(s, 640, 480), f[1])


def test_malformed_streams_are_rejected(oracle):
    rng = np.random.default_rng(5)
    f = (1500 + np.cumsum(rng.integers(-2, 3, size=64 * 8))).astype(np.uint16).reshape(8, 64)
    f[2, 10:20] = 0
    good = oracle.codec_encode(f)
    assert np.array_equal(oracle.codec_decode(good, 64, 8), f)

    def bad(mut):
        s = good.copy()
        mut(s)
        return oracle.codec_decode(s, 64, 8)

    assert oracle.codec_decode(good[:-1], 64, 8) is None              # truncated
    assert oracle.codec_decode(good[:10], 64, 8) is None              # shorter than a header
    assert oracle.codec_decode(good, 32, 16) is None                  # wrong geometry
    assert bad(lambda s: s.__setitem__(0, 0x5A)) is None              # magic
    assert bad(lambda s: s.__setitem__(8, s[8] + 1)) is None          # block count
    assert bad(lambda s: s.__setitem__(12, s[12] ^ 1)) is None        # payload size
    assert bad(lambda s: s.__setitem__(16, s[16] + 1)) is None        # size table vs payload
    nb = 16
    assert bad(lambda s: s.__setitem__(16 + nb + 6, 17)) is None      # bit width > 16
    assert bad(lambda s: s.__setitem__(16 + nb + 6, (s[16 + nb + 6] + 1) % 17)) is None  # width vs block size


def test_codec_symbols_exported(pkg):
    from test_cabi import declared_functions

    names = declared_functions("youth_codec.h")
    assert "youth_codec_encode" in names and "youth_cuda_track_batch_packed" in names and len(names) >= 8
    lib = C.CDLL(pkg.lib_paths()["cuda"])
    for n in names:
        assert hasattr(lib, n), f"libyouth_cuda.so does not export {n}"
    assert lib.youth_codec_max_bytes.restype is not None
    lib.youth_codec_max_bytes.restype = C.c_size_t
    assert lib.youth_codec_max_bytes(640, 480) == 16 + 9600 + 9600 * 69


# --------------------------------------------------------------------------- GPU


def split(packed, offs):
    return [packed[int(offs[i]):int(offs[i + 1])] for i in range(len(offs) - 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(32, 1), (33, 1), (31, 3), (100, 7), (160, 120), (640, 480), (1280, 960)])
def test_device_encoder_is_byte_identical_and_round_trips(pkg, oracle, shape):
    w, h = shape
    rng = np.random.default_rng(w * 1000 + h)
    frames = np.stack(list(edge_frames(w, h, rng).values()))
    cd = pkg.Codec(w, h, max_frames=len(frames))
    packed, offs = cd.encode(frames)
    assert offs[0] == 0 and offs[-1] == len(packed)
    for f, s in zip(frames, split(packed, offs)):
        assert bytes(s) == bytes(oracle.codec_encode(f))               # canonical stream, byte for byte
    back = cd.decode(packed, offs)
    assert np.array_equal(back, frames)
    # streams packed by the CPU statement decode on the device too (arbitrary byte alignment of each stream)
    cpu = [oracle.codec_encode(f) for f in frames]
    offs2 = np.concatenate([[0], np.cumsum([len(s) for s in cpu])]).astype(np.uint64)
    assert np.array_equal(cd.decode(np.concatenate(cpu), offs2), frames)
    cd.close()


@pytest.mark.gpu
def test_device_codec_on_synthetic_sequence(pkg, oracle):
    frames = pkg.synth_sequence(24, noise=1)
    cd = pkg.Codec(640, 480, max_frames=24)
    packed, offs = cd.encode(frames)
    assert len(packed) < frames.nbytes / 2.2
    for i in (0, 7, 23):
        assert bytes(packed[int(offs[i]):int(offs[i + 1])]) == bytes(oracle.codec_encode(frames[i]))
    assert np.array_equal(cd.decode(packed, offs), frames)
    # sub-batches and repeated use of one context
    p2, o2 = cd.encode(frames[5:9])
    assert bytes(p2) == bytes(packed[int(offs[5]):int(offs[9])])
    assert np.array_equal(cd.decode(p2, o2), frames[5:9])
    cd.close()


@pytest.mark.gpu
def test_device_decoder_rejects_malformed_streams(pkg, oracle):
    rng = np.random.default_rng(9)
    f = (1500 + np.cumsum(rng.integers(-2, 3, size=160 * 120))).astype(np.uint16).reshape(1, 120, 160)
    cd = pkg.Codec(160, 120, max_frames=2)
    packed, offs = cd.encode(f)
    nb = 160 * 120 // 32

    def rejects(mut, o=offs):
        s = packed.copy()
        mut(s)
        with pytest.raises(RuntimeError):
            cd.decode(s, o)

    rejects(lambda s: s.__setitem__(0, 0x5A))                          # magic (host check)
    rejects(lambda s: s.__setitem__(12, s[12] ^ 1))                    # payload size (host check)
    rejects(lambda s: s.__setitem__(16 + 5, s[16 + 5] + 1))            # size table vs payload (device check)
    rejects(lambda s: s.__setitem__(16 + nb + 6, 17))                  # bit width (device check)
    rejects(lambda s: None, o=np.array([0, len(packed) - 1], dtype=np.uint64))  # truncated
    assert np.array_equal(cd.decode(packed, offs), f)                  # the context still works afterwards
    cd.close()


@pytest.mark.gpu
def test_full_size_properties(pkg):
    """BASELINE-size batch: 300 frames; encode -> decode is the identity and offsets are consistent."""
    frames = pkg.synth_sequence(300)
    cd = pkg.Codec(640, 480, max_frames=300)
    packed, offs = cd.encode(frames)
    assert np.all(np.diff(offs.astype(np.int64)) > 16 + 9600) and offs[-1] == len(packed)
    assert 2.5 < frames.nbytes / len(packed) < 4.0
    assert np.array_equal(cd.decode(packed, offs), frames)
    cd.close()


@pytest.mark.gpu
def test_tracker_fed_from_packed_streams_matches_raw_input(pkg):
    from slam_rgbd_b200 import binding as B

    frames = np.stack([pkg.synth_sequence(40, sequence=s) for s in range(2)])
    cd = pkg.Codec(640, 480, max_frames=40)
    enc = [cd.encode(frames[s]) for s in range(2)]
    cd.close()
    cfg = pkg.default_config(batch=24, n_streams=2, traj_capacity=64)
    raw = B.Tracker(cfg)
    a = np.concatenate([raw.track_batch([frames[0][:24], frames[1][:24]]),
                        raw.track_batch([frames[0][24:], frames[1][24:]])], axis=1)
    raw.close()
    pk = B.Tracker(cfg)
    streams = [e[0] for e in enc]
    b1 = pk.track_batch_packed(streams, [e[1][:25].copy() for e in enc], 24)
    b2 = pk.track_batch_packed(streams, [e[1][24:].copy() for e in enc], 16)
    b = np.concatenate([b1, b2], axis=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # a corrupted stream is reported, not silently tracked
    bad = streams[0].copy()
    bad[int(enc[0][1][3]) + 16 + 9600 + 6] = 17
    pk.reset()
    with pytest.raises(RuntimeError, match="malformed"):
        pk.track_batch_packed([bad, streams[1]], [e[1][:25].copy() for e in enc], 24)
    pk.close()
