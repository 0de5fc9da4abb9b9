"""Committed fixture tests/golden/golden_ext_160x120.npz (made by tests/golden/make_golden_ext.py):
YD16 streams, a frame-to-model trajectory, a fused volume and ray-cast model maps for the four
frames of golden_160x120.npz.  CPU: the C statements still produce it.  GPU (-m gpu): the device
produces it too, through the C ABI, without the statements being re-run."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
W, H = 160, 120
SMALL = dict(width=W, height=H, fx=570.3 * W / 640, fy=570.3 * W / 640, cx=W / 2.0, cy=H / 2.0)
TSDF = dict(dim=(64, 32, 64), voxel_m=0.1, origin=(-3.2, -1.6, -1.2), trunc_m=0.3)
IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    return (np.load(os.path.join(HERE, "golden", "golden_160x120.npz")),
            np.load(os.path.join(HERE, "golden", "golden_ext_160x120.npz")))


def test_cpu_statements_reproduce_the_fixture(oracle, gold):
    g, e = gold
    frames = g["frames"]
    cfg, t = oracle.default_config(**SMALL), oracle.tsdf_config(**TSDF)
    streams = [oracle.codec_encode(f) for f in frames]
    assert np.array_equal(streams[0], e["yd16_stream_0"])
    assert [len(s) for s in streams] == list(e["yd16_sizes"]) and [sha(s) for s in streams] == list(e["yd16_sha"])
    assert np.array_equal(oracle.codec_decode(e["yd16_stream_0"], W, H), frames[0])
    poses, status = oracle.track_sequence_model(cfg, t, frames)
    assert np.array_equal(poses.view(np.uint32), e["model_poses"].view(np.uint32)) and np.array_equal(status, e["model_status"])
    f0 = oracle.OFrame(cfg, frames[0])
    vol = oracle.tsdf_new(t)
    oracle.tsdf_integrate(cfg, t, vol, f0.depth(0), IDENT)
    assert sha(vol) == str(e["volume_sha"]) and int((vol[..., 1] > 0).sum()) == int(e["volume_observed"])
    for level in range(3):
        vm, nm = oracle.tsdf_raycast(cfg, t, vol, IDENT, level, hint=f0.depth(level))
        assert sha(vm) == str(e[f"model_vmap_sha_l{level}"]) and sha(nm) == str(e[f"model_nmap_sha_l{level}"])


@pytest.mark.gpu
def test_device_reproduces_the_fixture(pkg, gold):
    from slam_rgbd_b200 import binding as B

    g, e = gold
    frames = np.ascontiguousarray(g["frames"])
    cd = pkg.Codec(W, H, max_frames=4)
    packed, offs = cd.encode(frames)
    assert list(np.diff(offs.astype(np.int64))) == list(e["yd16_sizes"])
    assert np.array_equal(packed[:int(offs[1])], e["yd16_stream_0"])
    assert [sha(packed[int(offs[i]):int(offs[i + 1])]) for i in range(4)] == list(e["yd16_sha"])
    assert np.array_equal(cd.decode(packed, offs), frames)
    cd.close()
    trk = B.Tracker(pkg.default_config(batch=4, traj_capacity=8, **SMALL))
    trk.enable_model(pkg.tsdf_config(**TSDF))
    poses = trk.track_batch([frames[:1]])[0]  # frame 0: fused at the identity, model ray cast with its depth as hint
    assert np.array_equal(poses[0], IDENT)
    assert sha(trk.read_volume()) == str(e["volume_sha"])
    for level in range(3):
        assert sha(trk.read_model(B.DBG_VERTEX, level)) == str(e[f"model_vmap_sha_l{level}"])
        assert sha(trk.read_model(B.DBG_NORMAL, level)) == str(e[f"model_nmap_sha_l{level}"])
    rest = trk.track_batch([frames[1:]])[0]
    got = np.concatenate([poses, rest])
    assert np.array_equal(got.view(np.uint32), e["model_poses"].view(np.uint32))
    _, _, st = trk.trajectory()
    assert np.array_equal(st, e["model_status"])
    trk.close()
