"""Frame-to-model tracking (include/youth_model.h).

CPU part: the C statement (oracle/youth_tsdf_oracle.c) against closed-form cases -- a fronto-parallel
plane fuses to tsdf = (1 - z)/mu and ray-casts back to z = 1 with the normal convention of stage 2 --
a numpy restatement of the fusion rule, and the drift of the model tracker versus frame-to-frame.
GPU part (-m gpu): fused volumes, ray-cast model maps and whole trajectories must be BIT-identical
to the CPU statement; full-size properties (determinism, reset, batching independence, accuracy).
The reference has no volumetric model (its map lives in un-vendored ORB-SLAM3, SLAM.cpp:54):
parity unpinned, the oracle defines the arithmetic."""
import ctypes as C

import numpy as np
import pytest

SMALL = dict(width=160, height=120, fx=570.3 / 4, fy=570.3 / 4, cx=80.0, cy=60.0)
SMALL_T = dict(dim=(64, 32, 64), voxel_m=0.1, origin=(-3.2, -1.6, -1.2), trunc_m=0.3)
IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)


def small_cfgs(oracle, **kw):
    return oracle.default_config(**SMALL, **kw), oracle.tsdf_config(**SMALL_T)


def pose_of(rx=0.0, ry=0.0, rz=0.0, t=(0, 0, 0)):
    from scipy.spatial.transform import Rotation

    R = Rotation.from_euler("xyz", [rx, ry, rz]).as_matrix()
    return np.concatenate([R, np.array(t, dtype=np.float64).reshape(3, 1)], axis=1).astype(np.float32).reshape(12)


# --------------------------------------------------------------------------- CPU


def test_defaults_agree_between_product_and_oracle(pkg, oracle):
    a, b = pkg.tsdf_config(), oracle.tsdf_config()
    assert C.sizeof(type(a)) == C.sizeof(type(b)) == 44
    assert bytes(a) == bytes(b)
    assert list(a.dim) == [256, 128, 256] and a.max_weight == 64


def test_plane_fuses_and_raycasts_in_closed_form(oracle):
    cfg, t = small_cfgs(oracle, bilateral=0)
    depth0 = np.full((120, 160), 1000.0, dtype=np.float32)  # plane z = 1 m seen from the identity pose
    vol = oracle.tsdf_new(t)
    assert (vol[..., 0] == 32767).all() and (vol[..., 1] == 0).all()
    oracle.tsdf_integrate(cfg, t, vol, depth0, IDENT)
    # voxel centres along the optical axis: x = y = 0.05 (ix = 32, iy = 16), z = -1.15 + 0.1 k
    col = vol[:, 16, 32, :]
    zc = np.float32(-1.2) + (np.arange(64, dtype=np.float32) + np.float32(0.5)) * np.float32(0.1)
    for k in range(64):
        z = float(zc[k])
        if 0 < z < 0.13:
            continue  # the voxel centre (0.05, 0.05, z) projects outside the 160x120 image this close
        if z <= 0 or 1.0 - z < -0.3:  # behind the camera / beyond the truncation band: untouched
            assert tuple(col[k]) == (32767, 0), k
        else:
            want = min(1.0, (1.0 - z) / 0.3)
            assert col[k, 1] == 1 and abs(col[k, 0] / 32767.0 - want) < 2e-4, k
    # a second observation of the same plane leaves the tsdf where it is and raises the weight
    before = vol.copy()
    oracle.tsdf_integrate(cfg, t, vol, depth0, IDENT)
    seen = before[..., 1] > 0
    assert (vol[..., 1][seen] == 2).all() and np.abs(vol[..., 0].astype(int) - before[..., 0]).max() <= 1
    vmap, nmap = oracle.tsdf_raycast(cfg, t, vol, IDENT, 0)
    ok = vmap[..., 3] > 0
    # one view only: rays within a voxel (10 cm = 14 px here) of the frustum border touch unobserved voxels
    assert ok[20:100, 24:136].all() and ok.mean() > 0.6
    assert np.abs(vmap[..., 2][ok] - 1.0).max() < 2e-3  # the zero crossing is the plane
    u, v = np.meshgrid(np.arange(160, dtype=np.float32), np.arange(120, dtype=np.float32))
    assert np.allclose(vmap[..., 0][ok], ((u - 80.0) * vmap[..., 2] / np.float32(570.3 / 4))[ok], rtol=1e-6)
    nok = nmap[..., 3] > 0
    assert nok[28:92, 40:120].all() and np.abs(nmap[..., :3][nok] - [0, 0, 1]).max() < 2e-2  # stage-2 convention: away from the camera


def test_model_normals_follow_the_frame_convention(pkg, oracle):
    cfg, t = small_cfgs(oracle)
    f = pkg.synth_sequence(5, 160, 120)
    gt = pkg.synth_gt(5, 160, 120).astype(np.float32)
    fr = oracle.OFrame(cfg, f[0])
    vol = oracle.tsdf_new(t)
    for i in range(5):  # several views at their true poses fill the 2 % dropout holes of a single view
        oracle.tsdf_integrate(cfg, t, vol, oracle.OFrame(cfg, f[i]).depth(0), gt[i])
    vmap, nmap = oracle.tsdf_raycast(cfg, t, vol, IDENT, 0)
    both = (nmap[..., 3] > 0) & (fr.nmap(0)[..., 3] > 0)
    assert both.mean() > 0.35  # 10 cm voxels: grazing surfaces (floor, ceiling) keep holes, walls and objects are dense
    dots = (nmap[..., :3] * fr.nmap(0)[..., :3]).sum(-1)[both]
    assert np.median(dots) > 0.97 and (dots > 0).mean() > 0.98
    vboth = (vmap[..., 3] > 0) & (fr.vmap(0)[..., 3] > 0)
    assert np.median(np.abs(vmap[..., 2] - fr.vmap(0)[..., 2])[vboth]) < 0.02  # model surface = measured surface


def test_fusion_rule_against_numpy(pkg, oracle):
    """independent float32 numpy restatement of yo_tsdf_integrate for an arbitrary pose"""
    cfg, t = small_cfgs(oracle)
    f = pkg.synth_sequence(2, 160, 120)
    d0 = oracle.OFrame(cfg, f[1]).depth(0).copy()
    pose = pose_of(0.02, -0.03, 0.01, (0.05, -0.02, 0.03))
    vol = oracle.tsdf_new(t)
    oracle.tsdf_integrate(cfg, t, vol, oracle.OFrame(cfg, f[0]).depth(0), IDENT)
    before = vol.copy()
    oracle.tsdf_integrate(cfg, t, vol, d0, pose)
    f32 = np.float32
    P = pose.reshape(3, 4).astype(np.float64)
    Ri, ti = P[:, :3].T, -P[:, :3].T @ P[:, 3]
    idx = np.stack(np.meshgrid(np.arange(64), np.arange(32), np.arange(64), indexing="ij"), -1)  # [x][y][z]
    w = np.array(SMALL_T["origin"]) + (idx + 0.5) * 0.1
    pc = w @ Ri.T + ti
    z = pc[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        ur = pc[..., 0] * (570.3 / 4) / z + 80.5
        vr = pc[..., 1] * (570.3 / 4) / z + 60.5
    inside = (z > 0) & (ur >= 0) & (ur < 160) & (vr >= 0) & (vr < 120)
    ui, vi = np.where(inside, ur, 0).astype(int), np.where(inside, vr, 0).astype(int)
    D = d0[vi, ui]
    sdf = D / 1000.0 - z
    upd = inside & (D > 0) & (sdf >= -0.3)
    tsdf = np.minimum(1.0, sdf / 0.3)
    b = before.transpose(2, 1, 0, 3).astype(np.float64)  # -> [x][y][z]
    a = vol.transpose(2, 1, 0, 3)
    want = (b[..., 0] / 32767.0 * b[..., 1] + tsdf) / (b[..., 1] + 1.0)
    # the float64 restatement may flip a gate or a rounding on a knife edge: allow a handful of voxels
    gate_diff = (a[..., 1] != np.where(upd, np.minimum(b[..., 1] + 1, 64), b[..., 1])).sum()
    assert gate_diff <= 40, gate_diff
    same = upd & (a[..., 1] == b[..., 1] + 1)
    assert np.abs(a[..., 0][same] - np.rint(want[same] * 32767.0)).max() <= 2
    assert (a[~upd & (a[..., 1] == b[..., 1])] == before.transpose(2, 1, 0, 3)[~upd & (a[..., 1] == b[..., 1])]).all()
    assert upd.sum() > 5000 and f32(1) == 1


def test_model_tracker_drifts_less_than_frame_to_frame(pkg, oracle):
    n = 30
    fr = pkg.synth_sequence(n, 160, 120)
    gt = pkg.synth_gt(n, 160, 120)
    cfg, t = small_cfgs(oracle)
    pm, st = oracle.track_sequence_model(cfg, t, fr)
    pf, _, _ = oracle.track_sequence(cfg, fr)
    assert st[0] == 1 and (st[1:] == 0).all()
    assert np.array_equal(pm[0], IDENT)
    em = np.linalg.norm(pm.reshape(-1, 3, 4)[:, :, 3] - gt.reshape(-1, 3, 4)[:, :, 3], axis=1)
    ef = np.linalg.norm(pf.reshape(-1, 3, 4)[:, :, 3] - gt.reshape(-1, 3, 4)[:, :, 3], axis=1)
    assert em.max() < 0.02 and em[-1] < ef[-1] * 1.5  # coarse 10 cm voxels at 160x120 already hold the trajectory


def test_map_size_counts_surface_voxels(oracle):
    cfg, t = small_cfgs(oracle, bilateral=0)
    vol = oracle.tsdf_new(t)
    assert oracle.tsdf_surface_voxels(t, vol) == 0
    oracle.tsdf_integrate(cfg, t, vol, np.full((120, 160), 1000.0, dtype=np.float32), IDENT)
    n = oracle.tsdf_surface_voxels(t, vol)
    # plane z = 1 m: the sign changes between the voxel layers centred at z = 0.95 and 1.05; count the observed columns
    layer = vol[21]  # iz = 21 <-> z = 0.95
    assert (layer[..., 1] > 0).any() and (layer[..., 0][layer[..., 1] > 0] >= 0).all()
    both = (vol[21][..., 1] > 0) & (vol[22][..., 1] > 0)
    assert n >= int(both.sum()) > 50
    # numpy restatement of the definition
    tsdf, w = vol[..., 0].astype(int), vol[..., 1]
    hit = np.zeros(w.shape, dtype=bool)
    for ax in range(3):
        a = [slice(None)] * 3
        b = [slice(None)] * 3
        a[ax], b[ax] = slice(0, -1), slice(1, None)
        a, b = tuple(a), tuple(b)
        hit[a] |= (w[a] > 0) & (w[b] > 0) & ((tsdf[a] < 0) != (tsdf[b] < 0))
    assert n == int(hit.sum())


def test_model_symbols_exported(pkg):
    from test_cabi import declared_functions

    names = declared_functions("youth_model.h")
    assert "youth_cuda_enable_model" in names and len(names) >= 7
    lib = C.CDLL(pkg.lib_paths()["cuda"])
    for nme in names:
        assert hasattr(lib, nme), f"libyouth_cuda.so does not export {nme}"


# --------------------------------------------------------------------------- GPU


def make_model_tracker(pkg, oracle, small=True, **kw):
    from slam_rgbd_b200 import binding as B

    cfg = pkg.default_config(**(SMALL if small else {}), **kw)
    tcfg = pkg.tsdf_config(**SMALL_T) if small else pkg.tsdf_config()
    trk = B.Tracker(cfg)
    trk.enable_model(tcfg)
    return trk, oracle.config_from(cfg), oracle.tsdf_config_from(tcfg)


@pytest.mark.gpu
def test_fusion_and_raycast_kernels_bit_exact(pkg, oracle):
    from slam_rgbd_b200 import binding as B

    trk, ocfg, otcfg = make_model_tracker(pkg, oracle, batch=4, traj_capacity=16)
    frames = pkg.synth_sequence(3, 160, 120)
    trk.track_batch([frames])  # makes the frames resident (and tracks them)
    trk.sync()
    # standalone kernels on a fresh volume with hand-picked poses
    trk.reset()  # clears the volume; the ring still holds the frames?  no: reset forgets them
    trk.track_batch([frames])
    vol_tracked = trk.read_volume()
    assert (vol_tracked[..., 1] > 0).sum() > 10000
    poses = [IDENT, pose_of(0.01, 0.02, -0.01, (0.02, 0.0, -0.01)), pose_of(-0.03, 0.01, 0.02, (-0.04, 0.03, 0.05))]
    ovol = vol_tracked.copy()
    for i, p in enumerate(poses):
        trk.debug_integrate(i, p)
        oracle.tsdf_integrate(ocfg, otcfg, ovol, oracle.OFrame(ocfg, frames[i]).depth(0), p)
        assert np.array_equal(trk.read_volume(), ovol), f"fused volume differs after frame {i}"
    cases = [(p, -1) for p in poses + [pose_of(0.1, -0.2, 0.05, (0.3, -0.1, 0.4))]] + [(poses[1], 1), (poses[2], 2)]
    for p, hint_frame in cases:  # hint_frame >= 0: rays start in front of the depth that frame measured
        trk.debug_raycast(p, hint_frame=hint_frame)
        hf = oracle.OFrame(ocfg, frames[hint_frame]) if hint_frame >= 0 else None
        for level in range(3):
            vm, nm = oracle.tsdf_raycast(ocfg, otcfg, ovol, p, level, hint=hf.depth(level) if hf else None)
            dv, dn = trk.read_model(B.DBG_VERTEX, level), trk.read_model(B.DBG_NORMAL, level)
            assert np.array_equal(dv.view(np.uint32), vm.view(np.uint32)), f"model vertex map differs at level {level}"
            assert np.array_equal(dn.view(np.uint32), nm.view(np.uint32)), f"model normal map differs at level {level}"
            assert (vm[..., 3] > 0).mean() > 0.25
    trk.close()


@pytest.mark.gpu
def test_model_trajectory_bit_exact_small(pkg, oracle):
    trk, ocfg, otcfg = make_model_tracker(pkg, oracle, batch=5, traj_capacity=32)
    frames = pkg.synth_sequence(12, 160, 120, noise=1)
    got = np.concatenate([trk.track_batch([frames[:5]])[0], trk.track_batch([frames[5:10]])[0],
                          trk.track_batch([frames[10:]])[0]])
    want, st = oracle.track_sequence_model(ocfg, otcfg, frames)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    _, _, dst = trk.trajectory()
    assert np.array_equal(dst, st)
    # tolerances of north_star, stated: 1e-4 rad / 1e-4 m (bit equality implies them)
    assert np.abs(got - want).max() <= 1e-4
    # frame-by-frame calls give the same chain (grouping never changes results)
    trk.reset()
    one = np.stack([trk.track(f) for f in frames])
    assert np.array_equal(one.view(np.uint32), want.view(np.uint32))
    # the size of the map (observed voxels the surface passes through) equals the statement's count
    assert trk.surface_voxels() == oracle.tsdf_surface_voxels(otcfg, trk.read_volume()) > 1000
    trk.close()


@pytest.mark.gpu
def test_model_survives_empty_and_partly_blind_frames(pkg, oracle):
    """a frame without readings is flagged LOST, keeps the pose, is NOT fused; a half-blind frame still tracks;
    the chain stays bit-identical to the CPU statement throughout"""
    from slam_rgbd_b200 import binding as B

    trk, ocfg, otcfg = make_model_tracker(pkg, oracle, batch=4, traj_capacity=16)
    frames = pkg.synth_sequence(10, 160, 120).copy()
    frames[4] = 0                 # sensor drop-out
    frames[7][:, 80:] = 0         # right half blind
    frames[8][:60] = 65535        # top half out of range (> depth_max)
    got = np.concatenate([trk.track_batch([frames[a:a + 4]])[0] for a in range(0, 10, 4)])
    want, st = oracle.track_sequence_model(ocfg, otcfg, frames)
    _, _, dst = trk.trajectory()
    assert np.array_equal(dst, st) and st[4] == B.STATUS_LOST and (st[[1, 2, 3, 5, 6, 7, 9]] == 0).all()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got[4], got[3])  # pose kept
    # the volume after the run equals the statement's (so the empty frame was not fused on either side)
    vol = oracle.tsdf_new(otcfg)
    for i in range(10):
        if st[i] & B.STATUS_LOST:
            continue
        fr = oracle.OFrame(ocfg, frames[i])
        oracle.tsdf_integrate(ocfg, otcfg, vol, fr.depth(0), want[i])
    assert np.array_equal(trk.read_volume(), vol)
    trk.close()


@pytest.mark.gpu
def test_model_trajectory_bit_exact_full_size(pkg, oracle):
    trk, ocfg, otcfg = make_model_tracker(pkg, oracle, small=False, batch=4, traj_capacity=16)
    frames = pkg.synth_sequence(4)
    got = trk.track_batch([frames])[0]
    want, _ = oracle.track_sequence_model(ocfg, otcfg, frames)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert trk.last_inliers() > 100000
    trk.close()


@pytest.mark.gpu
def test_model_mode_properties_full_size(pkg, oracle):
    """100 frames at 640x480 with the default volume: deterministic, independent of grouping and of the
    number of co-tracked sequences, resettable, and more accurate than frame-to-frame."""
    from slam_rgbd_b200 import binding as B

    n = 100
    seqs = [pkg.synth_sequence(n, sequence=s) for s in range(2)]
    gt = pkg.synth_gt(n)
    tcfg = pkg.tsdf_config()
    two = B.Tracker(pkg.default_config(batch=25, n_streams=2, traj_capacity=n))
    two.enable_model(tcfg)
    for a in range(0, n, 25):
        two.track_batch([s[a:a + 25] for s in seqs], want_poses=False)
    p0, _, st0 = two.trajectory(0)
    p1, _, _ = two.trajectory(1)
    assert (st0[1:] == 0).all()
    two.reset()
    for a in range(0, n, 20):
        two.track_batch([s[a:a + 20] for s in seqs], want_poses=False)
    q0, _, _ = two.trajectory(0)
    assert np.array_equal(p0.view(np.uint32), q0.view(np.uint32))  # reset + other grouping: same bits
    two.close()
    one = B.Tracker(pkg.default_config(batch=50, n_streams=1, traj_capacity=n))
    one.enable_model(tcfg)
    for a in range(0, n, 50):
        one.track_batch([seqs[0][a:a + 50]], want_poses=False)
    r0, _, _ = one.trajectory(0)
    assert np.array_equal(p0.view(np.uint32), r0.view(np.uint32))  # co-tracked sequences do not interact
    assert not np.array_equal(p0, p1)
    one.close()
    f2f = B.Tracker(pkg.default_config(batch=n, traj_capacity=n))
    f2f.track_batch([seqs[0]], want_poses=False)
    pf, _, _ = f2f.trajectory(0)
    f2f.close()
    em = np.linalg.norm(p0.reshape(-1, 3, 4)[:, :, 3] - gt.reshape(-1, 3, 4)[:, :, 3], axis=1)
    ef = np.linalg.norm(pf.reshape(-1, 3, 4)[:, :, 3] - gt.reshape(-1, 3, 4)[:, :, 3], axis=1)
    assert em.max() < 0.01 and em[-1] < ef[-1]
    # YD16-packed input feeds the same chain (unpacked on the device)
    cd = pkg.Codec(640, 480, max_frames=30)
    packed, offs = cd.encode(seqs[0][:30])
    cd.close()
    t2 = B.Tracker(pkg.default_config(batch=20, traj_capacity=32))
    t2.enable_model(tcfg)
    a = t2.track_batch_packed([packed], [offs[:21].copy()], 20)[0]
    b = t2.track_batch_packed([packed], [offs[20:].copy()], 10)[0]
    t2.close()
    assert np.array_equal(np.concatenate([a, b]).view(np.uint32), p0[:30].view(np.uint32))
