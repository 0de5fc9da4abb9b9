"""GPU parity: libyouth_cuda.so (through its C ABI) against the CPU oracle on the same
seeded inputs.  Tolerances from BASELINE.json north_star: masks / pyramid counts /
correspondence indices bit-exact, vertex+normal maps <= 1e-5 relative, poses <= 1e-4 rad
and <= 1e-4 m.  Both sides use the same FMA-free operation sequence, so the tests first
assert bit-equality and only report the tolerance-level comparison when that fails."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_MAP_REL = 1e-5
TOL_POSE_RAD = 1e-4
TOL_POSE_M = 1e-4


def rot_angle(Ra, Rb):
    """Angle between two rotations; chordal form, well conditioned near zero
    (|Ra - Rb|_F = 2*sqrt(2)*sin(angle/2)), unlike arccos of the trace."""
    return float(2.0 * np.arcsin(min(1.0, np.linalg.norm(Ra - Rb) / (2.0 * np.sqrt(2.0)))))


def assert_pose_close(pa, pb):
    A, B = np.asarray(pa, dtype=np.float64).reshape(3, 4), np.asarray(pb, dtype=np.float64).reshape(3, 4)
    assert np.linalg.norm(A[:, 3] - B[:, 3]) <= TOL_POSE_M
    assert rot_angle(A[:, :3], B[:, :3]) <= TOL_POSE_RAD


def assert_map_close(g, o):
    if np.array_equal(g, o):
        return
    scale = np.maximum(np.abs(o), 1e-3)
    assert np.max(np.abs(g - o) / scale) <= TOL_MAP_REL


def make_tracker(pkg, **kw):
    from slam_rgbd_b200.binding import Tracker

    return Tracker(pkg.default_config(**kw))


@pytest.mark.parametrize("bilateral", [1, 0])
def test_preprocess_parity(pkg, oracle, small_seq, bilateral):
    """stage 1+2: depth pyramid, pyramid sample counts, masks, vertex and normal maps."""
    from slam_rgbd_b200 import binding as B

    frames, _ = small_seq
    trk = make_tracker(pkg, bilateral=bilateral, batch=4)
    with pytest.raises(Exception):
        trk.debug_read(B.DBG_DEPTH, 0, 0)  # nothing tracked yet
    trk.enable_debug_maps()  # the product instantiation does not store the float depth pyramid / sample counts
    ocfg = oracle.config_from(trk.cfg)
    trk.track_batch([frames[:3]])
    for fi in range(3):
        of = oracle.OFrame(ocfg, frames[fi])
        for level in range(trk.cfg.levels):
            d = trk.debug_read(B.DBG_DEPTH, fi, level)
            assert np.array_equal(d, of.depth(level)), f"depth frame {fi} level {level}"
            m = trk.debug_read(B.DBG_MASK, fi, level)
            assert np.array_equal(m, of.mask(level)), f"mask frame {fi} level {level}"
            if level > 0:
                c = trk.debug_read(B.DBG_PYRCNT, fi, level)
                assert np.array_equal(c, of.pyrcnt(level)), f"pyramid count frame {fi} level {level}"
            assert_map_close(trk.debug_read(B.DBG_VERTEX, fi, level), of.vmap(level))
            assert_map_close(trk.debug_read(B.DBG_NORMAL, fi, level), of.nmap(level))
    trk.close()


def test_preprocess_bit_exact(pkg, oracle, small_seq):
    """The FMA-free contract makes the maps bit-identical, not merely within 1e-5."""
    from slam_rgbd_b200 import binding as B

    frames, _ = small_seq
    trk = make_tracker(pkg, batch=2)
    ocfg = oracle.config_from(trk.cfg)
    trk.track_batch([frames[:2]])
    of = oracle.OFrame(ocfg, frames[1])
    for level in range(trk.cfg.levels):
        assert np.array_equal(trk.debug_read(B.DBG_VERTEX, 1, level).view(np.uint32), of.vmap(level).view(np.uint32))
        assert np.array_equal(trk.debug_read(B.DBG_NORMAL, 1, level).view(np.uint32), of.nmap(level).view(np.uint32))
    trk.close()


@pytest.mark.parametrize("ppt", [1, 16, 64, 256])
def test_icp_sums_and_correspondences(pkg, oracle, small_seq, ppt):
    """stage 3+4: correspondence indices (and reject codes) bit-exact; 29 sums bit-exact."""
    frames, gt = small_seq
    trk = make_tracker(pkg, batch=2, icp_ppt=ppt)
    ocfg = oracle.config_from(trk.cfg)
    trk.track_batch([frames[:2]])
    prev, cur = oracle.OFrame(ocfg, frames[0]), oracle.OFrame(ocfg, frames[1])
    poses = [np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32), gt[1].astype(np.float32)]
    # a deliberately wrong pose exercises the distance / angle / out-of-image rejections
    bad = poses[0].copy()
    bad[3], bad[7] = 0.06, -0.04
    poses.append(bad)
    for level in range(trk.cfg.levels):
        for pose in poses:
            gs, gc = trk.debug_icp(1, level, pose)
            os_, oc = oracle.icp_sums(ocfg, level, cur, prev, pose)
            assert np.array_equal(gc, oc), f"correspondence map level {level}"
            assert gs[31] == os_[31]
            assert np.array_equal(gs.view(np.uint64), os_.view(np.uint64)), f"sums level {level}: {gs - os_}"
    trk.close()


@pytest.mark.parametrize("ppt", [64, 128])
def test_track_sequence_pose_parity(pkg, oracle, small_seq, ppt):
    """stages 1-5 end to end over 6 frames: poses vs oracle within 1e-4 m / 1e-4 rad (and,
    by construction, bit-identical), inlier count identical.  icp_ppt 64 = library default, 128 = what
    bench.py uses for launches of many pairs."""
    frames, gt = small_seq
    trk = make_tracker(pkg, batch=4, icp_ppt=ppt)
    ocfg = oracle.config_from(trk.cfg)
    poses_g = np.concatenate([trk.track_batch([frames[:4]])[0], trk.track_batch([frames[4:6]])[0]])
    poses_o, st_o, _ = oracle.track_sequence(ocfg, frames)
    traj, ts, st = trk.trajectory()
    assert np.array_equal(traj, poses_g)
    assert np.array_equal(st, st_o)
    for i in range(6):
        assert_pose_close(poses_g[i], poses_o[i])
    assert np.array_equal(poses_g.view(np.uint32), poses_o.view(np.uint32)), "poses not bit-identical"
    # and the estimate is right: within 2 mm / 0.05 deg of the synthetic ground truth
    for i in range(6):
        G = gt[i].reshape(3, 4)
        P = poses_g[i].astype(np.float64).reshape(3, 4)
        assert np.linalg.norm(G[:, 3] - P[:, 3]) < 2e-3
        assert rot_angle(G[:, :3], P[:, :3]) < np.radians(0.05)
    trk.close()


def test_streaming_equals_batched(pkg, small_seq):
    """youth_cuda_track one frame at a time == youth_cuda_track_batch (pairs are independent)."""
    frames, _ = small_seq
    a = make_tracker(pkg, batch=1)
    pa = np.stack([a.track(frames[i], ts=33 * i) for i in range(5)])
    b = make_tracker(pkg, batch=5)
    pb = b.track_batch([frames[:5]])[0]
    assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32))
    assert a.last_inliers() == b.last_inliers() > 100000
    a.close()
    b.close()


def test_determinism(pkg, small_seq):
    """same input twice -> bitwise identical sums and poses (fixed-order reduction, no atomics)."""
    frames, _ = small_seq
    out = []
    for _ in range(2):
        t = make_tracker(pkg, batch=3)
        p = t.track_batch([frames[:3]])[0]
        s, _c = t.debug_icp(2, 0, np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32), want_corr=False)
        out.append((p.copy(), s.copy()))
        t.close()
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    assert np.array_equal(out[0][1].view(np.uint64), out[1][1].view(np.uint64))


def test_multi_stream_matches_single(pkg):
    """two sequences tracked in one handle == each tracked alone."""
    fa, fb = pkg.synth_sequence(4, sequence=0), pkg.synth_sequence(4, sequence=5)
    both = make_tracker(pkg, n_streams=2, batch=4)
    pboth = both.track_batch([fa, fb])
    for k, f in enumerate((fa, fb)):
        one = make_tracker(pkg, batch=4)
        p = one.track_batch([f])[0]
        assert np.array_equal(p.view(np.uint32), pboth[k].view(np.uint32))
        one.close()
    both.close()


def test_edge_cases(pkg, oracle):
    """empty (all-invalid) frames, a frame with no overlap, reset, and capacity errors."""
    from slam_rgbd_b200 import binding as B

    trk = make_tracker(pkg, batch=2, traj_capacity=5)
    ocfg = oracle.config_from(trk.cfg)
    H, W = trk.cfg.height, trk.cfg.width
    empty = np.zeros((2, H, W), dtype=np.uint16)
    p = trk.track_batch([empty])[0]
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    assert np.array_equal(p[0], ident) and np.array_equal(p[1], ident)
    _, _, st = trk.trajectory()
    assert st[0] == B.STATUS_FIRST and st[1] == B.STATUS_LOST  # no inliers -> pose kept, flagged
    po, so, _ = oracle.track_sequence(ocfg, empty)
    assert np.array_equal(po, p) and np.array_equal(so, st)
    assert trk.last_inliers() == 0
    # maximum raw values are outside the depth gate -> invalid everywhere
    full = np.full((2, H, W), 65535, dtype=np.uint16)
    trk.track_batch([full])
    assert not trk.debug_read(B.DBG_MASK, 3, 0).any()
    with pytest.raises(RuntimeError):  # capacity 5: frames 4,5 do not fit
        trk.track_batch([empty])
    trk.reset()
    assert trk.frame_count() == 0
    assert np.array_equal(trk.track_batch([empty[:1]])[0][0], ident)
    trk.close()


RANDOM_CONFIGS = [
    # width, height, levels, extra config, noise
    (96, 64, 1, dict(iters=[6, 0, 0, 0], sigma_range_mm=10.0, sigma_space_px=2.0), 0),
    (128, 96, 2, dict(iters=[5, 3, 0, 0], depth_factor=5000.0, depth_max_mm=40000, bilateral=0), 1),
    (200, 152, 3, dict(iters=[4, 3, 2, 0], dist_thresh_m=0.03, cos_thresh=0.98, icp_ppt=2), 1),
    (320, 240, 4, dict(iters=[3, 2, 2, 2], sigma_range_mm=60.0, icp_ppt=128, depth_min_mm=700), 0),
    (72, 40, 2, dict(iters=[4, 4, 0, 0], min_inliers=100000), 0),  # never enough inliers: every pair flagged lost
    (640, 480, 3, dict(iters=[2, 1, 1, 0], fx=525.0, fy=531.5, cx=311.25, cy=247.75, icp_ppt=32), 1),
]


@pytest.mark.parametrize("w,h,levels,extra,noise", RANDOM_CONFIGS)
def test_parity_across_configurations(pkg, oracle, w, h, levels, extra, noise):
    """sizes with partial ingest tiles, 1-4 levels, other intrinsics / depth scales / gates /
    reduction geometries: depth pyramid, masks, maps, poses and status all bit-identical."""
    from slam_rgbd_b200 import binding as B

    scale = w / 640.0
    base = dict(width=w, height=h, levels=levels, fx=570.3 * scale, fy=570.3 * scale, cx=w / 2.0, cy=h / 2.0, batch=3)
    base.update(extra)
    frames = pkg.synth_sequence(3, w, h, sequence=7, noise=noise)
    if base.get("depth_factor", 1000.0) != 1000.0:
        frames = (frames.astype(np.uint32) * 5).astype(np.uint16)  # 0.2 mm units
    trk = make_tracker(pkg, **base)
    ocfg = oracle.config_from(trk.cfg)
    poses = trk.track_batch([frames])[0]
    want, st_o, _ = oracle.track_sequence(ocfg, frames)
    _, _, st = trk.trajectory()
    assert np.array_equal(st, st_o)
    assert np.array_equal(poses.view(np.uint32), want.view(np.uint32))
    of = oracle.OFrame(ocfg, frames[2])
    with pytest.raises(Exception):  # the product path does not store the depth pyramid: the read-back refuses
        trk.debug_read(B.DBG_DEPTH, 2, 0)
    dbg = make_tracker(pkg, **base)
    dbg.enable_debug_maps()
    assert np.array_equal(dbg.track_batch([frames])[0].view(np.uint32), want.view(np.uint32))
    for level in range(levels):
        assert np.array_equal(dbg.debug_read(B.DBG_DEPTH, 2, level), of.depth(level))
        assert np.array_equal(dbg.debug_read(B.DBG_VERTEX, 2, level).view(np.uint32), of.vmap(level).view(np.uint32))
    dbg.close()
    for level in range(levels):
        assert np.array_equal(trk.debug_read(B.DBG_MASK, 2, level), of.mask(level))
        assert np.array_equal(trk.debug_read(B.DBG_VERTEX, 2, level).view(np.uint32), of.vmap(level).view(np.uint32))
        assert np.array_equal(trk.debug_read(B.DBG_NORMAL, 2, level).view(np.uint32), of.nmap(level).view(np.uint32))
    if extra.get("min_inliers", 0) > 50000:
        assert list(st) == [1, 2, 2]
    trk.close()


def test_independent_handles_from_two_threads(pkg, small_seq):
    """handles are independent (include/youth_cuda.h, Threading): two trackers driven concurrently from two
    host threads give the bits of a tracker driven alone"""
    import threading

    from slam_rgbd_b200 import binding as B

    frames, _ = small_seq
    alone = B.Tracker(pkg.default_config(batch=6))
    want = alone.track_batch([frames])[0]
    alone.close()
    out = [None, None]

    errors = [None, None]

    def work(k):
        try:
            t = B.Tracker(pkg.default_config(batch=3 if k else 2))
            n = 3 if k else 2
            res = []
            for rep in range(4):
                t.reset()
                res = [t.track_batch([frames[a:a + n]])[0] for a in range(0, 6, n)]
            out[k] = np.concatenate(res)
            t.close()
        except Exception as e:  # surfaced below: an exception in a thread would otherwise only be a warning
            errors[k] = repr(e)

    ths = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for th in ths:
        th.start()
    for th in ths:
        th.join(timeout=120)
    assert all(not th.is_alive() for th in ths)
    assert errors == [None, None], errors
    for k in range(2):
        assert np.array_equal(out[k].view(np.uint32), want.view(np.uint32)), f"thread {k}"


def test_reciprocal_is_correctly_rounded_over_all_normals(pkg):
    """Stage 3's written-out reciprocal (approximation + one Newton step, no range test) against the
    IEEE reciprocal for EVERY positive normal float below 2^126 -- the only inputs whose quotient the
    front gate (v'.z >= FLT_MIN; |v'| < 2^21 for finite poses) lets through."""
    trk = make_tracker(pkg, batch=1, width=64, height=32, levels=1)
    lo = 0x00800000  # FLT_MIN
    hi = 0x7E800000 - 1  # just below 2^126
    assert trk.debug_rcp_check(lo, hi) == 0
    # the hook itself detects a difference where one exists (denormal inputs take the IEEE slow path)
    assert trk.debug_rcp_check(0x00000001, 0x000FFFFF) > 0
    trk.close()


def test_icp_schedule_does_not_change_a_bit(pkg, small_seq):
    """youth_cuda_set_icp_schedule: pair groups on several queues give the trajectory of the plain
    one-launch-per-iteration schedule, bit for bit (pairs are independent; the reduction order of a
    pair does not depend on which launch carries it)."""
    frames, _ = small_seq
    n = min(len(frames), 12)
    ref = None
    for group, queues in [(0, 1), (1, 1), (3, 2), (5, 4), (4, 8)]:
        trk = make_tracker(pkg, batch=n, traj_capacity=n)
        trk.set_icp_schedule(group, queues)
        poses = trk.track_batch([frames[:n]])[0]
        trk.close()
        if ref is None:
            ref = poses
        assert np.array_equal(poses.view(np.uint32), ref.view(np.uint32)), (group, queues)


@pytest.mark.parametrize("sigma_range_mm", [30.0, 41.9, 42.5, 60.0])
def test_bilateral_table_and_generic_paths_match_the_oracle(pkg, oracle, small_seq, sigma_range_mm, monkeypatch):
    """k_ingest has two bilateral variants: the product table s_wt[class][|diff|] (3 sigma_range + 2 <= 128
    entries, i.e. up to 42 mm) and the generic per-tap product (any sigma_range).  Both must give the
    oracle's filtered depth bit for bit; the table variant is also compared with the generic one forced
    through YOUTH_INGEST_GENERIC=1."""
    from slam_rgbd_b200 import binding as B

    frames, _ = small_seq
    out = {}
    for forced in ("0", "1"):
        monkeypatch.setenv("YOUTH_INGEST_GENERIC", forced)
        trk = make_tracker(pkg, batch=2, sigma_range_mm=sigma_range_mm)
        trk.enable_debug_maps()
        ocfg = oracle.config_from(trk.cfg)
        trk.track_batch([frames[:2]])
        of = oracle.OFrame(ocfg, frames[1])
        for level in range(trk.cfg.levels):
            d = trk.debug_read(B.DBG_DEPTH, 1, level)
            assert np.array_equal(d, of.depth(level)), f"depth level {level} (generic forced: {forced})"
            out[(forced, level)] = d
        trk.close()
    for level in range(3):
        assert np.array_equal(out[("0", level)], out[("1", level)])


def test_two_groups_in_flight_give_the_blocking_results(pkg, small_seq):
    """youth_cuda_read_trajectory_async / youth_cuda_wait_ticket: submit group g+1 before collecting group g
    (host-fed, pinned); every collected trajectory equals the blocking call's, bit for bit, and tickets
    complete in order."""
    import ctypes as C

    from slam_rgbd_b200 import binding as B

    frames, _ = small_seq
    n = len(frames)
    trk = make_tracker(pkg, batch=n, traj_capacity=n)
    ref = trk.track_batch([frames])[0].copy()
    pin = trk.lib.youth_cuda_host_alloc(frames.nbytes)
    C.memmove(pin, frames.ctypes.data, frames.nbytes)
    res = [trk.lib.youth_cuda_host_alloc(n * 48) for _ in range(2)]
    views = [np.ctypeslib.as_array((C.c_float * (n * 12)).from_address(p)).reshape(n, 12) for p in res]
    pending, seen = [], []
    for i in range(5):
        trk.reset()
        trk.track_batch_ptrs([pin], n, B.MEM_HOST_PINNED)
        views[i % 2][...] = -1.0  # step i-2 has been collected: poison the buffer so that the read-back must land
        got, ticket = trk.read_trajectory_async(res[i % 2], n)
        assert got == n
        pending.append((i, ticket))
        if len(pending) > 1:
            j, t = pending.pop(0)
            trk.wait_ticket(t)
            seen.append(t)
            assert np.array_equal(views[j % 2].view(np.uint32), ref.view(np.uint32)), f"step {j}"
    j, t = pending.pop(0)
    trk.wait_ticket(t)
    seen.append(t)
    assert np.array_equal(views[j % 2].view(np.uint32), ref.view(np.uint32))
    assert seen == sorted(seen)
    trk.wait_ticket(seen[0])  # waiting again on a finished ticket returns at once
    for p in res + [pin]:
        trk.lib.youth_cuda_host_free(p)
    trk.close()


def test_reciprocal_form_divisions_do_not_change_a_bit(pkg, oracle, small_seq, monkeypatch):
    """The back-projection divides by depth_factor, fx, fy (viewerModule.c:343-345).  The device uses a host-side
    reciprocal and two fused multiply-adds (div_cfg) after checking at init, exhaustively, that this gives the IEEE
    quotient for the configured divisors; YOUTH_NO_FAST_DIV=1 keeps the divisions.  Both must give the oracle's maps
    and poses bit for bit (also for a non-default depth scale and odd intrinsics)."""
    from slam_rgbd_b200 import binding as B

    frames, _ = small_seq
    for extra in ({}, dict(depth_factor=5000.0, fx=525.0, fy=531.7, cx=319.5, cy=239.5)):
        fr = frames if not extra else (frames.astype(np.uint32) * 5).astype(np.uint16)
        got = {}
        for no_fast in ("0", "1"):
            if no_fast == "1":
                monkeypatch.setenv("YOUTH_NO_FAST_DIV", "1")
            else:
                monkeypatch.delenv("YOUTH_NO_FAST_DIV", raising=False)
            trk = make_tracker(pkg, batch=6, **extra)
            poses = trk.track_batch([fr])[0]
            maps = [trk.debug_read(what, 5, level) for level in range(3) for what in (B.DBG_VERTEX, B.DBG_NORMAL)]
            got[no_fast] = (poses, maps)
            ocfg = oracle.config_from(trk.cfg)
            trk.close()
        want, _, _ = oracle.track_sequence(ocfg, fr)
        of = oracle.OFrame(ocfg, fr[5])
        for no_fast in ("0", "1"):
            poses, maps = got[no_fast]
            assert np.array_equal(poses.view(np.uint32), want.view(np.uint32)), (extra, no_fast)
            for level in range(3):
                assert np.array_equal(maps[2 * level].view(np.uint32), of.vmap(level).view(np.uint32)), (extra, no_fast, level)
                assert np.array_equal(maps[2 * level + 1].view(np.uint32), of.nmap(level).view(np.uint32)), (extra, no_fast, level)


def test_many_streams_two_steps_in_flight_wait_really_waits(pkg):
    """16 sequences in one handle, one asynchronous read-back per sequence and step, two steps in flight: 32 tickets
    are issued between a ticket and its wait, and a step's buffers are only read after its LAST ticket was waited
    for.  Regression test: with 16 event slots the wait on a recycled slot used to return without synchronising.
    Also waits on a ticket that is several hundred reads old (its slot was re-used many times)."""
    import ctypes as C

    from slam_rgbd_b200 import binding as B

    S, n, w, h = 16, 3, 160, 120
    small = dict(width=w, height=h, fx=570.3 / 4, fy=570.3 / 4, cx=80.0, cy=60.0)
    seqs = [pkg.synth_sequence(n, w, h, sequence=s) for s in range(S)]
    trk = make_tracker(pkg, batch=n, n_streams=S, traj_capacity=n, **small)
    ref = [p.copy() for p in trk.track_batch(seqs)]
    pins = []
    for s in range(S):
        pin = trk.lib.youth_cuda_host_alloc(seqs[s].nbytes)
        C.memmove(pin, seqs[s].ctypes.data, seqs[s].nbytes)
        pins.append(pin)
    res = [[trk.lib.youth_cuda_host_alloc(n * 48) for _ in range(S)] for _ in range(2)]
    views = [[np.ctypeslib.as_array((C.c_float * (n * 12)).from_address(p)).reshape(n, 12) for p in r] for r in res]
    pending, first_ticket = [], None
    for i in range(12):
        trk.reset()
        trk.track_batch_ptrs(pins, n, B.MEM_HOST_PINNED)
        tickets = []
        for s in range(S):
            views[i % 2][s][...] = -1.0
            got, t = trk.read_trajectory_async(res[i % 2][s], n, stream=s)
            assert got == n
            tickets.append(t)
        if first_ticket is None:
            first_ticket = tickets[0]
        pending.append((i, tickets))
        if len(pending) > 1:
            j, ts = pending.pop(0)
            trk.wait_ticket(ts[-1])  # 2 * S - 1 newer tickets exist by now
            for s in range(S):
                assert np.array_equal(views[j % 2][s].view(np.uint32), ref[s].view(np.uint32)), f"step {j} sequence {s}"
    j, ts = pending.pop(0)
    trk.wait_ticket(ts[-1])
    for s in range(S):
        assert np.array_equal(views[j % 2][s].view(np.uint32), ref[s].view(np.uint32))
    trk.wait_ticket(first_ticket)  # 191 reads old
    with pytest.raises(Exception):
        trk.wait_ticket(ts[-1] + 1)  # never issued
    for p in pins + res[0] + res[1]:
        trk.lib.youth_cuda_host_free(p)
    trk.close()


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("iters", [(10, 5, 4), (3, 0, 2), (0, 0, 3)])
def test_fused_and_per_iteration_icp_match_the_oracle(pkg, oracle, small_seq, fused, iters, monkeypatch):
    """YOUTH_ICP_FUSED=1: few pairs per launch take k_icp_fused (the whole coarse-to-fine schedule in one
    launch, the CTAs of a pair handing over through a generation counter) instead of one k_icp launch per
    iteration (the default: measured faster).  Both must give the oracle's poses and status words bit for
    bit, also when a level has no iterations, over several groups (the counters are left zero) and for two
    sequences in one launch."""
    monkeypatch.setenv("YOUTH_ICP_FUSED", fused)
    frames, _ = small_seq
    other = pkg.synth_sequence(6, sequence=3)
    trk = make_tracker(pkg, batch=2, n_streams=2, iters=list(iters) + [0])
    ocfg = oracle.config_from(trk.cfg)
    got = [[], []]
    for g in range(3):
        out = trk.track_batch([frames[2 * g:2 * g + 2], other[2 * g:2 * g + 2]])
        for s in range(2):
            got[s].append(out[s])
    for s, seq in enumerate((frames, other)):
        poses_o, st_o, inl_o = oracle.track_sequence(ocfg, seq)
        poses_g = np.concatenate(got[s])
        assert np.array_equal(poses_g.view(np.uint32), poses_o.view(np.uint32)), f"sequence {s}: poses not bit-identical"
        _, _, st = trk.trajectory(s)
        assert np.array_equal(st, st_o)
    trk.close()
