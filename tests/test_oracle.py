"""CPU oracle: known-answer tests with closed-form results, the conventions the reference
pins (depth unit, validity, pinhole back-projection: ViewerModule/viewerModule.c:341-345;
intrinsics: config/astra_orb_slam3_rgbd.yaml:9-12), independent numpy/scipy cross-checks of
the solve and the reduction, and the committed golden fixture."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

HERE = os.path.dirname(os.path.abspath(__file__))
IDENT = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)
# slot of A[i][j] (i <= j) in the 32 sums: pairs sized for fma.rn.f32x2 (include/youth_cuda.h)
SLOT_A = [[0, 1, 2, 3, 4, 5], [1, 7, 8, 9, 10, 11], [2, 8, 12, 13, 14, 15], [3, 9, 13, 17, 18, 19],
          [4, 10, 14, 18, 20, 21], [5, 11, 15, 19, 21, 23]]


def small_cfg(O, w=160, h=120, **kw):
    return O.default_config(width=w, height=h, fx=570.3 * w / 640, fy=570.3 * w / 640, cx=w / 2.0, cy=h / 2.0, **kw)


def test_defaults_match_reference_yaml(oracle):
    cfg = oracle.default_config()
    assert (cfg.width, cfg.height) == (640, 480)
    assert np.float32(cfg.fx) == np.float32(570.3) and np.float32(cfg.fy) == np.float32(570.3)
    assert (cfg.cx, cfg.cy, cfg.depth_factor) == (320.0, 240.0, 1000.0)
    assert list(cfg.iters)[:3] == [10, 5, 4] and cfg.levels == 3


def test_level_geometry(oracle):
    cfg = oracle.default_config()
    g1, g2 = oracle.level_geometry(cfg, 1), oracle.level_geometry(cfg, 2)
    assert (g1.w, g1.h, g2.w, g2.h) == (320, 240, 160, 120)
    assert g1.fx == np.float32(570.3) / 2 and g2.fx == np.float32(570.3) / 4
    assert g1.cx == (320.0 - 0.5) / 2 and g2.cx == ((320.0 - 0.5) / 2 - 0.5) / 2


def test_backprojection_follows_viewer_formula(oracle):
    """restatement of viewerModule.c:341-345 in numpy float32: valid iff depth > 0,
    z = d / 1000, x = (u - W/2) * z / 570.3, y = (v - H/2) * z / 570.3."""
    cfg = oracle.default_config(bilateral=0, levels=1)
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 6000, size=(480, 640), dtype=np.uint16)
    raw[rng.random(raw.shape) < 0.1] = 0
    f = oracle.OFrame(cfg, raw)
    v = f.vmap(0)
    u, vv = np.meshgrid(np.arange(640, dtype=np.float32), np.arange(480, dtype=np.float32))
    z = raw.astype(np.float32) / np.float32(1000.0)
    x = (u - np.float32(640 // 2)) * z / np.float32(570.3)
    y = (vv - np.float32(480 // 2)) * z / np.float32(570.3)
    valid = raw > 0
    assert np.array_equal(v[..., 3] == 1.0, valid)
    assert np.array_equal(v[..., 0][valid], x[valid])
    assert np.array_equal(v[..., 1][valid], y[valid])
    assert np.array_equal(v[..., 2][valid], z[valid])
    assert not v[~valid].any()


def test_plane_normals_and_validity(oracle):
    cfg = small_cfg(oracle)
    raw = np.full((120, 160), 1000, dtype=np.uint16)
    raw[40:43, 70:75] = 0  # a hole
    f = oracle.OFrame(cfg, raw)
    for level in range(3):
        n, v = f.nmap(level), f.vmap(level)
        ok = n[..., 3] == 1.0
        assert np.allclose(n[ok][:, :3], [0, 0, 1], atol=1e-4)  # fronto-parallel plane, forward differences
        assert not ok[-1, :].any() and not ok[:, -1].any()      # last row/column have no forward neighbour
        assert np.allclose(v[..., 2][v[..., 3] == 1.0], 1.0, rtol=1e-6)
    # a normal needs its right and lower neighbours: the hole invalidates one pixel up/left too
    m = f.mask(0)
    assert not (m[40:43, 70:75] & 1).any()
    assert not (m[40:43, 69] & 2).any() and not (m[39, 70:75] & 2).any()
    assert (m[38, 70:75] == 3).all()


def test_depth_gate_and_bilateral_properties(oracle):
    cfg = small_cfg(oracle, depth_min_mm=500, depth_max_mm=4000)
    raw = np.full((120, 160), 2000, dtype=np.uint16)
    raw[:10] = 100     # below the gate
    raw[-10:] = 60000  # above the gate
    raw[50:70, 80:] = 2500  # a 500 mm step: far beyond 3 sigma_r = 90 mm, must stay sharp
    d0 = np.empty((120, 160), dtype=np.float32)
    oracle.lib().yo_bilateral(C.byref(cfg), raw.ctypes.data, d0.ctypes.data)
    assert not d0[:10].any() and not d0[-10:].any()
    assert np.allclose(d0[10:50], 2000.0, rtol=1e-6)  # constant regions are fixed points (up to float rounding)
    assert np.allclose(d0[50:70, :80], 2000.0, rtol=1e-6) and np.allclose(d0[50:70, 80:], 2500.0, rtol=1e-6)
    # smoothing: +-2 mm noise is reduced
    rng = np.random.default_rng(0)
    noisy = (2000 + rng.integers(-2, 3, size=(120, 160))).astype(np.uint16)
    oracle.lib().yo_bilateral(C.byref(cfg), noisy.ctypes.data, d0.ctypes.data)
    assert d0[20:100, 20:140].std() < 0.5 * noisy[20:100, 20:140].astype(np.float64).std()
    cfg.bilateral = 0
    oracle.lib().yo_bilateral(C.byref(cfg), noisy.ctypes.data, d0.ctypes.data)
    assert np.array_equal(d0, noisy.astype(np.float32))


def _rn32(x):
    """a Fraction rounded to the nearest float32, ties to even, in exact arithmetic (no double rounding)"""
    from fractions import Fraction

    if x == 0:
        return np.float32(0.0)
    sign, a = (-1 if x < 0 else 1), abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length() - 24
    while a / Fraction(2) ** e >= 1 << 24:
        e += 1
    while e > -149 and a / Fraction(2) ** e < 1 << 23:
        e -= 1
    e = max(e, -149)  # subnormals: multiples of 2^-149
    return np.float32(sign * float(round(a / Fraction(2) ** e) * Fraction(2) ** e))  # round(): half to even; exact in double


def test_bilateral_against_exact_rational_restatement(oracle):
    """Stage 1 restated independently of the C oracle, tap by tap in exact rational arithmetic with one explicit
    float32 rounding per written operation (product of the two table weights, running weight sum, FUSED
    multiply-add of the numerator, final division) -- pure Python, so a small frame:
    holes, values outside the depth gate, a step beyond 3 sigma_r and all four borders are in it.  Bit-equal."""
    from fractions import Fraction as Fr

    w, h = 24, 16
    cfg = small_cfg(oracle, w=w, h=h, depth_min_mm=400, depth_max_mm=5000)
    rng = np.random.default_rng(11)
    raw = (1500 + rng.integers(-25, 26, size=(h, w))).astype(np.uint16)
    raw[:, 14:] += 400                      # a step far beyond the 90 mm range cut
    raw[rng.random((h, w)) < 0.06] = 0      # holes
    raw[3, 5], raw[9, 20] = 300, 6000       # outside the depth gate
    got = np.empty((h, w), dtype=np.float32)
    oracle.lib().yo_bilateral(C.byref(cfg), raw.ctypes.data, got.ctypes.data)
    ss, sr = float(cfg.sigma_space_px), float(cfg.sigma_range_mm)
    cut = int(np.float32(3.0) * np.float32(sr))
    ws = [[np.float32(np.exp(-float(dx * dx + dy * dy) / (2.0 * ss * ss))) for dx in range(-3, 4)] for dy in range(-3, 4)]
    wr = [np.float32(np.exp(-float(i) * float(i) / (2.0 * sr * sr))) for i in range(cut + 1)]
    valid = lambda d: 400 <= d <= 5000
    want = np.zeros((h, w), dtype=np.float32)
    for y in range(h):
        for x in range(w):
            dc = int(raw[y, x])
            if not valid(dc):
                continue
            sw, swd = np.float32(0), np.float32(0)
            for dy in range(-3, 4):
                for dx in range(-3, 4):
                    yy, xx = y + dy, x + dx
                    if not (0 <= yy < h and 0 <= xx < w):
                        continue
                    dk = int(raw[yy, xx])
                    if not valid(dk) or abs(dk - dc) > cut:
                        continue
                    wgt = _rn32(Fr(float(ws[dy + 3][dx + 3])) * Fr(float(wr[abs(dk - dc)])))
                    sw = _rn32(Fr(float(sw)) + Fr(float(wgt)))
                    swd = _rn32(Fr(float(wgt)) * dk + Fr(float(swd)))  # fma: one rounding
            want[y, x] = _rn32(Fr(float(swd)) / Fr(float(sw)))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (got > 0).sum() > 300 and np.abs(got[got > 0] - raw[got > 0]).max() > 1.0  # it did filter


def test_pyrdown_rule(oracle):
    cfg = small_cfg(oracle)
    src = np.zeros((4, 8), dtype=np.float32)
    src[0:2, 0:2] = [[1000, 1010], [1020, 1030]]   # all within 90 mm -> mean of 4
    src[0:2, 2:4] = [[0, 1000], [1200, 1010]]      # first valid is 1000; 1200 is outside the gate
    src[0:2, 4:6] = 0                               # nothing valid
    src[0:2, 6:8] = [[0, 0], [0, 777]]             # single sample
    dst = np.empty((2, 4), dtype=np.float32)
    cnt = np.empty((2, 4), dtype=np.uint8)
    oracle.lib().yo_pyrdown(C.byref(cfg), 8, 4, src.ctypes.data, dst.ctypes.data, cnt.ctypes.data)
    assert list(cnt[0]) == [4, 2, 0, 1]
    assert list(dst[0]) == [1015.0, 1005.0, 0.0, 777.0]
    assert not dst[1].any() and not cnt[1].any()


def test_identity_motion(oracle, pkg):
    """same frame as current and previous: every valid pixel matches itself, residuals are
    exactly zero, J^T r = 0, and the pose stays the identity."""
    cfg = small_cfg(oracle)
    raw = pkg.synth_sequence(1, 160, 120, sequence=1)[0]
    f = oracle.OFrame(cfg, raw)
    for level in range(3):
        sums, corr = oracle.icp_sums(cfg, level, f, f, IDENT)
        h, w = corr.shape
        idx = np.arange(h * w).reshape(h, w)
        both = f.mask(level) == 3
        assert np.array_equal(corr[both], idx[both])
        assert np.all(corr[~both] == -1)
        assert sums[31] == both.sum()
        assert not sums[24:31].any()
    rel, status, inl = oracle.track_pair(cfg, f, f)
    assert status == 0 and inl == (f.mask(0) == 3).sum()
    assert np.array_equal(rel, IDENT.astype(np.float64))


def test_known_motion_recovery(oracle, small_seq):
    frames, gt = small_seq
    cfg = oracle.default_config()
    poses, status, _ = oracle.track_sequence(cfg, frames[:3])
    assert list(status) == [1, 0, 0]
    for i in range(3):
        P, G = poses[i].astype(np.float64).reshape(3, 4), gt[i].reshape(3, 4)
        assert np.linalg.norm(P[:, 3] - G[:, 3]) < 1e-3
        assert np.linalg.norm(P[:, :3] - G[:, :3]) < 1e-3


def test_reduction_matches_float64_sum(oracle, small_seq):
    """the fixed-order float tree must agree with a plain float64 sum of the per-pixel terms."""
    frames, gt = small_seq
    cfg = oracle.default_config()
    prev, cur = oracle.OFrame(cfg, frames[0]), oracle.OFrame(cfg, frames[1])
    pose = gt[1].astype(np.float32)
    sums, corr = oracle.icp_sums(cfg, 0, cur, prev, pose)
    ok = corr >= 0
    P = pose.reshape(3, 4)
    vc = cur.vmap(0)[ok][:, :3]
    # the transformed point in float32 with the oracle's operation order (nested fused multiply-adds) (residuals are ~1e-4 m on
    # ~3 m coordinates, so float32 rounding of T*v is the dominant term and must be reproduced);
    # everything after it in float64
    def fma32(a, b, c):  # float32 fused multiply-add: the float64 product of two floats is exact
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    vt = np.stack([fma32(np.full_like(vc[:, 0], P[i, 0]), vc[:, 0],
                         fma32(np.full_like(vc[:, 0], P[i, 1]), vc[:, 1],
                               fma32(np.full_like(vc[:, 0], P[i, 2]), vc[:, 2], np.full_like(vc[:, 0], P[i, 3]))))
                   for i in range(3)], axis=1).astype(np.float64)
    q = corr[ok]
    vp = prev.vmap(0).reshape(-1, 4)[q][:, :3].astype(np.float64)
    npv = prev.nmap(0).reshape(-1, 4)[q][:, :3].astype(np.float64)
    r = np.einsum("ij,ij->i", npv, vp - vt)
    J = np.concatenate([np.cross(vt, npv), npv], axis=1)
    A, b = J.T @ J, J.T @ r
    for i in range(6):
        for j in range(i, 6):
            assert abs(sums[SLOT_A[i][j]] - A[i, j]) <= 2e-5 * max(1.0, abs(A[i, j]))
    assert np.allclose(sums[24:30], b, rtol=1e-3, atol=2e-4)
    assert sums[31] == ok.sum()
    for dup, twin in ((6, 1), (16, 13), (22, 21)):  # duplicate slots carry the same sums
        assert sums[dup] == sums[twin]
    for ppt in (1, 8, 256):  # a different run geometry regroups the sum but not its value
        cfg2 = oracle.default_config(icp_ppt=ppt)
        s2, c2 = oracle.icp_sums(cfg2, 0, cur, prev, pose)
        assert np.array_equal(c2, corr) and s2[31] == sums[31]
        assert np.allclose(s2[:31], sums[:31], rtol=1e-4, atol=1e-4)


def _sqrt32(x):
    """correctly rounded float32 square root of a positive float32, through an integer square root"""
    import math
    from fractions import Fraction

    q = Fraction(float(x)) * Fraction(4) ** 80  # sqrt(q) = sqrt(x) * 2^80, far more than 24 significant bits
    r = math.isqrt(q.numerator // q.denominator)
    exact = q.denominator == 1 and r * r == q.numerator
    proxy = Fraction(r) if exact else Fraction(2 * r + 1, 2)  # strictly inside (r, r+1): never a float32 tie
    return _rn32(proxy / Fraction(2) ** 80)


def test_pyramid_vertices_normals_against_exact_rational_restatement(oracle, pkg):
    """Stages 1b + 2 restated independently of the C oracle: 2x2 pyramid rule with
    its sample counts, back-projection (viewerModule.c:343-345) with the per-level intrinsics, cross-product normals
    with the correctly rounded square root and reciprocal -- every operation rounded once, none fused.  Bit-equal at
    all three levels, validity included."""
    from fractions import Fraction as Fr

    w, h = 64, 48
    cfg = small_cfg(oracle, w=w, h=h)
    raw = pkg.synth_sequence(1, w, h, sequence=5, noise=1)[0]
    fr = oracle.OFrame(cfg, raw)
    F = lambda v: Fr(float(v))
    mul = lambda a, b: _rn32(F(a) * F(b))
    sub = lambda a, b: _rn32(F(a) - F(b))
    add = lambda a, b: _rn32(F(a) + F(b))
    div = lambda a, b: _rn32(F(a) / F(b))
    f32 = np.float32
    thr = mul(f32(3.0), f32(cfg.sigma_range_mm))
    depth = [fr.depth(0).copy()]
    for level in range(1, 3):
        src = depth[-1]
        hh, ww = src.shape[0] // 2, src.shape[1] // 2
        dst, cnt = np.zeros((hh, ww), dtype=np.float32), np.zeros((hh, ww), dtype=np.uint8)
        for y in range(hh):
            for x in range(ww):
                smp = [src[2 * y, 2 * x], src[2 * y, 2 * x + 1], src[2 * y + 1, 2 * x], src[2 * y + 1, 2 * x + 1]]
                centre = next((v for v in smp if v > 0), None)
                if centre is None:
                    continue
                tot, n = f32(0.0), 0
                for v in smp:
                    if v > 0 and abs(sub(v, centre)) <= thr:
                        tot, n = add(tot, v), n + 1
                dst[y, x], cnt[y, x] = div(tot, f32(n)), n
        assert np.array_equal(fr.depth(level).view(np.uint32), dst.view(np.uint32))
        assert np.array_equal(fr.pyrcnt(level), cnt) and cnt.max() == 4
        depth.append(dst)
    factor = f32(cfg.depth_factor)
    for level in range(3):
        g = oracle.level_geometry(cfg, level)
        d = depth[level]
        hh, ww = d.shape
        V = np.zeros((hh, ww, 4), dtype=np.float32)
        for v in range(hh):
            for u in range(ww):
                if d[v, u] > 0:
                    z = div(d[v, u], factor)
                    V[v, u] = (div(mul(sub(f32(u), f32(g.cx)), z), f32(g.fx)), div(mul(sub(f32(v), f32(g.cy)), z), f32(g.fy)), z, 1.0)
        N = np.zeros((hh, ww, 4), dtype=np.float32)
        for v in range(hh - 1):
            for u in range(ww - 1):
                p, px, py = V[v, u], V[v, u + 1], V[v + 1, u]
                if p[3] == 0 or px[3] == 0 or py[3] == 0:
                    continue
                ax, ay, az = sub(px[0], p[0]), sub(px[1], p[1]), sub(px[2], p[2])
                bx, by, bz = sub(py[0], p[0]), sub(py[1], p[1]), sub(py[2], p[2])
                nx, ny, nz = sub(mul(ay, bz), mul(az, by)), sub(mul(az, bx), mul(ax, bz)), sub(mul(ax, by), mul(ay, bx))
                len2 = add(add(mul(nx, nx), mul(ny, ny)), mul(nz, nz))
                if not len2 > f32(1e-24):
                    continue
                inv = div(f32(1.0), _sqrt32(len2))
                N[v, u] = (mul(nx, inv), mul(ny, inv), mul(nz, inv), 1.0)
        assert np.array_equal(fr.vmap(level).view(np.uint32), V.view(np.uint32)), level
        assert np.array_equal(fr.nmap(level).view(np.uint32), N.view(np.uint32)), level
        assert N[..., 3].sum() > 0.5 * hh * ww


def test_association_and_reduction_against_exact_rational_restatement(oracle, pkg):
    """Stages 3 + 4 restated independently of the C oracle: per pixel the nested-fma
    transform, the reciprocal and projection fma, the gates in their order with their reject codes, residual and
    Jacobian as fma chains, the 32 fused accumulations; then the lane / run / chain reduction order.  Exact rational
    arithmetic with one explicit float32 (or, across runs, float64) rounding per written operation.  The
    correspondence image must be equal and the 32 sums bit-equal."""
    from fractions import Fraction as Fr

    w, h, ppt = 64, 48, 4
    cfg = small_cfg(oracle, w=w, h=h, levels=1, icp_ppt=ppt)
    cfg.iters[0] = 1
    raw = pkg.synth_sequence(2, w, h, sequence=3)
    prev, cur = oracle.OFrame(cfg, raw[0]), oracle.OFrame(cfg, raw[1])
    rot = Rotation.from_rotvec([0.004, -0.007, 0.003]).as_matrix()
    pose = np.concatenate([rot, [[0.006], [-0.004], [0.009]]], axis=1).astype(np.float32).reshape(12)
    sums, corr = oracle.icp_sums(cfg, 0, cur, prev, pose)

    F = lambda v: Fr(float(v))
    mul = lambda a, b: _rn32(F(a) * F(b))
    sub = lambda a, b: _rn32(F(a) - F(b))
    add = lambda a, b: _rn32(F(a) + F(b))
    fma = lambda a, b, c: _rn32(F(a) * F(b) + F(c))
    f32 = np.float32
    g = oracle.level_geometry(cfg, 0)
    fx, fy, cxh, cyh = f32(g.fx), f32(g.fy), add(f32(g.cx), f32(0.5)), add(f32(g.cy), f32(0.5))
    dist2_thr, cos_thr = mul(f32(cfg.dist_thresh_m), f32(cfg.dist_thresh_m)), f32(cfg.cos_thresh)
    P = [f32(v) for v in pose]
    vc, nc = cur.vmap(0).reshape(-1, 4), cur.nmap(0).reshape(-1, 4)
    vp, npv = prev.vmap(0).reshape(-1, 4), prev.nmap(0).reshape(-1, 4)
    PA = [0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6]
    PB = [0, 2, 4, 0, 2, 4, 2, 4, 2, 4, 4, 4, 0, 2, 4]
    FLT_MIN = f32(1.17549435e-38)

    def pixel(p, acc):
        if vc[p, 3] == 0 or nc[p, 3] == 0:
            return -1
        x, y, z = vc[p, 0], vc[p, 1], vc[p, 2]
        tx = fma(P[0], x, fma(P[1], y, fma(P[2], z, P[3])))
        ty = fma(P[4], x, fma(P[5], y, fma(P[6], z, P[7])))
        tz = fma(P[8], x, fma(P[9], y, fma(P[10], z, P[11])))
        if not tz >= FLT_MIN:
            return -2
        iz = _rn32(1 / F(tz))
        ur, vr = fma(mul(tx, fx), iz, cxh), fma(mul(ty, fy), iz, cyh)
        if not (0 <= ur < w and 0 <= vr < h):
            return -3
        q = int(vr) * w + int(ur)
        if vp[q, 3] == 0 or npv[q, 3] == 0:
            return -4
        dx, dy, dz = sub(vp[q, 0], tx), sub(vp[q, 1], ty), sub(vp[q, 2], tz)
        if not fma(dz, dz, fma(dy, dy, mul(dx, dx))) <= dist2_thr:
            return -5
        nx, ny, nz = nc[p, 0], nc[p, 1], nc[p, 2]
        rnx = fma(P[2], nz, fma(P[1], ny, mul(P[0], nx)))
        rny = fma(P[6], nz, fma(P[5], ny, mul(P[4], nx)))
        rnz = fma(P[10], nz, fma(P[9], ny, mul(P[8], nx)))
        n0, n1, n2 = npv[q, 0], npv[q, 1], npv[q, 2]
        if not fma(rnz, n2, fma(rny, n1, mul(rnx, n0))) >= cos_thr:
            return -6
        X = [fma(ty, n2, -mul(tz, n1)), fma(tz, n0, -mul(tx, n2)), fma(tx, n1, -mul(ty, n0)), n0, n1, n2,
             fma(n2, dz, fma(n1, dy, mul(n0, dx))), f32(1.0)]
        for k in range(15):
            acc[2 * k] = fma(X[PA[k]], X[PB[k]], acc[2 * k])
            acc[2 * k + 1] = fma(X[PA[k]], X[PB[k] + 1], acc[2 * k + 1])
        acc[30] = fma(X[6], X[6], acc[30])
        acc[31] = fma(X[7], X[7], acc[31])
        return q

    npix, lanes = w * h, 32
    nruns = -(-npix // (lanes * ppt))
    want_corr = np.empty(npix, dtype=np.int32)
    partial = np.zeros((nruns, 32), dtype=np.float32)
    for run in range(nruns):
        acc = [[f32(0.0)] * 32 for _ in range(lanes)]
        for j in range(ppt):
            for l in range(lanes):
                p = j * lanes * nruns + lanes * run + l
                if p < npix:
                    want_corr[p] = pixel(p, acc[l])
        for k in range(32):
            v = [acc[l][k] for l in range(lanes)]
            s_ = 16
            while s_ >= 1:
                for l in range(s_):
                    v[l] = add(v[l], v[l + s_])
                s_ >>= 1
            partial[run, k] = v[0]
    want = np.zeros(32, dtype=np.float64)
    for k in range(32):
        chains = []
        for c in range(8):
            d = 0.0
            for run in range(c, nruns, 8):
                d = d + float(partial[run, k])  # IEEE double addition
            chains.append(d)
        tot = chains[0]
        for c in range(1, 8):
            tot = tot + chains[c]
        want[k] = tot
    assert np.array_equal(corr.reshape(-1), want_corr)
    assert (want_corr >= 0).sum() > 1500 and len(set(want_corr[want_corr < 0])) >= 3  # matches and several reject kinds
    assert np.array_equal(sums.view(np.uint64), want.view(np.uint64))


def test_solve_update_against_numpy_scipy(oracle):
    rng = np.random.default_rng(5)
    cfg = oracle.default_config()
    for trial in range(20):
        J = rng.normal(size=(400, 6))
        xi = rng.normal(size=6) * (0.02 if trial < 15 else 1.0)  # also large rotations for the exp series
        r = J @ xi
        A, b = J.T @ J, J.T @ r
        sums = np.zeros(32)
        for i in range(6):
            for j in range(i, 6):
                sums[SLOT_A[i][j]] = A[i, j]
        sums[24:30] = b
        sums[31] = 400
        R0 = Rotation.from_rotvec(rng.normal(size=3) * 0.3).as_matrix()
        t0 = rng.normal(size=3)
        pose_d = np.concatenate([R0, t0[:, None]], axis=1).reshape(12).copy()
        pose_f = np.zeros(12, dtype=np.float32)
        assert oracle.lib().yo_solve_update(C.byref(cfg), sums.ctypes.data, pose_d.ctypes.data, pose_f.ctypes.data) == 1
        x = np.linalg.solve(A, b)
        w, u = x[:3], x[3:]
        Rinc = Rotation.from_rotvec(w).as_matrix()
        th = np.linalg.norm(w)
        Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * Wx + (th - np.sin(th)) / th**3 * Wx @ Wx
        want = np.concatenate([Rinc @ R0, (Rinc @ t0 + V @ u)[:, None]], axis=1).reshape(12)
        assert np.allclose(pose_d, want, rtol=0, atol=1e-9)
        assert np.array_equal(pose_f, pose_d.astype(np.float32))


def test_solve_and_pose_update_against_operation_by_operation_restatement(oracle):
    """Stage 5 restated independently of the C oracle in Python floats -- IEEE doubles,
    one rounding per written operation, correctly rounded sqrt, no fused operations: Cholesky with one reciprocal
    per column, ascending forward and DESCENDING back substitution, the three 12-term Horner series in theta^2, the
    SE(3) exponential and T <- exp(xi) T with the parenthesisation of the specification.  Bit-equal, small and large
    rotations."""
    import math

    rng = np.random.default_rng(9)
    cfg = oracle.default_config()
    fact = [float(math.factorial(n)) for n in range(28)]

    def mat3(a, b):
        return [(a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j] for i in range(3) for j in range(3)]

    def solve_update(sums, pose):
        A = [[float(sums[SLOT_A[i][j]]) for j in range(6)] for i in range(6)]
        b = [float(sums[24 + i]) for i in range(6)]
        scale = max(A[i][i] for i in range(6))
        L = [[0.0] * 6 for _ in range(6)]
        inv = [0.0] * 6
        for j in range(6):
            d = A[j][j]
            for m in range(j):
                d = d - L[j][m] * L[j][m]
            assert d > 1e-12 * scale
            L[j][j] = math.sqrt(d)
            inv[j] = 1.0 / L[j][j]
            for i in range(j + 1, 6):
                s_ = A[i][j]
                for m in range(j):
                    s_ = s_ - L[i][m] * L[j][m]
                L[i][j] = s_ * inv[j]
        y, x = [0.0] * 6, [0.0] * 6
        for i in range(6):
            s_ = b[i]
            for m in range(i):
                s_ = s_ - L[i][m] * y[m]
            y[i] = s_ * inv[i]
        for i in range(5, -1, -1):
            s_ = y[i]
            for m in range(5, i, -1):
                s_ = s_ - L[m][i] * x[m]
            x[i] = s_ * inv[i]
        wx, wy, wz = x[0], x[1], x[2]
        t2 = (wx * wx + wy * wy) + wz * wz
        a = bb = c = 0.0
        for k in range(11, -1, -1):
            sgn = -1.0 if k & 1 else 1.0
            a = a * t2 + sgn / fact[2 * k + 1]
            bb = bb * t2 + sgn / fact[2 * k + 2]
            c = c * t2 + sgn / fact[2 * k + 3]
        W = [0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0]
        W2 = mat3(W, W)
        ident = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
        Ri = [(ident[i] + a * W[i]) + bb * W2[i] for i in range(9)]
        V = [(ident[i] + bb * W[i]) + c * W2[i] for i in range(9)]
        ti = [(V[3 * i] * x[3] + V[3 * i + 1] * x[4]) + V[3 * i + 2] * x[5] for i in range(3)]
        R = [float(pose[4 * i + j]) for i in range(3) for j in range(3)]
        t = [float(pose[4 * i + 3]) for i in range(3)]
        Rn = mat3(Ri, R)
        tn = [((Ri[3 * i] * t[0] + Ri[3 * i + 1] * t[1]) + Ri[3 * i + 2] * t[2]) + ti[i] for i in range(3)]
        return np.array([[Rn[3 * i], Rn[3 * i + 1], Rn[3 * i + 2], tn[i]] for i in range(3)]).reshape(12)

    for trial in range(30):
        J = rng.normal(size=(300, 6))
        xi = rng.normal(size=6) * (0.01 if trial < 20 else 1.5)
        A, b = J.T @ J, J.T @ (J @ xi)
        sums = np.zeros(32)
        for i in range(6):
            for j in range(i, 6):
                sums[SLOT_A[i][j]] = A[i, j]
        sums[24:30] = b
        sums[31] = 300
        R0 = Rotation.from_rotvec(rng.normal(size=3) * 0.4).as_matrix()
        pose_d = np.concatenate([R0, rng.normal(size=(3, 1))], axis=1).reshape(12).copy()
        want = solve_update(sums, pose_d)
        pose_f = np.zeros(12, dtype=np.float32)
        assert oracle.lib().yo_solve_update(C.byref(cfg), sums.ctypes.data, pose_d.ctypes.data, pose_f.ctypes.data) == 1
        assert np.array_equal(pose_d.view(np.uint64), want.view(np.uint64)), trial
        assert np.array_equal(pose_f, want.astype(np.float32))


def test_solve_failure_policy(oracle):
    cfg = oracle.default_config()
    ident = IDENT.astype(np.float64).copy()
    pf = IDENT.copy()
    sums = np.zeros(32)
    sums[31] = 10  # too few inliers
    assert oracle.lib().yo_solve_update(C.byref(cfg), sums.ctypes.data, ident.ctypes.data, pf.ctypes.data) == 0
    sums[31] = 1000  # enough inliers but a singular (all-zero) system
    assert oracle.lib().yo_solve_update(C.byref(cfg), sums.ctypes.data, ident.ctypes.data, pf.ctypes.data) == 0
    sums[0] = np.nan
    assert oracle.lib().yo_solve_update(C.byref(cfg), sums.ctypes.data, ident.ctypes.data, pf.ctypes.data) == 0
    assert np.array_equal(ident, IDENT.astype(np.float64))  # pose untouched, never NaN


def test_empty_and_ragged_inputs(oracle):
    cfg = small_cfg(oracle)
    empty = np.zeros((3, 120, 160), dtype=np.uint16)
    poses, status, _ = oracle.track_sequence(cfg, empty)
    assert list(status) == [1, 2, 2] and all(np.array_equal(p, IDENT) for p in poses)
    half = np.full((2, 120, 160), 1500, dtype=np.uint16)
    half[:, :, 80:] = 0  # ragged validity: half the image
    poses, status, _ = oracle.track_sequence(cfg, half)
    assert status[1] in (0, 2) and np.isfinite(poses).all()


def test_golden_fixture(oracle):
    g = np.load(os.path.join(HERE, "golden", "golden_160x120.npz"))
    frames = g["frames"]
    cfg = small_cfg(oracle)
    poses, status, _ = oracle.track_sequence(cfg, frames)
    assert np.array_equal(poses.view(np.uint32), g["poses"].view(np.uint32))
    assert np.array_equal(status, g["status"])
    prev, cur = oracle.OFrame(cfg, frames[0]), oracle.OFrame(cfg, frames[1])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    for level in range(3):
        sums, corr = oracle.icp_sums(cfg, level, cur, prev, IDENT)
        assert np.array_equal(sums.view(np.uint64), g[f"sums_l{level}"].view(np.uint64))
        assert np.array_equal(corr, g[f"corr_l{level}"])
        assert np.array_equal(cur.mask(level), g[f"mask_l{level}"])
        assert sha(cur.depth(level)) == str(g[f"depth_sha_l{level}"])
        assert sha(cur.vmap(level)) == str(g[f"vmap_sha_l{level}"])
        assert sha(cur.nmap(level)) == str(g[f"nmap_sha_l{level}"])


def test_pose_agrees_with_an_order_independent_float64_icp(oracle):
    """The oracle's reduction order is the kernel's geometry (icp_ppt, lane -> run -> chain), so oracle == device bit
    for bit says nothing about whether that order matters.  Outside anchor: tests/numpy_icp.py solves the same
    pairs in float64 with order-free numpy sums, numpy.linalg.solve and a libm SE(3) exponential.  The oracle's
    relative poses must agree within the north-star tolerance (1e-4 m, 1e-4 rad) at BOTH reduction geometries
    the build uses (icp_ppt 64 and 128), with and without sensor noise."""
    import youth_pkg

    import numpy_icp as NI

    pkg = youth_pkg.load()
    for noise in (0, 1):
        frames = pkg.synth_sequence(4, noise=noise, first=40)
        for ppt in (64, 128):
            cfg = oracle.default_config(icp_ppt=ppt)
            ofr = [oracle.OFrame(cfg, f) for f in frames]
            geom = []
            for l in range(cfg.levels):
                g = oracle.level_geometry(cfg, l)
                geom.append((g.w, g.h, g.fx, g.fy, g.cx, g.cy))
            for i in range(1, len(frames)):
                rel, st, inl = oracle.track_pair(cfg, ofr[i], ofr[i - 1])
                maps = lambda fr: [(fr.vmap(l), fr.nmap(l)) for l in range(cfg.levels)]
                T, inl64 = NI.icp_pair(cfg.levels, list(cfg.iters), geom, maps(ofr[i]), maps(ofr[i - 1]),
                                       cfg.dist_thresh_m, cfg.cos_thresh, cfg.min_inliers)
                R = rel.reshape(3, 4)
                assert st == 0
                assert np.linalg.norm(R[:, 3] - T[:3, 3]) < 1e-4, (noise, ppt, i)
                assert NI.rot_angle(R[:, :3], T[:3, :3]) < 1e-4, (noise, ppt, i)
                assert abs(inl - inl64) <= max(50, inl // 500)  # threshold ties may fall either way in double


def test_pair_parallel_tracking_equals_the_serial_tracker(oracle, small_seq):
    frames, _ = small_seq
    cfg = oracle.default_config(icp_ppt=128)
    a, sa, _ = oracle.track_sequence(cfg, frames)
    b, sb, inl, _ = oracle.track_sequence_parallel(cfg, frames, threads=4)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(sa, sb)
    assert inl[0] == 0 and inl[1:].min() > 100000
