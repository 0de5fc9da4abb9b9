"""An INDEPENDENT float64 statement of stages 3-5 (numpy, vectorised, order-free sums, numpy.linalg.solve, a
closed-form SE(3) exponential): the outside anchor of the pose-level parity tests.  It shares no code and no
summation order with the C oracle or the kernels -- the oracle's reduction geometry follows the kernel's
(icp_ppt, lane / run / chain order), so bit equality between those two cannot by itself show that the order is
harmless; agreement with this order-independent double-precision solve within the north-star tolerance
(1e-4 m, 1e-4 rad) does.

Inputs are the vertex / normal maps of the two frames at every level (float [h][w][4]: x, y, z, valid), as the
oracle or the device (debug read-back) produce them; DESIGN.md section 3, items 5-8 is the text it follows.
Test infrastructure only."""
import numpy as np


def se3_exp(xi):
    """xi = (wx, wy, wz, tx, ty, tz) -> 4x4, Rodrigues + the V matrix, in float64 with libm."""
    w, u = np.asarray(xi[:3], dtype=np.float64), np.asarray(xi[3:], dtype=np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-9:
        A, B, Cc = 1.0 - th * th / 6.0, 0.5 - th * th / 24.0, 1.0 / 6.0 - th * th / 120.0
    else:
        A, B, Cc = np.sin(th) / th, (1.0 - np.cos(th)) / th ** 2, (th - np.sin(th)) / th ** 3
    R = np.eye(3) + A * K + B * (K @ K)
    V = np.eye(3) + B * K + Cc * (K @ K)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ u
    return T


def icp_pair(levels, iters, geom, cur_maps, prev_maps, dist_thresh=0.10, cos_thresh=None, min_inliers=100):
    """Relative pose prev<-cur of one frame pair, coarse to fine.
    geom[l] = (w, h, fx, fy, cx, cy); cur_maps[l] / prev_maps[l] = (vmap, nmap) float arrays [h][w][4].
    Returns (T 4x4 float64, inliers of the last iteration)."""
    if cos_thresh is None:
        cos_thresh = float(np.cos(np.deg2rad(20.0)))
    T = np.eye(4)
    inl = 0
    for l in range(levels - 1, -1, -1):
        w, h, fx, fy, cx, cy = geom[l]
        vc = cur_maps[l][0].reshape(-1, 4).astype(np.float64)
        nc = cur_maps[l][1].reshape(-1, 4).astype(np.float64)
        vp = prev_maps[l][0].reshape(-1, 4).astype(np.float64)
        npv = prev_maps[l][1].reshape(-1, 4).astype(np.float64)
        ok_c = (vc[:, 3] != 0) & (nc[:, 3] != 0)
        for _ in range(iters[l]):
            R, t = T[:3, :3], T[:3, 3]
            p = vc[:, :3] @ R.T + t
            n = nc[:, :3] @ R.T
            z = np.where(p[:, 2] > 0, p[:, 2], 1.0)
            u = np.floor(p[:, 0] * fx / z + cx + 0.5)
            v = np.floor(p[:, 1] * fy / z + cy + 0.5)
            ok = ok_c & (p[:, 2] > 0) & (u >= 0) & (u < w) & (v >= 0) & (v < h)
            q = np.where(ok, v * w + u, 0).astype(np.int64)
            ok &= (vp[q, 3] != 0) & (npv[q, 3] != 0)
            d = vp[q, :3] - p
            ok &= (d * d).sum(1) <= dist_thresh * dist_thresh
            ok &= (n * npv[q, :3]).sum(1) >= cos_thresh
            inl = int(ok.sum())
            if inl < min_inliers:
                break
            nn, pp, dd = npv[q, :3][ok], p[ok], d[ok]
            r = (nn * dd).sum(1)
            J = np.concatenate([np.cross(pp, nn), nn], axis=1)
            A = J.T @ J
            b = J.T @ r
            try:
                xi = np.linalg.solve(A, b)
            except np.linalg.LinAlgError:
                break
            T = se3_exp(xi) @ T
    return T, inl


def chain(rel_poses):
    """world poses (first = identity) from relative poses prev<-cur"""
    Wm = np.eye(4)
    out = [Wm.copy()]
    for T in rel_poses:
        Wm = Wm @ T
        out.append(Wm.copy())
    return out


def rot_angle(Ra, Rb):
    """angle of Ra^T Rb from its skew part (sin) and trace (cos): well conditioned near zero, where arccos of the
    trace alone turns a 1e-7 rounding of a matrix entry into 4e-4 rad"""
    D = np.asarray(Ra, dtype=np.float64).T @ np.asarray(Rb, dtype=np.float64)
    s = 0.5 * np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    c = (np.trace(D) - 1.0) / 2.0
    return float(np.arctan2(s, c))
