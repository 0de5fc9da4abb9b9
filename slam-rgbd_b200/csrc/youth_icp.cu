/*
 * youth_icp.cu -- stages 3-5 of the tracker on sm_100a (see youth_common.cuh for the contract).
 *
 *   k_icp        one warp per run of pixels, software-pipelined projective association + point-to-plane
 *                residual / Jacobian, 32 FFMA2-accumulated sums per lane, warp butterfly; the last run of each frame
 *                pair to finish then runs the fixed-order cross-run reduction (double), the 6x6 Cholesky solve and
 *                the SE(3) exponential update -- one launch per ICP iteration, no host sync, no block barrier, no
 *                atomics on data
 *   k_icp_fused  the whole coarse-to-fine schedule of a few pairs in one launch (opt-in, measured slower)
 *   k_compose    pose chain: world pose = world pose * relative pose, trajectory append
 *   k_rcp_check  parity hook for the reciprocal used by stage 3
 *
 * Pipeline depth: the loop keeps two pixels per lane in flight (streamed record two steps ahead, gather consumed
 * two steps after issue) at 96 registers = 5 CTAs of 4 warps per SM.  What was measured on B200 before settling
 * there (profiles/README.md, r2a-r2c; tools/membench2.cu): three- to six-deep register pipelines at 4 or 3 CTAs per
 * SM, the streamed half through a shared-memory ring filled by cp.async or by 1-D bulk copies, prefetch.global.L1 as
 * register-free depth, recomputing vx, vy instead of loading them -- none was faster; the bare access pattern itself
 * needs 432 us per 300-pair level-0 launch at this occupancy and depth, the kernel takes 487.
 */
#include "youth_common.cuh"

#include <string.h>

/* ------------------------------------------------------------------ stage 5 (device function) */

/* sin(t)/t, (1-cos t)/t^2, (t-sin t)/t^3 as 12-term Horner polynomials in t^2.  Called by a full
 * warp: lanes 0, 1, 2 (mod 3) each evaluate ONE of the three series (same operation sequence as the
 * CPU checker's three interleaved series), then the results are exchanged -- a third of the
 * dependent double-precision chain. */
__device__ __forceinline__ void so3_coeffs_warp(double t2, int lane, double* A, double* B, double* C) {
  const double f[28] = {1.0,
                        1.0,
                        2.0,
                        6.0,
                        24.0,
                        120.0,
                        720.0,
                        5040.0,
                        40320.0,
                        362880.0,
                        3628800.0,
                        39916800.0,
                        479001600.0,
                        6227020800.0,
                        87178291200.0,
                        1307674368000.0,
                        20922789888000.0,
                        355687428096000.0,
                        6402373705728000.0,
                        121645100408832000.0,
                        2432902008176640000.0,
                        51090942171709440000.0,
                        1124000727777607680000.0,
                        25852016738884976640000.0,
                        620448401733239439360000.0,
                        15511210043330985984000000.0,
                        403291461126605635584000000.0,
                        10888869450418352160768000000.0};
  const int which = lane % 3;
  double acc = 0.0;
#pragma unroll
  for (int k = 11; k >= 0; --k) {
    const double sgn = (k & 1) ? -1.0 : 1.0;
    const double ca = sgn / f[2 * k + 1], cb = sgn / f[2 * k + 2], cc = sgn / f[2 * k + 3]; /* folded at compile time */
    acc = acc * t2 + (which == 0 ? ca : (which == 1 ? cb : cc));
  }
  *A = __shfl_sync(0xffffffffu, acc, 0);
  *B = __shfl_sync(0xffffffffu, acc, 1);
  *C = __shfl_sync(0xffffffffu, acc, 2);
}

/* slot of A[i][j] in the 32 sums (pairs formed for fma.rn.f32x2, see icp_pixel) */
__device__ __forceinline__ int sums_slot_a(int i, int j) {
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  /* rows: 0 -> 0..5, 1 -> 6..11 (slot 6 duplicates A10), 2 -> 12..15 (from column 2),
   * 3 -> 16..19 (slot 16 duplicates A32), 4 -> 20..21, 5 -> 22..23 (slot 22 duplicates A54) */
  const int base = lo == 0 ? 0 : (lo == 1 ? 6 : (lo == 2 ? 10 : (lo == 3 ? 14 : (lo == 4 ? 16 : 18))));
  return base + hi;
}

/* Stage 5 on one warp: lane i owns row i of the 6x6 system (Cholesky with one reciprocal
 * per column, forward substitution ascending, back substitution descending), then lanes
 * 0..2 each produce one row of exp(xi) * T.  Every value is computed by the same operation
 * sequence as the CPU checker, so the result is bit-identical.  Must be called by all 32
 * lanes of a warp; returns 1 (warp-uniform) when the pose was updated. */
__device__ __forceinline__ int solve_update_warp(const double* tot, int min_inliers, double* pose_d, float* pose_f,
                                                 int lane) {
  const unsigned FULL = 0xffffffffu;
  if (!(tot[YOUTH_SUMS_COUNT] >= (double)min_inliers)) return 0;
  const int i = lane < 6 ? lane : 5;
  double a[6], l[6], inv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    a[j] = tot[sums_slot_a(i, j)];
    l[j] = 0.0;
  }
  double scale = tot[sums_slot_a(0, 0)];
#pragma unroll
  for (int j = 1; j < 6; ++j) {
    const double d = tot[sums_slot_a(j, j)];
    if (d > scale) scale = d;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = a[j];
#pragma unroll
    for (int m = 0; m < j; ++m) {
      const double Ljm = __shfl_sync(FULL, l[m], j);
      s = s - l[m] * Ljm;
    }
    const double diag = __shfl_sync(FULL, s, j);
    if (!(diag > 1e-12 * scale)) return 0;
    const double r = sqrt(diag);
    inv[j] = 1.0 / r;
    l[j] = (lane == j) ? r : s * inv[j];
  }
  double y[6], x[6];
  double t = tot[YOUTH_SUMS_B0 + i];
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    y[m] = __shfl_sync(FULL, t * inv[m], m);
    if (lane > m) t = t - l[m] * y[m];
  }
  t = y[0];
#pragma unroll
  for (int m = 1; m < 6; ++m)
    if (i == m) t = y[m];
#pragma unroll
  for (int m = 5; m >= 0; --m) {
    x[m] = __shfl_sync(FULL, t * inv[m], m);
#pragma unroll
    for (int ii = 0; ii < m; ++ii) {
      const double v = __shfl_sync(FULL, l[ii], m);
      if (lane == ii) t = t - v * x[m];
    }
  }
#pragma unroll
  for (int m = 0; m < 6; ++m)
    if (!(x[m] > -1e6 && x[m] < 1e6)) return 0;

  const double wx = x[0], wy = x[1], wz = x[2];
  const double t2 = (wx * wx + wy * wy) + wz * wz;
  double Ac, Bc, Cc;
  so3_coeffs_warp(t2, lane, &Ac, &Bc, &Cc);
  const double Wm[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
  const int r = lane < 3 ? lane : 2;
  double wrow[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) wrow[j] = r == 0 ? Wm[j] : (r == 1 ? Wm[3 + j] : Wm[6 + j]);
  double Ri[3], V[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double w2 = (wrow[0] * Wm[j] + wrow[1] * Wm[3 + j]) + wrow[2] * Wm[6 + j];
    const double id = (j == r) ? 1.0 : 0.0;
    Ri[j] = (id + Ac * wrow[j]) + Bc * w2;
    V[j] = (id + Bc * wrow[j]) + Cc * w2;
  }
  const double ti = (V[0] * x[3] + V[1] * x[4]) + V[2] * x[5];
  double R[9], tt[3];
#pragma unroll
  for (int a2 = 0; a2 < 3; ++a2) {
#pragma unroll
    for (int j = 0; j < 3; ++j) R[3 * a2 + j] = pose_d[4 * a2 + j];
    tt[a2] = pose_d[4 * a2 + 3];
  }
  double Rn[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) Rn[j] = (Ri[0] * R[j] + Ri[1] * R[3 + j]) + Ri[2] * R[6 + j];
  const double tn = ((Ri[0] * tt[0] + Ri[1] * tt[1]) + Ri[2] * tt[2]) + ti;
  __syncwarp();
  if (lane < 3) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      pose_d[4 * r + j] = Rn[j];
      pose_f[4 * r + j] = (float)Rn[j];
    }
    pose_d[4 * r + 3] = tn;
    pose_f[4 * r + 3] = (float)tn;
  }
  return 1;
}

/* ------------------------------------------------------------------ k_icp */

/* Stage 3 is split in two so that a lane can keep several pixels in flight:
 *   icp_front  current vertex/normal + pose -> projected previous-frame pixel index (or a
 *              negative reject code), transformed point and rotated normal;
 *   icp_back   gathered previous vertex/normal -> remaining gates, residual, Jacobian and the
 *              32 fused multiply-add accumulations (slot layout: include/youth_cuda.h).
 * Every __fmaf_rn / fma2 here is part of the arithmetic specification (the CPU checker calls
 * fmaf() at the same places); nothing else may be contracted (--fmad=false). */
struct IcpPend {
  float tx, ty, tz;    /* T v            */
  float rnx, rny, rnz; /* R n            */
  int q;               /* >= 0: previous-frame pixel index; < 0: reject code */
};

struct F3 {
  float x, y, z;
};

/* Both halves are written branch-free (selects instead of early returns) so that the compiler
 * can interleave the arithmetic of consecutive pixels and no reconvergence barriers sit inside
 * the pipelined loop.  The gate order still decides which reject code is reported. */
/* one map record = the three float2 planes (vx,vy) (vz,nx) (ny,nz) of a pixel */
struct Rec3 {
  float2 a, b, c;
};

/* one frame's maps: first plane and the byte distance between planes (uniform per warp) */
struct RecBase {
  const float2* a;
  long long plane_bytes;
};

/* Predicated record loads (ld.global.nc = the read-only path; the maps are not written while k_icp
 * runs).  The destination registers are read-write operands: when the predicate is false they keep
 * their contents, no zero-fill and no branch.  The plane addresses are formed inside the asm (a chain
 * of 64-bit adds of the plane size) so that the compiler does not re-associate them into longer index
 * arithmetic. */
__device__ __forceinline__ void ld_rec_gather(int q, const RecBase& base, Rec3& r) { /* loads iff q >= 0 */
  asm("{\n\t.reg .pred p;\n\t.reg .b64 pa, pb, pc;\n\t"
      "setp.ge.s32 p, %6, 0;\n\t"
      "mad.wide.s32 pa, %6, 8, %7;\n\t"
      "add.s64 pb, pa, %8;\n\t"
      "add.s64 pc, pb, %8;\n\t"
      "@p ld.global.nc.v2.f32 {%0, %1}, [pa];\n\t"
      "@p ld.global.nc.v2.f32 {%2, %3}, [pb];\n\t"
      "@p ld.global.nc.v2.f32 {%4, %5}, [pc];\n\t}"
      : "+f"(r.a.x), "+f"(r.a.y), "+f"(r.b.x), "+f"(r.b.y), "+f"(r.c.x), "+f"(r.c.y)
      : "r"(q), "l"(base.a), "l"(base.plane_bytes));
}

/* loads iff j < nj (the lane has that pixel); otherwise the record is marked invalid through its normal
 * (nx = YK_N_INVALID) and the other registers keep their contents.  (Moving the mark to the consumer -- one more
 * predicate term in icp_front -- was measured on B200: bit-identical, no gain; profiles/README.md, r1u.) */
__device__ __forceinline__ void ld_rec_stream(int j, int nj, const float2* pa, long long plane_bytes, Rec3& r) {
  asm("{\n\t.reg .pred p;\n\t.reg .b64 pb, pc;\n\t"
      "setp.lt.s32 p, %6, %7;\n\t"
      "add.s64 pb, %8, %9;\n\t"
      "add.s64 pc, pb, %9;\n\t"
      "@p ld.global.nc.v2.f32 {%0, %1}, [%8];\n\t"
      "@p ld.global.nc.v2.f32 {%2, %3}, [pb];\n\t"
      "@p ld.global.nc.v2.f32 {%4, %5}, [pc];\n\t"
      "@!p mov.f32 %3, 0f40000000;\n\t}"
      : "+f"(r.a.x), "+f"(r.a.y), "+f"(r.b.x), "+f"(r.b.y), "+f"(r.c.x), "+f"(r.c.y)
      : "r"(j), "r"(nj), "l"(pa), "l"(plane_bytes));
}



/* parity hook: counts the floats with bit patterns in [lo, hi] whose rcp_normal differs from the IEEE
 * reciprocal (__frcp_rn) */
__global__ void __launch_bounds__(256) k_rcp_check(uint32_t lo, uint32_t hi, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (unsigned long long b = (unsigned long long)lo + blockIdx.x * 256ull + threadIdx.x; b <= hi; b += 256ull * gridDim.x) {
    const float x = __uint_as_float((uint32_t)b);
    if (__float_as_uint(rcp_normal(x)) != __float_as_uint(__frcp_rn(x))) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

#define YK_Z_FRONT_MIN 1.17549435e-38f /* FLT_MIN: v'.z must be a positive normal float */

template <bool CODES>
__device__ __forceinline__ void icp_front(const LevelGeom& g, const F3 vc, const F3 nc, const float* P,
                                          const RecBase& prv, IcpPend& pd, Rec3& gr) {
  pd.tx = __fmaf_rn(P[0], vc.x, __fmaf_rn(P[1], vc.y, __fmaf_rn(P[2], vc.z, P[3])));
  pd.ty = __fmaf_rn(P[4], vc.x, __fmaf_rn(P[5], vc.y, __fmaf_rn(P[6], vc.z, P[7])));
  pd.tz = __fmaf_rn(P[8], vc.x, __fmaf_rn(P[9], vc.y, __fmaf_rn(P[10], vc.z, P[11])));
  if (CODES) {
    /* the debug kernel reports which gate rejected the pixel, with the gates as the specification words
     * them (IEEE division, float comparisons against the image size) */
    const bool valid = (vc.z > 0.0f) & YK_N_VALID(nc.x); /* vertex / normal validity is encoded in the values */
    const bool front_ok = pd.tz >= YK_Z_FRONT_MIN;
    const float iz = 1.0f / (front_ok ? pd.tz : 1.0f);
    const float ur = __fmaf_rn(pd.tx * g.fx, iz, g.cxh);
    const float vr = __fmaf_rn(pd.ty * g.fy, iz, g.cyh);
    const bool inside = (ur >= 0.0f) & (ur < (float)g.w) & (vr >= 0.0f) & (vr < (float)g.h);
    /* nearest pixel: floor(u + 0.5), the 0.5 is folded into cxh/cyh; cvt.rzi saturates (NaN -> 0), so
     * the discarded conversion of an out-of-image value is well defined on the device */
    const int q = __float2int_rz(vr) * g.w + __float2int_rz(ur);
    pd.q = !valid ? YOUTH_REJ_CUR_INVALID : (!front_ok ? YOUTH_REJ_BEHIND : (!inside ? YOUTH_REJ_OUT_OF_IMAGE : q));
    ld_rec_gather(pd.q, prv, gr);
  } else {
    /* the product only needs q < 0 for a rejected pixel.  Same gates, cheaper form: the quotient of a
     * rejected v'.z is never used, so no select in front of the reciprocal; 0 <= u + 1/2 < w is tested
     * on the floor-converted integer as one unsigned comparison (floor == truncation where it passes,
     * the conversion saturates, -0.0 converts to 0 as the float test accepts it; a NaN cannot occur:
     * poses are finite and v'.z is normal); the five gates chain through one predicate. */
    const float iz = rcp_normal(pd.tz);
    const float ur = __fmaf_rn(pd.tx * g.fx, iz, g.cxh);
    const float vr = __fmaf_rn(pd.ty * g.fy, iz, g.cyh);
    const int ui = __float2int_rd(ur), vi = __float2int_rd(vr);
    const int q = vi * g.w + ui;
    asm("{\n\t.reg .pred p;\n\t.reg .b64 pa, pb, pc;\n\t"
        "setp.lt.f32 p, %8, 0f3FC00000;\n\t"          /* YK_N_VALID: nx < 1.5 (implies a valid vertex, %7) */
        "setp.ge.and.f32 p, %9, 0f00800000, p;\n\t"   /* v'.z >= FLT_MIN */
        "setp.lt.and.u32 p, %10, %11, p;\n\t"
        "setp.lt.and.u32 p, %12, %13, p;\n\t"
        "selp.s32 %0, %14, -1, p;\n\t"
        /* gather of the matched previous-frame record, under the same predicate */
        "mad.wide.s32 pa, %14, 8, %15;\n\t"
        "add.s64 pb, pa, %16;\n\t"
        "add.s64 pc, pb, %16;\n\t"
        "@p ld.global.nc.v2.f32 {%1, %2}, [pa];\n\t"
        "@p ld.global.nc.v2.f32 {%3, %4}, [pb];\n\t"
        "@p ld.global.nc.v2.f32 {%5, %6}, [pc];\n\t}"
        : "=r"(pd.q), "+f"(gr.a.x), "+f"(gr.a.y), "+f"(gr.b.x), "+f"(gr.b.y), "+f"(gr.c.x), "+f"(gr.c.y)
        : "f"(vc.z), "f"(nc.x), "f"(pd.tz), "r"(ui), "r"(g.w), "r"(vi), "r"(g.h), "r"(q), "l"(prv.a), "l"(prv.plane_bytes));
  }
  pd.rnx = __fmaf_rn(P[2], nc.z, __fmaf_rn(P[1], nc.y, P[0] * nc.x));
  pd.rny = __fmaf_rn(P[6], nc.z, __fmaf_rn(P[5], nc.y, P[4] * nc.x));
  pd.rnz = __fmaf_rn(P[10], nc.z, __fmaf_rn(P[9], nc.y, P[8] * nc.x));
}

__device__ __forceinline__ int icp_back(float dist2_thr, float cos_thr, const IcpPend& pd, const F3 vp,
                                        const F3 np, float2* acc2) {
  const bool ok0 = pd.q >= 0;
  const bool ok1 = ok0 & YK_N_VALID(np.x); /* a valid normal implies a valid vertex (stage 2) */
  const float dx = vp.x - pd.tx, dy = vp.y - pd.ty, dz = vp.z - pd.tz;
  const float dist2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
  const bool ok2 = ok1 & (dist2 <= dist2_thr);
  const float cosang = __fmaf_rn(pd.rnz, np.z, __fmaf_rn(pd.rny, np.y, pd.rnx * np.x));
  const bool ok3 = ok2 & (cosang >= cos_thr);
  /* A rejected pixel gets the zero normal: J = (v' x 0, 0) and r = 0 . d are then +-0 (every other
   * operand is finite: map data, finite poses), its 32 products are +-0, and acc + (+-0) == acc -- the
   * accumulators never hold -0 -- so this is bit-identical to skipping the pixel (which is what the CPU
   * checker does), at four selects and no branch. */
  const float nx = ok3 ? np.x : 0.0f, ny = ok3 ? np.y : 0.0f, nz = ok3 ? np.z : 0.0f;
  const float one = ok3 ? 1.0f : 0.0f;
  const float r = __fmaf_rn(nz, dz, __fmaf_rn(ny, dy, nx * dx));
  const float J0 = __fmaf_rn(pd.ty, nz, -(pd.tz * ny));
  const float J1 = __fmaf_rn(pd.tz, nx, -(pd.tx * nz));
  const float J2 = __fmaf_rn(pd.tx, ny, -(pd.ty * nx));
  const float J3 = nx, J4 = ny, J5 = nz;
  const float2 P01 = make_float2(J0, J1), P23 = make_float2(J2, J3), P45 = make_float2(J4, J5);
  const float2 B0 = make_float2(J0, J0), B1 = make_float2(J1, J1), B2 = make_float2(J2, J2);
  const float2 B3 = make_float2(J3, J3), B4 = make_float2(J4, J4), B5 = make_float2(J5, J5);
  const float2 Br = make_float2(r, r), R1 = make_float2(r, one);
  acc2[0] = fma2(B0, P01, acc2[0]);
  acc2[1] = fma2(B0, P23, acc2[1]);
  acc2[2] = fma2(B0, P45, acc2[2]);
  acc2[3] = fma2(B1, P01, acc2[3]);
  acc2[4] = fma2(B1, P23, acc2[4]);
  acc2[5] = fma2(B1, P45, acc2[5]);
  acc2[6] = fma2(B2, P23, acc2[6]);
  acc2[7] = fma2(B2, P45, acc2[7]);
  acc2[8] = fma2(B3, P23, acc2[8]);
  acc2[9] = fma2(B3, P45, acc2[9]);
  acc2[10] = fma2(B4, P45, acc2[10]);
  acc2[11] = fma2(B5, P45, acc2[11]);
  acc2[12] = fma2(Br, P01, acc2[12]);
  acc2[13] = fma2(Br, P23, acc2[13]);
  acc2[14] = fma2(Br, P45, acc2[14]);
  acc2[15] = fma2(R1, R1, acc2[15]);
  return !ok0 ? pd.q
              : (!ok1 ? YOUTH_REJ_PREV_INVALID : (!ok2 ? YOUTH_REJ_DISTANCE : (!ok3 ? YOUTH_REJ_ANGLE : pd.q)));
}



/* transposing butterfly: 32 per-lane accumulators -> lane L holds slot L summed over the
 * warp with the pairwise tree of strides 16, 8, 4, 2, 1 (31 shuffles instead of 160) */
template <int M>
__device__ __forceinline__ void butterfly_step(float* acc, int lane) {
  const bool up = (lane & M) != 0;
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const float send = up ? acc[i] : acc[i + M];
    const float keep = up ? acc[i + M] : acc[i];
    acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
  }
}

/* k_icp: one WARP per run of 32*ppr pixels, no block-level synchronisation.
 * Lane l of run k walks pixels j*(32*nruns) + 32*k + l (j ascending) through a two-stage software pipeline
 * (streaming loads of pixel j+1 and the gather of pixel j are in flight while pixel j-1 is
 * finished), accumulates into 16 float2 registers with FFMA2, and the warp reduces with the
 * transposing butterfly.  The last run of a pair to arrive (ticket counter) sums the run
 * partials in the fixed order and runs the warp-parallel solve. */
/* LAST_CTA (the few-pairs, latency-bound launches of the live and frame-to-model paths): the ticket is
 * taken per CTA after one block barrier, and the four warps of the last CTA share the cross-run
 * reduction (two of the eight chains each) -- same order of additions, a quarter of the serial loads. */
template <bool DEBUG, bool LAST_CTA = false>
__global__ void __launch_bounds__(32 * YK_ICP_WARPS, LAST_CTA ? 2 : YK_ICP_MIN_BLOCKS) k_icp(const __grid_constant__ IcpParams P) {
  __shared__ double s_tot[YK_ICP_WARPS][32];
  __shared__ double s_chain[LAST_CTA ? 8 : 1][32];
  __shared__ unsigned int s_ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int run = blockIdx.x * YK_ICP_WARPS + warp;
  const int sq = blockIdx.z, fi = P.f0 + blockIdx.y; /* grid = (CTAs of a pair, frames of the range, sequences) */
  const int pair = sq * P.ring.n + fi;
  if (!LAST_CTA && run >= P.nruns) return;
  const bool has_run = run < P.nruns; /* LAST_CTA: warps without a run still meet the block barrier */
  int s, cur_slot, prev_slot;
  if (DEBUG && P.dbg_cur_slot >= 0) {
    s = P.dbg_stream;
    cur_slot = P.dbg_cur_slot;
    prev_slot = P.dbg_prev_slot;
  } else {
    s = sq;
    const int i = fi;
    if (P.seq_count[s] + i == 0) return; /* first frame of a sequence: no predecessor, pose stays identity */
    cur_slot = ring_slot(P.ring, i);
    prev_slot = (cur_slot + P.ring.R - 1) % P.ring.R;
  }
  const size_t stream_base = (size_t)s * P.ring.R, npx = (size_t)P.npix;
  const float2* __restrict__ cur = P.maps + (stream_base + cur_slot) * 3 * npx;  /* (vx,vy) (vz,nx) (ny,nz) */
  const float2* __restrict__ prv =
      P.model != nullptr ? P.model + (size_t)s * 3 * npx : P.maps + (stream_base + prev_slot) * 3 * npx;
  const float* pose_g = P.pose_f + pair * 12; /* prev<-cur pose of this pair */
  float pose[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) pose[k] = __ldg(pose_g + k);

  float2 acc2[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc2[k] = make_float2(0.0f, 0.0f);

  /* Software pipeline, per lane.  Pixel j of this lane = j * (32 * nruns) + 32 * run + lane: at
   * every step the runs of a pair read one contiguous span of the maps together.  Iteration j:
   *   back(j-2)  uses the gather issued two iterations ago
   *   front(j)   uses the streaming record loaded two iterations ago, issues the gather of pixel j
   *              (into the registers back(j-2) has just released)
   *   prefetch   streaming record of pixel j+2
   * so every load has two iterations of other work to hide behind (tools/membench.cu: SD=2, GD=2).
   * Loads are predicated, not zero-filled: a record that was not loaded keeps stale (finite or not,
   * it does not matter) register contents and is gated out -- the gather by q < 0, the streaming
   * record by its normal, forced to the invalid marker when the lane has no pixel j+2.
   * (measured on B200: predicated loads beat unconditional loads from clamped addresses -- 6.34 vs
   * 6.94 ms per 300 pairs x 10 iterations -- a rejected pixel's gather is pure cost) */
  const Rec3 zrec = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  const int npx_i = P.npix;
  const int p0 = run * 32 + lane;
  const int pstep = 32 * P.nruns;
  /* pixels of this lane: j < nj  <=>  p0 + j * pstep < npix */
  const int ppr = (LAST_CTA && !has_run) ? 0 : P.ppr; /* a warp without a run accumulates nothing */
  int nj = p0 < npx_i ? (npx_i - p0 + pstep - 1) / pstep : 0;
  nj = nj < ppr ? nj : ppr;
  const long long plane_bytes = (long long)npx * (long long)sizeof(float2);
  const RecBase prvb = {prv, plane_bytes};
  const float2* sp = cur + p0; /* streaming pointer: pixel of the next prefetch */
  Rec3 s0 = zrec, s1 = zrec;
  {
    ld_rec_stream(0, nj, sp, plane_bytes, s0);
    sp += pstep;
    ld_rec_stream(1, nj, sp, plane_bytes, s1);
    sp += pstep;
  }
  IcpPend pd0, pd1;
  pd0.tx = pd0.ty = pd0.tz = pd0.rnx = pd0.rny = pd0.rnz = 0.0f;
  pd0.q = YOUTH_REJ_CUR_INVALID;
  pd1 = pd0;
  Rec3 g0 = zrec, g1 = zrec;
#pragma unroll 2 /* a multiple of 2 makes the two-deep register rotation free; 4 and 6 spill (measured slower) */
  for (int j = 0; j < ppr; ++j) {
    {
      const F3 vp = F3{g0.a.x, g0.a.y, g0.b.x};
      const int code = icp_back(P.dist2_thr, P.cos_thr, pd0, vp, F3{g0.b.y, g0.c.x, g0.c.y}, acc2);
      if (DEBUG) {
        const int pk = p0 + (j - 2) * pstep;
        if (P.corr != nullptr && j >= 2 && pk < P.npix) P.corr[pk] = code;
      }
    }
    IcpPend pdn;
    Rec3 gn = g0; /* dead values: the predicated gather overwrites them when the pixel projects into the image */
    const float nx_c = s0.b.y;
    const F3 vc = F3{s0.a.x, s0.a.y, s0.b.x};
    icp_front<DEBUG>(P.g, vc, F3{nx_c, s0.c.x, s0.c.y}, pose, prvb, pdn, gn);
    Rec3 sn = s0; /* dead as well: the registers of the record front(j) has just consumed */
    ld_rec_stream(j + 2, nj, sp, plane_bytes, sn); /* streaming record of pixel j+2 */
    sp += pstep;
    s0 = s1;
    s1 = sn;
    pd0 = pd1;
    g0 = g1;
    pd1 = pdn;
    g1 = gn;
  }
#pragma unroll
  for (int t = 0; t < 2; ++t) { /* drain: pixels ppr-2 and ppr-1 */
    const F3 vp = F3{g0.a.x, g0.a.y, g0.b.x};
    const int code = icp_back(P.dist2_thr, P.cos_thr, pd0, vp, F3{g0.b.y, g0.c.x, g0.c.y}, acc2);
    if (DEBUG) {
      const int jj = P.ppr - 2 + t;
      const int pk = p0 + jj * pstep;
      if (P.corr != nullptr && jj >= 0 && pk < P.npix) P.corr[pk] = code;
    }
    pd0 = pd1;
    g0 = g1;
  }
  float acc[32];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    acc[2 * k] = acc2[k].x;
    acc[2 * k + 1] = acc2[k].y;
  }
  butterfly_step<16>(acc, lane);
  butterfly_step<8>(acc, lane);
  butterfly_step<4>(acc, lane);
  butterfly_step<2>(acc, lane);
  butterfly_step<1>(acc, lane);
  const float* part = P.partials + (size_t)pair * P.max_runs * 32;
  if (has_run) P.partials[((size_t)pair * P.max_runs + run) * 32 + lane] = acc[0];
  __threadfence(); /* publish this run's partial before taking a ticket */
  if (LAST_CTA) {
    __syncthreads(); /* every run of this CTA is published */
    if (threadIdx.x == 0) s_ticket = atomicAdd(P.tickets + pair, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return; /* CTA-uniform */
    /* last CTA of this pair: its warps share the eight chains (chain c = runs c, c+8, ... ascending, in
     * double), then warp 0 adds the eight chains in order -- the order of the specification */
    __threadfence();
    /* chains warp, warp + WARPS, ... (< 8) belong to this warp */
    constexpr int CPW = (8 + YK_ICP_WARPS - 1) / YK_ICP_WARPS;
    double cs[CPW];
#pragma unroll
    for (int k = 0; k < CPW; ++k) cs[k] = 0.0;
    int r = 0;
    constexpr int DEPTH = 32 / CPW; /* CPW x DEPTH = 32 loads in flight per lane */
    for (; r + 8 * DEPTH <= P.nruns; r += 8 * DEPTH) {
      float v[CPW][DEPTH];
#pragma unroll
      for (int u = 0; u < DEPTH; ++u)
#pragma unroll
        for (int k = 0; k < CPW; ++k) {
          const int c = warp + k * YK_ICP_WARPS;
          v[k][u] = c < 8 ? __ldcg(part + (size_t)(r + 8 * u + c) * 32 + lane) : 0.0f;
        }
#pragma unroll
      for (int u = 0; u < DEPTH; ++u)
#pragma unroll
        for (int k = 0; k < CPW; ++k)
          if (warp + k * YK_ICP_WARPS < 8) cs[k] = cs[k] + (double)v[k][u];
    }
    for (; r < P.nruns; r += 8) {
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int c = warp + k * YK_ICP_WARPS;
        if (c < 8 && r + c < P.nruns) cs[k] = cs[k] + (double)__ldcg(part + (size_t)(r + c) * 32 + lane);
      }
    }
#pragma unroll
    for (int k = 0; k < CPW; ++k) {
      const int c = warp + k * YK_ICP_WARPS;
      if (c < 8) s_chain[LAST_CTA ? c : 0][lane] = cs[k];
    }
    __syncthreads();
    if (warp != 0) return;
    double t = s_chain[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = t + s_chain[LAST_CTA ? w : 0][lane];
    s_tot[0][lane] = t;
    P.sums[pair * 32 + lane] = t;
    if (lane == 0) P.tickets[pair] = 0u; /* ready for the next launch */
    __syncwarp();
    if (P.do_solve) {
      if (!solve_update_warp(s_tot[0], P.min_inliers, P.pose_d + pair * 12, P.pose_f_out + pair * 12, lane)) {
        if (lane == 0) P.pair_status[pair] |= YOUTH_STATUS_LOST;
      }
    }
    return;
  }
  unsigned int ticket = 0;
  if (lane == 0) ticket = atomicAdd(P.tickets + pair, 1u);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != (unsigned int)(P.nruns - 1)) return;

  /* last run of this pair: fixed-order cross-run reduction in double.  Chain w (0..7) adds
   * runs w, w+8, ... in ascending order; lane = slot; up to 32 loads in flight. */
  __threadfence();
  double ch[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) ch[w] = 0.0;
  int r = 0;
  for (; r + 64 <= P.nruns; r += 64) { /* the accumulators are dead here: 64 loads in flight per lane */
    float v[64];
#pragma unroll
    for (int u = 0; u < 64; ++u) v[u] = __ldcg(part + (size_t)(r + u) * 32 + lane);
#pragma unroll
    for (int u = 0; u < 64; ++u) ch[u & 7] = ch[u & 7] + (double)v[u];
  }
  for (; r + 16 <= P.nruns; r += 16) {
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldcg(part + (size_t)(r + u) * 32 + lane);
#pragma unroll
    for (int u = 0; u < 16; ++u) ch[u & 7] = ch[u & 7] + (double)v[u];
  }
  for (; r < P.nruns; r += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (r + u < P.nruns) ? __ldcg(part + (size_t)(r + u) * 32 + lane) : 0.0f;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r + u < P.nruns) ch[u] = ch[u] + (double)v[u];
  }
  double t = ch[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) t = t + ch[w];
  s_tot[warp][lane] = t;
  P.sums[pair * 32 + lane] = t;
  if (lane == 0) P.tickets[pair] = 0u; /* ready for the next launch */
  __syncwarp();
  if (P.do_solve) {
    if (!solve_update_warp(s_tot[warp], P.min_inliers, P.pose_d + pair * 12, P.pose_f_out + pair * 12, lane)) {
      if (lane == 0) P.pair_status[pair] |= YOUTH_STATUS_LOST;
    }
  }
}

/* ------------------------------------------------------------------ k_icp_fused */

/* The whole coarse-to-fine schedule of a few pairs in ONE launch (live frames, frame-to-model tracking,
 * pair groups): the CTAs of a pair stay resident, and after every iteration they wait on a per-pair
 * generation counter that the pair's last CTA advances once it has reduced, solved and published the
 * pose.  Same sweep, same reduction order, same solve as k_icp<false, true>; the 19 kernel boundaries are
 * replaced by a release/acquire hand-off.
 * MEASURED ON B200: bit-identical, and SLOWER than one (graph-captured) launch per iteration -- frame-to-model
 * 2995 vs 3374 frames/s, pair groups of 2..11 pairs on 1..5 streams 12.2..43 ms vs 8.3 ms per 300 frames: an
 * iteration of a few pairs is a chain of dependent L2 round trips (pose, first records, partials, ticket,
 * reduction, solve) of ~10-15 us either way, and the polled counter adds to it.  Opt-in (YOUTH_ICP_FUSED=1),
 * kept because it is the building block of an L2-resident multi-pair schedule (what comes next).
 * Requirements the host enforces: every CTA of the grid is resident at the same time (cooperative launch,
 * grid <= occupancy x SMs); gen[] and tickets[] are zero at launch and are left zero. */
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(32 * YK_ICP_WARPS, 3) k_icp_fused(const __grid_constant__ IcpFusedParams P) {
  __shared__ double s_tot[32];
  __shared__ double s_chain[8][32];
  __shared__ unsigned int s_ticket;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sq = blockIdx.z, fi = P.f0 + blockIdx.y; /* grid = (CTAs of a pair at the finest level, frames, sequences) */
  const int pair = sq * P.ring.n + fi;
  if (P.seq_count[sq] + fi == 0) return; /* first frame of a sequence (all CTAs of the pair leave together) */
  const int cur_slot = ring_slot(P.ring, fi);
  const int prev_slot = (cur_slot + P.ring.R - 1) % P.ring.R;
  const size_t stream_base = (size_t)sq * P.ring.R;
  const float* pose_g = P.pose_f + pair * 12;
  const float* part = P.partials + (size_t)pair * P.max_runs * 32;
  unsigned int done = 0; /* iterations of this pair that must be complete before this CTA's next sweep */
  const Rec3 zrec = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};

  for (int li = 0; li < P.nlv; ++li) {
    const IcpFusedLevel& L = P.lv[li];
    const int ctas = (L.nruns + YK_ICP_WARPS - 1) / YK_ICP_WARPS;
    if ((int)blockIdx.x >= ctas) { /* a coarse level needs fewer CTAs: the others join at the first level that has work for them */
      done += (unsigned int)L.iters;
      continue;
    }
    const bool last_level = li == P.nlv - 1;
    const int run = blockIdx.x * YK_ICP_WARPS + warp;
    const bool has_run = run < L.nruns;
    const size_t npx = (size_t)L.npix;
    const float2* __restrict__ cur = L.maps + (stream_base + cur_slot) * 3 * npx;
    const float2* __restrict__ prv = L.model != nullptr ? L.model + (size_t)sq * 3 * npx : L.maps + (stream_base + prev_slot) * 3 * npx;
    const LevelGeom g = L.g;
    const int npx_i = L.npix;
    const int p0 = run * 32 + lane;
    const int pstep = 32 * L.nruns;
    const int ppr = has_run ? L.ppr : 0;
    int nj = p0 < npx_i ? (npx_i - p0 + pstep - 1) / pstep : 0;
    nj = nj < ppr ? nj : ppr;
    const long long plane_bytes = (long long)npx * (long long)sizeof(float2);
    const RecBase prvb = {prv, plane_bytes};

    for (int it = 0; it < L.iters; ++it) {
      /* the pose of iteration `done` (identity from k_ingest when done == 0) */
      if (done > 0) {
        if (threadIdx.x == 0) {
          while (ld_acquire_gpu(P.gen + pair) < done) {
          }
        }
        __syncthreads();
      }
      float pose[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) pose[k] = __ldcg(pose_g + k);

      float2 acc2[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc2[k] = make_float2(0.0f, 0.0f);
      const float2* sp = cur + p0;
      Rec3 s0 = zrec, s1 = zrec;
      ld_rec_stream(0, nj, sp, plane_bytes, s0);
      sp += pstep;
      ld_rec_stream(1, nj, sp, plane_bytes, s1);
      sp += pstep;
      IcpPend pd0, pd1;
      pd0.tx = pd0.ty = pd0.tz = pd0.rnx = pd0.rny = pd0.rnz = 0.0f;
      pd0.q = YOUTH_REJ_CUR_INVALID;
      pd1 = pd0;
      Rec3 g0 = zrec, g1 = zrec;
#pragma unroll 2
      for (int j = 0; j < ppr; ++j) { /* the two-deep software pipeline of k_icp */
        icp_back(P.dist2_thr, P.cos_thr, pd0, F3{g0.a.x, g0.a.y, g0.b.x}, F3{g0.b.y, g0.c.x, g0.c.y}, acc2);
        IcpPend pdn;
        Rec3 gn = g0;
        icp_front<false>(g, F3{s0.a.x, s0.a.y, s0.b.x}, F3{s0.b.y, s0.c.x, s0.c.y}, pose, prvb, pdn, gn);
        Rec3 sn = s0;
        ld_rec_stream(j + 2, nj, sp, plane_bytes, sn);
        sp += pstep;
        s0 = s1;
        s1 = sn;
        pd0 = pd1;
        g0 = g1;
        pd1 = pdn;
        g1 = gn;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        icp_back(P.dist2_thr, P.cos_thr, pd0, F3{g0.a.x, g0.a.y, g0.b.x}, F3{g0.b.y, g0.c.x, g0.c.y}, acc2);
        pd0 = pd1;
        g0 = g1;
      }
      float acc[32];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        acc[2 * k] = acc2[k].x;
        acc[2 * k + 1] = acc2[k].y;
      }
      butterfly_step<16>(acc, lane);
      butterfly_step<8>(acc, lane);
      butterfly_step<4>(acc, lane);
      butterfly_step<2>(acc, lane);
      butterfly_step<1>(acc, lane);
      if (has_run) P.partials[((size_t)pair * P.max_runs + run) * 32 + lane] = acc[0];
      __threadfence(); /* publish this run's partial before the CTA takes its ticket */
      __syncthreads();
      if (threadIdx.x == 0) s_ticket = atomicAdd(P.tickets + pair, 1u);
      __syncthreads();
      done += 1;
      if (s_ticket != (unsigned int)(ctas - 1)) continue; /* CTA-uniform */

      /* last CTA of the pair in this iteration: cross-run reduction in the order of the specification
       * (chain c = runs c, c+8, ... ascending, in double; then the eight chains in order), shared by
       * the warps exactly as in k_icp<false, true> */
      __threadfence();
      constexpr int CPW = (8 + YK_ICP_WARPS - 1) / YK_ICP_WARPS;
      double cs[CPW];
#pragma unroll
      for (int k = 0; k < CPW; ++k) cs[k] = 0.0;
      int r = 0;
      constexpr int DEPTH = 32 / CPW;
      for (; r + 8 * DEPTH <= L.nruns; r += 8 * DEPTH) {
        float v[CPW][DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u)
#pragma unroll
          for (int k = 0; k < CPW; ++k) {
            const int c = warp + k * YK_ICP_WARPS;
            v[k][u] = c < 8 ? __ldcg(part + (size_t)(r + 8 * u + c) * 32 + lane) : 0.0f;
          }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u)
#pragma unroll
          for (int k = 0; k < CPW; ++k)
            if (warp + k * YK_ICP_WARPS < 8) cs[k] = cs[k] + (double)v[k][u];
      }
      for (; r < L.nruns; r += 8) {
#pragma unroll
        for (int k = 0; k < CPW; ++k) {
          const int c = warp + k * YK_ICP_WARPS;
          if (c < 8 && r + c < L.nruns) cs[k] = cs[k] + (double)__ldcg(part + (size_t)(r + c) * 32 + lane);
        }
      }
#pragma unroll
      for (int k = 0; k < CPW; ++k) {
        const int c = warp + k * YK_ICP_WARPS;
        if (c < 8) s_chain[c][lane] = cs[k];
      }
      __syncthreads();
      if (warp == 0) {
        double t = s_chain[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) t = t + s_chain[w][lane];
        s_tot[lane] = t;
        P.sums[pair * 32 + lane] = t;
        if (lane == 0) P.tickets[pair] = 0u; /* every CTA of this iteration has taken its ticket */
        __syncwarp();
        if (!solve_update_warp(s_tot, P.min_inliers, P.pose_d + pair * 12, P.pose_f + pair * 12, lane)) {
          if (lane == 0) P.pair_status[pair] |= YOUTH_STATUS_LOST;
        }
        __syncwarp();
        if (lane == 0) {
          /* hand-off: the ticket reset, the sums and the pose are ordered before the new generation.  After the
           * final iteration nobody waits any more (every CTA of the grid row works at the finest level and has
           * passed all earlier waits before the last ticket was taken): leave the counter at zero. */
          const bool final_iteration = last_level && it == L.iters - 1;
          __threadfence();
          st_release_gpu(P.gen + pair, final_iteration ? 0u : done);
        }
      }
      /* the other warps of this CTA go on to the next iteration's wait like every other CTA */
    }
  }
}

/* ------------------------------------------------------------------ k_compose */

/* One CTA per sequence.  The pose chain is inherently sequential (and its operation order is
 * part of the specification), so one thread multiplies, but the relative poses are staged
 * into shared memory by the whole CTA first and the results are written back coalesced. */
#define YK_COMPOSE_CHUNK 64
__global__ void __launch_bounds__(128) k_compose(const __grid_constant__ ComposeParams P) {
  __shared__ double s_rel[YK_COMPOSE_CHUNK][12];
  __shared__ float s_out[YK_COMPOSE_CHUNK][12];
  __shared__ uint32_t s_st[YK_COMPOSE_CHUNK];
  __shared__ int s_in[YK_COMPOSE_CHUNK];
  __shared__ double s_w[12];
  __shared__ int s_inl;
  const int s = blockIdx.x, tid = threadIdx.x;
  if (s >= P.ring.S) return;
  const int c0 = P.seq_count[s];
  if (tid < 12) s_w[tid] = P.world[s * 12 + tid];
  if (tid == 0) s_inl = 0;
  for (int base = 0; base < P.ring.n; base += YK_COMPOSE_CHUNK) {
    const int cn = min(YK_COMPOSE_CHUNK, P.ring.n - base);
    __syncthreads();
    for (int k = tid; k < cn * 12; k += 128) s_rel[k / 12][k % 12] = P.pose_d[(size_t)(s * P.ring.n + base) * 12 + k];
    for (int k = tid; k < cn; k += 128) { /* status + inlier count of every pair, staged like the poses */
      s_st[k] = P.pair_status[s * P.ring.n + base + k];
      s_in[k] = (int)P.sums[(size_t)(s * P.ring.n + base + k) * 32 + YOUTH_SUMS_COUNT];
    }
    __syncthreads();
    if (tid == 0) {
      double Wd[12];
      for (int k = 0; k < 12; ++k) Wd[k] = s_w[k];
      int inl = s_inl;
      for (int i = 0; i < cn; ++i) {
        const int fi = c0 + base + i;
        uint32_t st;
        if (fi == 0) {
          for (int k = 0; k < 12; ++k) Wd[k] = (k == 0 || k == 5 || k == 10) ? 1.0 : 0.0;
          st = YOUTH_STATUS_FIRST;
          inl = 0;
        } else {
          const double* r = s_rel[i];
          double T[12];
#pragma unroll
          for (int a2 = 0; a2 < 3; ++a2) {
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2)
              T[4 * a2 + b2] = (Wd[4 * a2] * r[b2] + Wd[4 * a2 + 1] * r[4 + b2]) + Wd[4 * a2 + 2] * r[8 + b2];
            T[4 * a2 + 3] = ((Wd[4 * a2] * r[3] + Wd[4 * a2 + 1] * r[7]) + Wd[4 * a2 + 2] * r[11]) + Wd[4 * a2 + 3];
          }
#pragma unroll
          for (int k = 0; k < 12; ++k) Wd[k] = T[k];
          st = s_st[i];
          inl = s_in[i];
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) s_out[i][k] = (float)Wd[k];
        s_st[i] = st;
      }
      for (int k = 0; k < 12; ++k) s_w[k] = Wd[k];
      s_inl = inl;
    }
    __syncthreads();
    for (int k = tid; k < cn * 12; k += 128) {
      const int fi = c0 + base + k / 12;
      if (fi < P.cap) P.traj[((size_t)s * P.cap + fi) * 12 + k % 12] = s_out[k / 12][k % 12];
    }
    for (int k = tid; k < cn; k += 128) {
      const int fi = c0 + base + k;
      if (fi < P.cap) P.traj_status[(size_t)s * P.cap + fi] = s_st[k];
    }
  }
  __syncthreads();
  if (tid < 12) {
    P.world[s * 12 + tid] = s_w[tid];
    if (P.world_f != nullptr) P.world_f[s * 12 + tid] = (float)s_w[tid];
  }
  if (tid == 0) {
    if (P.last_status != nullptr) P.last_status[s] = s_st[(P.ring.n - 1) % YK_COMPOSE_CHUNK];
    P.seq_count[s] = c0 + P.ring.n;
    P.last_inliers[s] = s_inl;
    if (s == 0) *P.head = (*P.head + P.ring.n) % P.ring.R; /* every kernel of this group has already run */
  }
}


/* ------------------------------------------------------------------ launchers */

void yk_launch_icp(bool last_cta, dim3 grid, cudaStream_t st, const IcpParams& p) {
  if (last_cta) k_icp<false, true><<<grid, 32 * YK_ICP_WARPS, 0, st>>>(p);
  else k_icp<false, false><<<grid, 32 * YK_ICP_WARPS, 0, st>>>(p);
}

void yk_launch_icp_debug(dim3 grid, cudaStream_t st, const IcpParams& p) { k_icp<true, false><<<grid, 32 * YK_ICP_WARPS, 0, st>>>(p); }

cudaError_t yk_icp_fused_ctas_per_sm(int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_icp_fused, 32 * YK_ICP_WARPS, 0);
}

cudaError_t yk_launch_icp_fused(dim3 grid, cudaStream_t st, bool cooperative, const IcpFusedParams& p) {
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = grid;
  lc.blockDim = dim3(32 * YK_ICP_WARPS);
  lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; /* every CTA resident at once, or the launch waits its turn */
  at[0].val.cooperative = 1;
  lc.attrs = at;
  lc.numAttrs = cooperative ? 1 : 0;
  return cudaLaunchKernelEx(&lc, k_icp_fused, p);
}

void yk_launch_compose(int sequences, cudaStream_t st, const ComposeParams& p) { k_compose<<<sequences, 128, 0, st>>>(p); }

void yk_launch_rcp_check(cudaStream_t st, uint32_t lo_bits, uint32_t hi_bits, unsigned long long* d_mismatches) {
  k_rcp_check<<<148 * 8, 256, 0, st>>>(lo_bits, hi_bits, d_mismatches);
}
