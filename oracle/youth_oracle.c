/*
 * youth_oracle.c -- CPU oracle of the frame-to-frame depth tracking path (C99).
 * TEST INFRASTRUCTURE ONLY -- see youth_oracle.h for who may call this and for the
 * "parity unpinned" statement.  Build with -O2 -ffp-contract=off (oracle/Makefile).
 *
 * Stage map:
 *   stage 1  yo_bilateral, yo_pyrdown   depth ingest (replaces the int16 -> float metres
 *                                       conversion of reference SLAM.cpp:133-134,153-155)
 *   stage 2  yo_vertex_normal           back-projection per reference viewerModule.c:341-345
 *   stage 3  icp_pixel                  projective association + point-to-plane residual/Jacobian
 *   stage 4  yo_icp_sums                fixed-order reduction of the 27 (+2, +3 duplicate) sums
 *   stage 5  yo_solve_update            6x6 solve + SE(3) update
 */
#define _POSIX_C_SOURCE 200809L
#include "youth_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ config */

void yo_default_config(yo_config* c) {
  memset(c, 0, sizeof(*c));
  c->width = 640;  /* Camera.width  (astra_orb_slam3_rgbd.yaml:19) */
  c->height = 480; /* Camera.height (:20) */
  c->fx = 570.3f;  /* Camera.fx (:9) == the literal in viewerModule.c:344 */
  c->fy = 570.3f;  /* Camera.fy (:10) */
  c->cx = 320.0f;  /* Camera.cx (:11) == width/2 in viewerModule.c:344 */
  c->cy = 240.0f;  /* Camera.cy (:12) */
  c->depth_factor = 1000.0f; /* DepthMapFactor (:35) */
  c->levels = 3;
  c->iters[0] = 10;
  c->iters[1] = 5;
  c->iters[2] = 4;
  c->iters[3] = 4;
  c->depth_min_mm = 1; /* viewerModule.c:341: valid iff depth > 0 */
  c->depth_max_mm = 10000;
  c->bilateral = 1;
  c->sigma_space_px = 4.5f;
  c->sigma_range_mm = 30.0f;
  c->dist_thresh_m = 0.10f;
  c->cos_thresh = 0.93969262f; /* cos(20 deg) */
  c->min_inliers = 100;
  c->icp_ppt = 64;
}

int yo_level_geometry(const yo_config* c, int level, yo_level* g) {
  if (level < 0 || level >= c->levels || level >= YO_MAX_LEVELS) return 0;
  int w = c->width, h = c->height;
  float fx = c->fx, fy = c->fy, cx = c->cx, cy = c->cy;
  for (int l = 0; l < level; ++l) {
    w /= 2;
    h /= 2;
    fx = fx * 0.5f;
    fy = fy * 0.5f;
    /* a level-(l+1) pixel is the mean of a 2x2 block whose centre is at 2X+0.5 */
    cx = (cx - 0.5f) * 0.5f;
    cy = (cy - 0.5f) * 0.5f;
  }
  g->w = w;
  g->h = h;
  g->fx = fx;
  g->fy = fy;
  g->cx = cx;
  g->cy = cy;
  return 1;
}

yo_frame* yo_frame_alloc(const yo_config* c) {
  yo_frame* f = (yo_frame*)calloc(1, sizeof(yo_frame));
  for (int l = 0; l < c->levels; ++l) {
    yo_level g;
    yo_level_geometry(c, l, &g);
    size_t n = (size_t)g.w * g.h;
    f->depth[l] = (float*)calloc(n, sizeof(float));
    f->pyrcnt[l] = (uint8_t*)calloc(n, 1);
    f->vmap[l] = (float*)calloc(n * 4, sizeof(float));
    f->nmap[l] = (float*)calloc(n * 4, sizeof(float));
  }
  return f;
}

void yo_frame_free(yo_frame* f) {
  if (!f) return;
  for (int l = 0; l < YO_MAX_LEVELS; ++l) {
    free(f->depth[l]);
    free(f->pyrcnt[l]);
    free(f->vmap[l]);
    free(f->nmap[l]);
  }
  free(f);
}

/* ------------------------------------------------------------------ stage 1 */

static int range_cut(const yo_config* c) {
  int cut = (int)(3.0f * c->sigma_range_mm);
  if (cut > YO_RANGE_LUT_MAX - 2) cut = YO_RANGE_LUT_MAX - 2;
  if (cut < 0) cut = 0;
  return cut;
}

/* weight tables: the only libm use, and it depends on the config alone */
static void bilateral_tables(const yo_config* c, float ws[7][7], float* wr, int cut) {
  double ss = (double)c->sigma_space_px, sr = (double)c->sigma_range_mm;
  for (int dy = -3; dy <= 3; ++dy)
    for (int dx = -3; dx <= 3; ++dx)
      ws[dy + 3][dx + 3] = (float)exp(-(double)(dx * dx + dy * dy) / (2.0 * ss * ss));
  for (int i = 0; i <= cut; ++i) wr[i] = (float)exp(-((double)i * (double)i) / (2.0 * sr * sr));
  wr[cut + 1] = 0.0f;
}

static int raw_valid(const yo_config* c, int d) { return d >= c->depth_min_mm && d <= c->depth_max_mm; }

void yo_bilateral(const yo_config* c, const uint16_t* raw, float* out) {
  const int W = c->width, H = c->height;
  if (!c->bilateral) {
    for (int i = 0; i < W * H; ++i) out[i] = raw_valid(c, raw[i]) ? (float)raw[i] : 0.0f;
    return;
  }
  float ws[7][7];
  float wr[YO_RANGE_LUT_MAX];
  const int cut = range_cut(c);
  bilateral_tables(c, ws, wr, cut);
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      const int dc = raw[y * W + x];
      if (!raw_valid(c, dc)) {
        out[y * W + x] = 0.0f;
        continue;
      }
      float sw = 0.0f, swd = 0.0f;
      for (int dy = -3; dy <= 3; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -3; dx <= 3; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          const int dk = raw[yy * W + xx];
          if (!raw_valid(c, dk)) continue;
          int diff = dk - dc;
          if (diff < 0) diff = -diff;
          if (diff > cut) continue;
          const float w = ws[dy + 3][dx + 3] * wr[diff];
          sw = sw + w;
          swd = fmaf(w, (float)dk, swd); /* specified fused multiply-add */
        }
      }
      out[y * W + x] = swd / sw; /* centre tap has weight ws[3][3]*wr[0] = 1, so sw >= 1 */
    }
  }
}

void yo_pyrdown(const yo_config* c, int w, int h, const float* src, float* dst, uint8_t* cnt) {
  const int w2 = w / 2, h2 = h / 2;
  const float thr = 3.0f * c->sigma_range_mm;
  for (int y = 0; y < h2; ++y) {
    for (int x = 0; x < w2; ++x) {
      /* block samples in scan order (0,0) (1,0) (0,1) (1,1) */
      float s[4];
      s[0] = src[(2 * y) * w + 2 * x];
      s[1] = src[(2 * y) * w + 2 * x + 1];
      s[2] = src[(2 * y + 1) * w + 2 * x];
      s[3] = src[(2 * y + 1) * w + 2 * x + 1];
      float centre = 0.0f;
      for (int k = 0; k < 4; ++k)
        if (s[k] > 0.0f) {
          centre = s[k];
          break;
        }
      float sum = 0.0f;
      int n = 0;
      if (centre > 0.0f) {
        for (int k = 0; k < 4; ++k) {
          if (s[k] > 0.0f && fabsf(s[k] - centre) <= thr) {
            sum = sum + s[k];
            ++n;
          }
        }
      }
      dst[y * w2 + x] = n ? sum / (float)n : 0.0f;
      if (cnt) cnt[y * w2 + x] = (uint8_t)n;
    }
  }
}

/* ------------------------------------------------------------------ stage 2 */

void yo_vertex_normal(const yo_config* c, int level, const float* depth, float* vmap, float* nmap) {
  yo_level g;
  yo_level_geometry(c, level, &g);
  const int W = g.w, H = g.h;
  for (int v = 0; v < H; ++v) {
    for (int u = 0; u < W; ++u) {
      const float d = depth[v * W + u];
      float* o = vmap + 4 * (size_t)(v * W + u);
      if (d > 0.0f) {
        /* viewerModule.c:343-345: z = d/1000; x = (u - cx) * z / fx; y = (v - cy) * z / fy
         * (the viewer's sign flip at :357 is a GL display convention and is not carried) */
        const float z = d / c->depth_factor;
        o[0] = ((float)u - g.cx) * z / g.fx;
        o[1] = ((float)v - g.cy) * z / g.fy;
        o[2] = z;
        o[3] = 1.0f;
      } else {
        o[0] = o[1] = o[2] = o[3] = 0.0f;
      }
    }
  }
  yo_normals_from_vertices(W, H, vmap, nmap);
}

/* normal = normalize((V(u+1,v) - V) x (V(u,v+1) - V)); valid iff the three vertices are */
void yo_normals_from_vertices(int W, int H, const float* vmap, float* nmap) {
  for (int v = 0; v < H; ++v) {
    for (int u = 0; u < W; ++u) {
      float* o = nmap + 4 * (size_t)(v * W + u);
      o[0] = o[1] = o[2] = o[3] = 0.0f;
      if (u + 1 >= W || v + 1 >= H) continue;
      const float* p = vmap + 4 * (size_t)(v * W + u);
      const float* px = vmap + 4 * (size_t)(v * W + u + 1);
      const float* py = vmap + 4 * (size_t)((v + 1) * W + u);
      if (p[3] == 0.0f || px[3] == 0.0f || py[3] == 0.0f) continue;
      const float ax = px[0] - p[0], ay = px[1] - p[1], az = px[2] - p[2];
      const float bx = py[0] - p[0], by = py[1] - p[1], bz = py[2] - p[2];
      const float nx = ay * bz - az * by;
      const float ny = az * bx - ax * bz;
      const float nz = ax * by - ay * bx;
      const float len2 = (nx * nx + ny * ny) + nz * nz;
      if (!(len2 > 1e-24f)) continue;
      const float inv = 1.0f / sqrtf(len2);
      o[0] = nx * inv;
      o[1] = ny * inv;
      o[2] = nz * inv;
      o[3] = 1.0f;
    }
  }
}

void yo_preprocess(const yo_config* c, const uint16_t* raw, yo_frame* f) {
  yo_bilateral(c, raw, f->depth[0]);
  memset(f->pyrcnt[0], 0, (size_t)c->width * c->height);
  for (int l = 1; l < c->levels; ++l) {
    yo_level g;
    yo_level_geometry(c, l - 1, &g);
    yo_pyrdown(c, g.w, g.h, f->depth[l - 1], f->depth[l], f->pyrcnt[l]);
  }
  for (int l = 0; l < c->levels; ++l) yo_vertex_normal(c, l, f->depth[l], f->vmap[l], f->nmap[l]);
}

/* ------------------------------------------------------------------ stage 3 */

/* slot layout of the 32 sums: 16 pairs, pair p = X[PA[p]] * (X[PB[p]], X[PB[p]+1]) with
 * X = (J0..J5, r, 1); chosen so that the device can form every pair with one packed
 * fma.rn.f32x2.  Slots 6, 16, 22 duplicate A10, A32, A54 and are ignored by the solve. */
static const int YO_PA[16] = {0, 0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, -1};
static const int YO_PB[16] = {0, 2, 4, 0, 2, 4, 2, 4, 2, 4, 4, 4, 0, 2, 4, -1};
/* slot of A[i][j], i <= j */
static const int YO_SLOT_A[6][6] = {{0, 1, 2, 3, 4, 5},      {1, 7, 8, 9, 10, 11},    {2, 8, 12, 13, 14, 15},
                                    {3, 9, 13, 17, 18, 19},  {4, 10, 14, 18, 20, 21}, {5, 11, 15, 19, 21, 23}};
#define YO_SLOT_B0 24
#define YO_SLOT_RR 30
#define YO_SLOT_COUNT 31

/* One pixel of the current frame.  Returns the matched previous-frame pixel index or a
 * negative reject code; on a match ADDS its 32 terms into acc[] with one fused
 * multiply-add per slot: acc[k] = fma(a, b, acc[k]).
 * J = [ (T v) x n' , n' ],  r = n' . (v' - T v).
 * Every fmaf() below is a single-rounding IEEE fused multiply-add and is part of the
 * specification (the device issues fma.rn.f32 / fma.rn.f32x2 at the same places);
 * everything else is one rounding per written operation. */
static int icp_pixel(const yo_level* g, float dist2_thr, float cos_thr, const float* vc4,
                     const float* nc4, const float* vprev, const float* nprev, const float* P,
                     float* acc) {
  if (vc4[3] == 0.0f || nc4[3] == 0.0f) return YO_REJ_CUR_INVALID;
  const float x = vc4[0], y = vc4[1], z = vc4[2];
  const float tx = fmaf(P[0], x, fmaf(P[1], y, fmaf(P[2], z, P[3])));
  const float ty = fmaf(P[4], x, fmaf(P[5], y, fmaf(P[6], z, P[7])));
  const float tz = fmaf(P[8], x, fmaf(P[9], y, fmaf(P[10], z, P[11])));
  if (!(tz >= FLT_MIN)) return YO_REJ_BEHIND; /* in front of the camera: a positive normal float */
  const float iz = 1.0f / tz;
  const float ur = fmaf(tx * g->fx, iz, g->cx + 0.5f);
  const float vr = fmaf(ty * g->fy, iz, g->cy + 0.5f);
  if (!(ur >= 0.0f && ur < (float)g->w && vr >= 0.0f && vr < (float)g->h)) return YO_REJ_OUT_OF_IMAGE;
  const int ui = (int)ur, vi = (int)vr; /* nearest pixel: floor(u + 0.5) */
  const int q = vi * g->w + ui;
  const float* vp = vprev + 4 * (size_t)q;
  const float* np = nprev + 4 * (size_t)q;
  if (vp[3] == 0.0f || np[3] == 0.0f) return YO_REJ_PREV_INVALID;
  const float dx = vp[0] - tx, dy = vp[1] - ty, dz = vp[2] - tz;
  const float dist2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  if (!(dist2 <= dist2_thr)) return YO_REJ_DISTANCE;
  const float nx = nc4[0], ny = nc4[1], nz = nc4[2];
  const float rnx = fmaf(P[2], nz, fmaf(P[1], ny, P[0] * nx));
  const float rny = fmaf(P[6], nz, fmaf(P[5], ny, P[4] * nx));
  const float rnz = fmaf(P[10], nz, fmaf(P[9], ny, P[8] * nx));
  const float cosang = fmaf(rnz, np[2], fmaf(rny, np[1], rnx * np[0]));
  if (!(cosang >= cos_thr)) return YO_REJ_ANGLE;
  float X[8];
  X[6] = fmaf(np[2], dz, fmaf(np[1], dy, np[0] * dx)); /* r */
  X[0] = fmaf(ty, np[2], -(tz * np[1]));
  X[1] = fmaf(tz, np[0], -(tx * np[2]));
  X[2] = fmaf(tx, np[1], -(ty * np[0]));
  X[3] = np[0];
  X[4] = np[1];
  X[5] = np[2];
  X[7] = 1.0f;
  for (int p = 0; p < 15; ++p) {
    acc[2 * p] = fmaf(X[YO_PA[p]], X[YO_PB[p]], acc[2 * p]);
    acc[2 * p + 1] = fmaf(X[YO_PA[p]], X[YO_PB[p] + 1], acc[2 * p + 1]);
  }
  acc[30] = fmaf(X[6], X[6], acc[30]);
  acc[31] = fmaf(X[7], X[7], acc[31]);
  return q;
}

/* ------------------------------------------------------------------ stage 4 */

/* Fixed-order reduction (identical tree on the device):
 *   nruns  = ceil(npix / (32*ppr)) runs per frame, ppr = max(1, icp_ppt >> 2*level);
 *   lane l of run k adds the pixels (row-major linear index) j*(32*nruns) + 32*k + l,
 *            j = 0..ppr-1, in that order (all runs sweep the image together), into 32
 *            float accumulators that start at +0, one fused multiply-add per slot (rejected
 *            pixels add nothing);
 *   run    : the 32 lanes are combined by the pairwise tree with strides 16, 8, 4, 2, 1;
 *   frame  : run partials accumulated in double: chain w (0..7) adds runs w, w+8, ...
 *            starting from 0.0, then the 8 chains are added in order starting from chain 0. */
static int run_ppr(const yo_config* c, int level) {
  int ppr = c->icp_ppt >> (2 * level);
  return ppr < 1 ? 1 : ppr;
}

void yo_icp_sums(const yo_config* c, int level, const yo_frame* cur, const yo_frame* prev,
                 const float pose[12], double* sums, int32_t* corr) {
  yo_level g;
  yo_level_geometry(c, level, &g);
  const int npix = g.w * g.h;
  const int ppr = run_ppr(c, level);
  const int T = YO_ICP_LANES * ppr;
  const int nruns = (npix + T - 1) / T;
  const float dist2_thr = c->dist_thresh_m * c->dist_thresh_m;
  const float* vc = cur->vmap[level];
  const float* nc = cur->nmap[level];
  const float* vp = prev->vmap[level];
  const float* np = prev->nmap[level];

  float* partial = (float*)malloc(sizeof(float) * YO_SUM_SLOTS * (size_t)nruns);
  for (int run = 0; run < nruns; ++run) {
    float acc[YO_ICP_LANES][YO_SUM_SLOTS];
    memset(acc, 0, sizeof(acc));
    for (int j = 0; j < ppr; ++j) {
      for (int l = 0; l < YO_ICP_LANES; ++l) {
        const int p = j * (YO_ICP_LANES * nruns) + YO_ICP_LANES * run + l;
        if (p >= npix) continue;
        const int q = icp_pixel(&g, dist2_thr, c->cos_thresh, vc + 4 * (size_t)p, nc + 4 * (size_t)p, vp,
                                np, pose, acc[l]);
        if (corr) corr[p] = q;
      }
    }
    for (int k = 0; k < YO_SUM_SLOTS; ++k) {
      float v[YO_ICP_LANES];
      for (int l = 0; l < YO_ICP_LANES; ++l) v[l] = acc[l][k];
      for (int s = 16; s >= 1; s >>= 1)
        for (int l = 0; l < s; ++l) v[l] = v[l] + v[l + s];
      partial[(size_t)run * YO_SUM_SLOTS + k] = v[0];
    }
  }
  for (int k = 0; k < YO_SUM_SLOTS; ++k) {
    double chain[8];
    for (int w = 0; w < 8; ++w) {
      double d = 0.0;
      for (int run = w; run < nruns; run += 8) d = d + (double)partial[(size_t)run * YO_SUM_SLOTS + k];
      chain[w] = d;
    }
    double tot = chain[0];
    for (int w = 1; w < 8; ++w) tot = tot + chain[w];
    sums[k] = tot;
  }
  free(partial);
}

/* ------------------------------------------------------------------ stage 5 */

/* sin(t)/t, (1-cos t)/t^2, (t-sin t)/t^3 as Horner polynomials in t^2 (double, 12 terms):
 * deterministic on any IEEE machine, no libm. */
static void so3_coeffs(double t2, double* A, double* B, double* C) {
  static const double f[28] = {1.0,
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
  double a = 0.0, b = 0.0, c = 0.0;
  for (int k = 11; k >= 0; --k) {
    const double sgn = (k & 1) ? -1.0 : 1.0;
    a = a * t2 + sgn / f[2 * k + 1];
    b = b * t2 + sgn / f[2 * k + 2];
    c = c * t2 + sgn / f[2 * k + 3];
  }
  *A = a;
  *B = b;
  *C = c;
}

static void mat3_mul(const double* a, const double* b, double* o) { /* 3x3 row-major */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = (a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j]) + a[3 * i + 2] * b[6 + j];
}

int yo_solve_update(const yo_config* c, const double* sums, double pose_d[12], float pose_f[12]) {
  if (!(sums[YO_SLOT_COUNT] >= (double)c->min_inliers)) return 0;
  double A[6][6], b[6], L[6][6], yv[6], x[6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) A[i][j] = sums[YO_SLOT_A[i][j]];
  for (int i = 0; i < 6; ++i) b[i] = sums[YO_SLOT_B0 + i];
  double scale = A[0][0];
  for (int i = 1; i < 6; ++i)
    if (A[i][i] > scale) scale = A[i][i];
  /* Cholesky A = L L^T with one reciprocal per column (divisions by the pivot become
   * multiplications), forward substitution in ascending and back substitution in
   * DESCENDING column order -- the order a row-per-lane device solve produces naturally. */
  double inv[6];
  memset(L, 0, sizeof(L));
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
    for (int m = 0; m < j; ++m) d = d - L[j][m] * L[j][m];
    if (!(d > 1e-12 * scale)) return 0;
    L[j][j] = sqrt(d);
    inv[j] = 1.0 / L[j][j];
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
      for (int m = 0; m < j; ++m) s = s - L[i][m] * L[j][m];
      L[i][j] = s * inv[j];
    }
  }
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int m = 0; m < i; ++m) s = s - L[i][m] * yv[m];
    yv[i] = s * inv[i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = yv[i];
    for (int m = 5; m > i; --m) s = s - L[m][i] * x[m];
    x[i] = s * inv[i];
  }
  for (int i = 0; i < 6; ++i)
    if (!(x[i] > -1e6 && x[i] < 1e6)) return 0; /* also rejects NaN */

  /* SE(3) exponential of xi = (w, u):  R = I + A W + B W^2,  V = I + B W + C W^2 */
  const double wx = x[0], wy = x[1], wz = x[2];
  const double t2 = (wx * wx + wy * wy) + wz * wz;
  double Ac, Bc, Cc;
  so3_coeffs(t2, &Ac, &Bc, &Cc);
  const double W[9] = {0.0, -wz, wy, wz, 0.0, -wx, -wy, wx, 0.0};
  double W2[9];
  mat3_mul(W, W, W2);
  double Ri[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    Ri[i] = (id + Ac * W[i]) + Bc * W2[i];
    V[i] = (id + Bc * W[i]) + Cc * W2[i];
  }
  double ti[3];
  for (int i = 0; i < 3; ++i) ti[i] = (V[3 * i] * x[3] + V[3 * i + 1] * x[4]) + V[3 * i + 2] * x[5];
  /* T <- exp(xi) * T */
  double R[9], t[3], Rn[9], tn[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[3 * i + j] = pose_d[4 * i + j];
    t[i] = pose_d[4 * i + 3];
  }
  mat3_mul(Ri, R, Rn);
  for (int i = 0; i < 3; ++i) tn[i] = ((Ri[3 * i] * t[0] + Ri[3 * i + 1] * t[1]) + Ri[3 * i + 2] * t[2]) + ti[i];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) pose_d[4 * i + j] = Rn[3 * i + j];
    pose_d[4 * i + 3] = tn[i];
  }
  for (int i = 0; i < 12; ++i) pose_f[i] = (float)pose_d[i];
  return 1;
}

static void pose_identity(double* p) {
  for (int i = 0; i < 12; ++i) p[i] = 0.0;
  p[0] = p[5] = p[10] = 1.0;
}

uint32_t yo_track_pair(const yo_config* c, const yo_frame* cur, const yo_frame* prev, double rel[12],
                       int32_t* inliers_out) {
  float pf[12];
  double sums[YO_SUM_SLOTS];
  uint32_t status = 0;
  pose_identity(rel); /* every pair starts from the identity: pairs are independent */
  for (int i = 0; i < 12; ++i) pf[i] = (float)rel[i];
  int32_t inl = 0;
  for (int level = c->levels - 1; level >= 0; --level) {
    for (int it = 0; it < c->iters[level]; ++it) {
      yo_icp_sums(c, level, cur, prev, pf, sums, NULL);
      inl = (int32_t)sums[YO_SLOT_COUNT];
      if (!yo_solve_update(c, sums, rel, pf)) status |= YO_STATUS_LOST;
    }
  }
  if (inliers_out) *inliers_out = inl;
  return status;
}

void yo_compose(const double a[12], const double r[12], double o[12]) {
  double tmp[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      tmp[4 * i + j] = (a[4 * i] * r[j] + a[4 * i + 1] * r[4 + j]) + a[4 * i + 2] * r[8 + j];
    tmp[4 * i + 3] = ((a[4 * i] * r[3] + a[4 * i + 1] * r[7]) + a[4 * i + 2] * r[11]) + a[4 * i + 3];
  }
  memcpy(o, tmp, sizeof(tmp));
}

/* ------------------------------------------------------------------ sequence tracker */

struct yo_tracker {
  yo_config cfg;
  yo_frame* fr[2];
  int newest; /* index of the newest frame in fr[] */
  int count;
  double world[12];
  int32_t inliers;
};

yo_tracker* yo_tracker_create(const yo_config* c) {
  yo_tracker* t = (yo_tracker*)calloc(1, sizeof(*t));
  t->cfg = *c;
  t->fr[0] = yo_frame_alloc(c);
  t->fr[1] = yo_frame_alloc(c);
  yo_tracker_reset(t);
  return t;
}

void yo_tracker_destroy(yo_tracker* t) {
  if (!t) return;
  yo_frame_free(t->fr[0]);
  yo_frame_free(t->fr[1]);
  free(t);
}

void yo_tracker_reset(yo_tracker* t) {
  t->newest = 1;
  t->count = 0;
  t->inliers = 0;
  pose_identity(t->world);
}

uint32_t yo_tracker_track(yo_tracker* t, const uint16_t* raw, float pose_out[12]) {
  const int slot = t->newest ^ 1;
  yo_preprocess(&t->cfg, raw, t->fr[slot]);
  uint32_t status;
  if (t->count == 0) {
    status = YO_STATUS_FIRST;
    t->inliers = 0;
  } else {
    double rel[12];
    status = yo_track_pair(&t->cfg, t->fr[slot], t->fr[t->newest], rel, &t->inliers);
    yo_compose(t->world, rel, t->world);
  }
  t->newest = slot;
  t->count++;
  if (pose_out)
    for (int i = 0; i < 12; ++i) pose_out[i] = (float)t->world[i];
  return status;
}

int32_t yo_tracker_last_inliers(const yo_tracker* t) { return t->inliers; }

const yo_frame* yo_tracker_frame(const yo_tracker* t, int which) {
  return t->fr[which ? (t->newest ^ 1) : t->newest];
}

double yo_track_sequence(const yo_config* c, const uint16_t* frames, int n, float* poses_out,
                         uint32_t* status_out) {
  yo_tracker* t = yo_tracker_create(c);
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  const size_t stride = (size_t)c->width * c->height;
  for (int i = 0; i < n; ++i) {
    uint32_t st = yo_tracker_track(t, frames + stride * i, poses_out ? poses_out + 12 * (size_t)i : NULL);
    if (status_out) status_out[i] = st;
  }
  clock_gettime(CLOCK_MONOTONIC, &b);
  yo_tracker_destroy(t);
  return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
