/*
 * youth_ingest.cu -- stages 1-2 of the tracker on sm_100a (see youth_common.cuh for the contract).
 *
 *   k_ingest   uint16 depth -> validity + 7x7 bilateral (smem tile, 128-bit loads, two pixels per thread in packed
 *              FFMA2 lock step) -> level-0 vertex AND normal maps (tile + 1 halo) -> in-tile pyramid + vertex maps
 *              of levels >= 1
 *   k_normals  cross-product normal maps of levels >= 1, one launch
 */
#include "youth_common.cuh"

#ifndef YK_INGEST_MIN_BLOCKS
#define YK_INGEST_MIN_BLOCKS 6 /* resident k_ingest CTAs per SM the register budget is sized for (40 registers, 35 KB smem) */
#endif
#define YK_TILE_W 64
#define YK_TILE_H 16
#define YK_HALO 3
#define YK_SMEM_W (YK_TILE_W + 16) /* 8-pixel (128-bit) aligned halo on both sides */
#define YK_SMEM_H (YK_TILE_H + 1 + 2 * YK_HALO) /* +1: halo row for the fused level-0 normals */
#define YK_SENTINEL 1.0e9f

/* ------------------------------------------------------------------ k_ingest */

__device__ __forceinline__ float pyr_combine(float s0, float s1, float s2, float s3, float thr, int* cnt) {
  float centre = 0.0f;
  if (s0 > 0.0f) centre = s0;
  else if (s1 > 0.0f) centre = s1;
  else if (s2 > 0.0f) centre = s2;
  else if (s3 > 0.0f) centre = s3;
  float sum = 0.0f;
  int n = 0;
  if (centre > 0.0f) {
    if (s0 > 0.0f && fabsf(s0 - centre) <= thr) { sum = sum + s0; ++n; }
    if (s1 > 0.0f && fabsf(s1 - centre) <= thr) { sum = sum + s1; ++n; }
    if (s2 > 0.0f && fabsf(s2 - centre) <= thr) { sum = sum + s2; ++n; }
    if (s3 > 0.0f && fabsf(s3 - centre) <= thr) { sum = sum + s3; ++n; }
  }
  *cnt = n;
  return n ? sum / (float)n : 0.0f;
}

/* counts the dividends a = +-2^e * 1.m, e in [-64, 64), whose div_cfg differs from the IEEE quotient */
__global__ void __launch_bounds__(256) k_div_check(float b, float r, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  const unsigned long long n = 128ull << 24; /* 128 binades x 2^23 mantissas x 2 signs */
  for (unsigned long long k = blockIdx.x * 256ull + threadIdx.x; k < n; k += 256ull * gridDim.x) {
    const uint32_t sign = (uint32_t)(k & 1), m = (uint32_t)(k >> 1) & 0x7FFFFFu, e = (uint32_t)(k >> 24) + (127u - 64u);
    const float a = __uint_as_float((sign << 31) | (e << 23) | m);
    if (__float_as_uint(div_cfg(a, b, r)) != __float_as_uint(a / b)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

/* back-projection (reference viewerModule.c:343-345) into plane 0 (vx,vy) and the .x half of
 * plane 1 (vz,nx) of a slot; an invalid pixel is stored as (0,0,0): a valid vertex has z > 0 */
__device__ __forceinline__ void store_vertex(float2* slot_base, size_t npix, size_t o, float d, int u, int v,
                                             const LevelGeom& g, float depth_factor, float r_df = 0.f, float r_fx = 0.f,
                                             float r_fy = 0.f, int fast_div = 0) {
  float x = 0.0f, y = 0.0f, z = 0.0f;
  if (fast_div) {
    if (d > 0.0f) {
      z = div_cfg(d, depth_factor, r_df);
      x = div_cfg(((float)u - g.cx) * z, g.fx, r_fx);
      y = div_cfg(((float)v - g.cy) * z, g.fy, r_fy);
    }
  } else if (d > 0.0f) {
    z = d / depth_factor;
    x = ((float)u - g.cx) * z / g.fx;
    y = ((float)v - g.cy) * z / g.fy;
  }
  slot_base[o] = make_float2(x, y);
  reinterpret_cast<float*>(slot_base + npix + o)[0] = z;
}

/* 7x7 bilateral for the two horizontally adjacent pixels A = (xo, y) and B = (xo+1, y) of the
 * tile (xo even).  `tile` holds raw depth as float with a far sentinel for invalid pixels; row 0
 * of the tile is image row y0-3 and column 0 is image column x0-8.
 * Per window row the eight columns xo-3 .. xo+4 are visited once each: column k is tap k of A and
 * tap k-1 of B, so one broadcast value feeds both halves of every packed operation and no
 * register pairs have to be assembled.  `wsp[dy][k]` = (ws[dy][k], ws[dy][k-1]) with 0 where a
 * pixel has no such tap: a zero weight adds +0 to sw and leaves swd unchanged, so each pixel sees
 * exactly its 49 taps in row-major order (the order of the specification).  swd uses a fused
 * multiply-add (specified: the CPU checker calls fmaf at the same place). */
__device__ __forceinline__ float2 bilateral_pair(const float (*tile)[YK_SMEM_W], const float* s_wr, const float2* wsp,
                                                 float cutf, int xo, int y) {
  const float2 c = make_float2(tile[y + YK_HALO][xo + 8], tile[y + YK_HALO][xo + 9]);
  float2 sw = make_float2(0.0f, 0.0f), swd = make_float2(0.0f, 0.0f);
  const float2 negc = make_float2(-c.x, -c.y);
  const char* lut0 = reinterpret_cast<const char*>(s_wr) - 0x4A000000; /* bits(2^21) = 0x4A000000 */
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    /* window columns xo-4 .. xo+5 as five aligned 64-bit shared loads; columns 1..8 are used */
    float w[10];
    const float2* row = reinterpret_cast<const float2*>(&tile[y + dy][xo + 4]);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float2 t = row[k];
      w[2 * k] = t.x;
      w[2 * k + 1] = t.y;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 fk = make_float2(w[1 + k], w[1 + k]);
      const float2 df = add2(fk, negc); /* exact: integer-valued floats */
      const float2 dcl = make_float2(fminf(fabsf(df.x), cutf), fminf(fabsf(df.y), cutf));
      /* small non-negative integer-valued float v -> byte offset 4*v: adding 2^21 leaves 4*v in the low
       * mantissa bits (ulp 0.25), so one integer add forms the shared-memory address of the LUT entry */
      const float2 mg = add2(dcl, make_float2(2097152.0f, 2097152.0f));
      const float2 wr = make_float2(*reinterpret_cast<const float*>(lut0 + __float_as_int(mg.x)),
                                    *reinterpret_cast<const float*>(lut0 + __float_as_int(mg.y)));
      const float2 wt = mul2(wsp[dy * 8 + k], wr);
      sw = add2(sw, wt);
      swd = fma2(wt, fk, swd);
    }
  }
  return make_float2(c.x != YK_SENTINEL ? swd.x / sw.x : 0.0f, c.y != YK_SENTINEL ? swd.y / sw.y : 0.0f);
}

/* Product-table variant (range_cut + 2 <= YK_WT_STRIDE, i.e. sigma_range up to 42 mm): the tap weight
 * ws[dy][dx] * wr[|diff|] is read ready-made from a shared table s_wt[class][|diff|] that every CTA fills
 * with the same single-precision products the generic path forms per tap (class = (|dy|, |dx|): the
 * spatial weight depends on dx^2 + dy^2 only).
 * The table address is formed in the floating-point pipe: for the small integer-valued float v,
 * fma(v, bits(4), bits(base)) is the subnormal whose bit pattern is base + 4*v -- exact, no conversion,
 * no integer add -- and the row of the tap is an immediate offset of the shared load.  Per column and
 * pixel pair: FADD2 (difference), 2 FMNMX (|.| and clamp), FFMA2 (address), 2 LDS, FADD2, FFMA2. */
__host__ __device__ constexpr int yk_wt_class(int dy, int k) { /* window row dy, window column k (0..6) */
  return (dy < 3 ? 3 - dy : dy - 3) * 4 + (k < 3 ? 3 - k : k - 3);
}

template <int DY, int K>
__device__ __forceinline__ void bilateral_col(float fkv, float2 negc, float cutf, float2 ulp4, float2 lbase, float2& sw,
                                              float2& swd) {
  const float2 fk = make_float2(fkv, fkv);
  const float2 df = add2(fk, negc); /* exact: integer-valued floats */
  /* column 0 is a tap of pixel A only, column 7 of pixel B only: the other half gets weight +0 without a
   * lookup (adds +0 to sw and leaves swd unchanged, exactly like a looked-up zero) */
  const float2 dcl = make_float2(fminf(fabsf(df.x), cutf), fminf(fabsf(df.y), cutf));
  const float2 ad = fma2(dcl, ulp4, lbase);
  float2 wt = make_float2(0.0f, 0.0f);
  if (K <= 6) asm("ld.shared.f32 %0, [%1+%2];" : "=f"(wt.x) : "r"(__float_as_int(ad.x)), "n"(yk_wt_class(DY, K <= 6 ? K : 0) * YK_WT_STRIDE * 4));
  if (K >= 1) asm("ld.shared.f32 %0, [%1+%2];" : "=f"(wt.y) : "r"(__float_as_int(ad.y)), "n"(yk_wt_class(DY, K >= 1 ? K - 1 : 0) * YK_WT_STRIDE * 4));
  sw = add2(sw, wt);
  swd = fma2(wt, fk, swd);
}

template <int DY>
__device__ __forceinline__ void bilateral_row(const float (*tile)[YK_SMEM_W], int xo, int y, float2 negc, float cutf,
                                              float2 ulp4, float2 lbase, float2& sw, float2& swd) {
  /* window columns xo-4 .. xo+5 as five aligned 64-bit shared loads; columns 1..8 are used */
  const float2* row = reinterpret_cast<const float2*>(&tile[y + DY][xo + 4]);
  const float2 t0 = row[0], t1 = row[1], t2 = row[2], t3 = row[3], t4 = row[4];
  bilateral_col<DY, 0>(t0.y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 1>(t1.x, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 2>(t1.y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 3>(t2.x, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 4>(t2.y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 5>(t3.x, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 6>(t3.y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_col<DY, 7>(t4.x, negc, cutf, ulp4, lbase, sw, swd);
}

__device__ __forceinline__ float2 bilateral_pair_wt(const float (*tile)[YK_SMEM_W], const float* s_wt, float cutf, int xo,
                                                    int y) {
  const float2 c = make_float2(tile[y + YK_HALO][xo + 8], tile[y + YK_HALO][xo + 9]);
  float2 sw = make_float2(0.0f, 0.0f), swd = make_float2(0.0f, 0.0f);
  const float2 negc = make_float2(-c.x, -c.y);
  const float ulp = __int_as_float(4), lb = __int_as_float((int)__cvta_generic_to_shared(s_wt));
  const float2 ulp4 = make_float2(ulp, ulp), lbase = make_float2(lb, lb);
  bilateral_row<0>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_row<1>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_row<2>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_row<3>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_row<4>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_row<5>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  bilateral_row<6>(tile, xo, y, negc, cutf, ulp4, lbase, sw, swd);
  return make_float2(c.x != YK_SENTINEL ? swd.x / sw.x : 0.0f, c.y != YK_SENTINEL ? swd.y / sw.y : 0.0f);
}

/* The same for TWO vertically adjacent pixel pairs, rows y and y + 1 of the tile: their 7-row windows share six of
 * eight rows, so every window row is loaded once (five LDS.64) and feeds both -- row R of the eight is window row
 * dy = R of the upper pair and dy = R - 1 of the lower pair.  Each pixel still sees its 49 taps in row-major order
 * (the order of the specification); only the interleaving between the four pixels changes.  40 instead of 70 row
 * loads per four pixels: shared-memory wavefronts are this kernel's tightest resource (ncu: ~90 % of the LSU
 * data-pipe peak), the table lookups are one per tap either way. */
template <int R>
__device__ __forceinline__ void bilateral_row2(const float (*tile)[YK_SMEM_W], int xo, int y, float2 negcA, float2 negcB,
                                               float cutf, float2 ulp4, float2 lbase, float2& swA, float2& swdA, float2& swB,
                                               float2& swdB) {
  const float2* row = reinterpret_cast<const float2*>(&tile[y + R][xo + 4]);
  const float2 t0 = row[0], t1 = row[1], t2 = row[2], t3 = row[3], t4 = row[4];
  if (R <= 6) {
    constexpr int DY = R <= 6 ? R : 0;
    bilateral_col<DY, 0>(t0.y, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 1>(t1.x, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 2>(t1.y, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 3>(t2.x, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 4>(t2.y, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 5>(t3.x, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 6>(t3.y, negcA, cutf, ulp4, lbase, swA, swdA);
    bilateral_col<DY, 7>(t4.x, negcA, cutf, ulp4, lbase, swA, swdA);
  }
  if (R >= 1) {
    constexpr int DY = R >= 1 ? R - 1 : 0;
    bilateral_col<DY, 0>(t0.y, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 1>(t1.x, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 2>(t1.y, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 3>(t2.x, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 4>(t2.y, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 5>(t3.x, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 6>(t3.y, negcB, cutf, ulp4, lbase, swB, swdB);
    bilateral_col<DY, 7>(t4.x, negcB, cutf, ulp4, lbase, swB, swdB);
  }
}

/* rows y and y + 1: .x/.y of dA are the pixels (xo, y), (xo+1, y); dB the same one row down */
__device__ __forceinline__ void bilateral_quad_wt(const float (*tile)[YK_SMEM_W], const float* s_wt, float cutf, int xo, int y,
                                                  float2& dA, float2& dB) {
  const float2 cA = make_float2(tile[y + YK_HALO][xo + 8], tile[y + YK_HALO][xo + 9]);
  const float2 cB = make_float2(tile[y + 1 + YK_HALO][xo + 8], tile[y + 1 + YK_HALO][xo + 9]);
  float2 swA = make_float2(0.0f, 0.0f), swdA = swA, swB = swA, swdB = swA;
  const float2 negcA = make_float2(-cA.x, -cA.y), negcB = make_float2(-cB.x, -cB.y);
  const float ulp = __int_as_float(4), lb = __int_as_float((int)__cvta_generic_to_shared(s_wt));
  const float2 ulp4 = make_float2(ulp, ulp), lbase = make_float2(lb, lb);
  bilateral_row2<0>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<1>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<2>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<3>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<4>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<5>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<6>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  bilateral_row2<7>(tile, xo, y, negcA, negcB, cutf, ulp4, lbase, swA, swdA, swB, swdB);
  dA = make_float2(cA.x != YK_SENTINEL ? swdA.x / swA.x : 0.0f, cA.y != YK_SENTINEL ? swdA.y / swA.y : 0.0f);
  dB = make_float2(cB.x != YK_SENTINEL ? swdB.x / swB.x : 0.0f, cB.y != YK_SENTINEL ? swdB.y / swB.y : 0.0f);
}

#define YK_D0_W (YK_TILE_W + 4) /* level-0 depth tile with the +1 halo column/row the normals need (66 used) */
#define YK_D0_H (YK_TILE_H + 1)

/* FD: the vertex divisions in the reciprocal form (div_cfg; the host picks the instantiation after k_div_check).
 * DBG: also store the filtered float depth pyramid and the pyramid sample counts -- nothing on the product path reads
 * them (0.5 GB per 300 frames and 485 MB per handle saved); parity read-back and the model ray cast's depth hint do. */
template <int MODE, bool FD, bool DBG>
__global__ void __launch_bounds__(256, YK_INGEST_MIN_BLOCKS) k_ingest(const __grid_constant__ IngestParams P) {
  constexpr bool BILATERAL = MODE != YK_INGEST_RAW;
  __shared__ __align__(16) float tile[YK_SMEM_H][YK_SMEM_W];
  __shared__ __align__(8) float d0s[YK_D0_H][YK_D0_W];
  /* 65 columns used; the pitch of 66 keeps every row 16-byte (vxy) / 8-byte (vz) aligned, so the normals pass
   * reads a pixel pair with one 128-bit / 64-bit load instead of stride-2 64-bit / 32-bit loads (which cost two
   * wavefronts per ideal one: 15 % of this kernel's shared-memory wavefronts, its tightest resource) */
  __shared__ __align__(16) float2 vxy[YK_D0_H][YK_TILE_W + 2];
  __shared__ __align__(8) float vz[YK_D0_H][YK_TILE_W + 2];
  __shared__ float d1s[YK_TILE_H / 2][YK_TILE_W / 2];
  __shared__ float d2s[YK_TILE_H / 4][YK_TILE_W / 4];
  __shared__ __align__(16) float s_wr[MODE == YK_INGEST_BILATERAL ? YK_RANGE_LUT_MAX : (MODE == YK_INGEST_BILATERAL_WT ? YK_WT_ROWS * YK_WT_STRIDE : 1)];

  const int tid = threadIdx.x;
  /* one sequence (the usual case): no division at the head of every CTA's dependency chain */
  const int s = P.ring.S == 1 ? 0 : blockIdx.z / P.chunk_n, i = P.frame0 + (blockIdx.z - s * P.chunk_n);
  const int frame = s * P.ring.n + i; /* pair index inside the group */
  const int slot = ring_slot(P.ring, i);
  const int W = P.lv[0].w, H = P.lv[0].h;
  const int x0 = blockIdx.x * YK_TILE_W, y0 = blockIdx.y * YK_TILE_H;
  const uint16_t* raw = P.raw[s] + (size_t)i * W * H;
  const size_t slot_idx = (size_t)s * P.ring.R + slot;

  if (blockIdx.x == 0 && blockIdx.y == 0 && tid < 12) { /* every pair starts from the identity */
    const double v = (tid == 0 || tid == 5 || tid == 10) ? 1.0 : 0.0;
    P.pose_d[frame * 12 + tid] = v;
    P.pose_f[frame * 12 + tid] = (float)v;
    if (tid == 0) P.pair_status[frame] = 0u;
  }
  if (MODE == YK_INGEST_BILATERAL) {
    for (int k = tid; k < P.range_cut + 2; k += 256) s_wr[k] = P.wr[k];
  }
  if (MODE == YK_INGEST_BILATERAL_WT) { /* s_wt[class][|diff|] = ws * wr: two 128-bit loads per thread from the 8 KB table */
    for (int k = tid; k < YK_WT_ROWS * YK_WT_STRIDE / 4; k += 256) reinterpret_cast<float4*>(s_wr)[k] = __ldg(P.wt + k);
  }
  /* stage raw depth through shared memory: 128-bit loads of 8 pixels, converted to float
   * with invalid / out-of-image pixels replaced by a far sentinel.  Tile rows y0-3 .. y0+19,
   * columns x0-8 .. x0+71 (the extra row/column feed the +1 halo of the normals). */
  constexpr int VEC_PER_ROW = YK_SMEM_W / 8;
  for (int k = tid; k < YK_SMEM_H * VEC_PER_ROW; k += 256) {
    const int r = k / VEC_PER_ROW, c8 = k - r * VEC_PER_ROW;
    const int gy = y0 - YK_HALO + r, gx = x0 - 8 + c8 * 8;
    float f[8];
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const uint4 v = ldg_nc_u4(reinterpret_cast<const uint4*>(raw + (size_t)gy * W + gx));
      const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int a = (int)(wds[j] & 0xFFFFu), b = (int)(wds[j] >> 16);
        f[2 * j] = (a >= P.dmin && a <= P.dmax) ? (float)a : YK_SENTINEL;
        f[2 * j + 1] = (b >= P.dmin && b <= P.dmax) ? (float)b : YK_SENTINEL;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = YK_SENTINEL;
    }
    float4* dst = reinterpret_cast<float4*>(&tile[r][c8 * 8]);
    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();

  /* level-0 depth: work items are pixel pairs (xo, xo+1); 32 x 16 items cover the tile, 49 more
   * cover the halo row y = 16 and the halo column pair (64, 65) */
  const float cutf = (float)(P.range_cut + 1);
  if (MODE == YK_INGEST_BILATERAL_WT) {
    /* the tile proper in ONE pass: every thread filters a 2 x 2 block (two pixel pairs in rows 2w, 2w + 1 of warp w);
     * the window rows the two rows share are loaded once */
    const int xo = 2 * (tid & 31), y = 2 * (tid >> 5);
    float2 dA, dB;
    bilateral_quad_wt(tile, s_wr, cutf, xo, y, dA, dB);
    *reinterpret_cast<float2*>(&d0s[y][xo]) = dA;
    *reinterpret_cast<float2*>(&d0s[y + 1][xo]) = dB;
  }
#pragma unroll 1
  for (int pass = (MODE == YK_INGEST_BILATERAL_WT ? 2 : 0); pass < 3; ++pass) {
    int xo, y;
    if (pass < 2) {
      xo = 2 * (tid & 31);
      y = (tid >> 5) + 8 * pass;
    } else if (tid < 32) {
      xo = 2 * tid;
      y = YK_TILE_H;
    } else if (tid < 32 + YK_D0_H) {
      xo = YK_TILE_W;
      y = tid - 32;
    } else {
      break;
    }
    float2 d;
    if (BILATERAL) {
      d = MODE == YK_INGEST_BILATERAL_WT ? bilateral_pair_wt(tile, s_wr, cutf, xo, y)
                                         : bilateral_pair(tile, s_wr, P.wsp, cutf, xo, y);
    } else {
      const float a = tile[y + YK_HALO][xo + 8], b = tile[y + YK_HALO][xo + 9];
      d = make_float2(a != YK_SENTINEL ? a : 0.0f, b != YK_SENTINEL ? b : 0.0f);
    }
    d0s[y][xo] = d.x;
    d0s[y][xo + 1] = d.y;
  }
  __syncthreads();
  /* level-0 vertices of the tile + halo into shared memory (the normals need right/lower neighbours) */
  for (int k = tid; k < YK_D0_H * (YK_TILE_W + 1); k += 256) {
    const int y = k / (YK_TILE_W + 1), x = k - y * (YK_TILE_W + 1);
    const float d = d0s[y][x];
    float vx = 0.0f, vy = 0.0f, vzz = 0.0f;
    if (FD) {
      if (d > 0.0f) {
        vzz = div_cfg(d, P.depth_factor, P.r_df);
        vx = div_cfg(((float)(x0 + x) - P.lv[0].cx) * vzz, P.lv[0].fx, P.r_fx[0]);
        vy = div_cfg(((float)(y0 + y) - P.lv[0].cy) * vzz, P.lv[0].fy, P.r_fy[0]);
      }
    } else if (d > 0.0f) { /* reference viewerModule.c:343-345 */
      vzz = d / P.depth_factor;
      vx = ((float)(x0 + x) - P.lv[0].cx) * vzz / P.lv[0].fx;
      vy = ((float)(y0 + y) - P.lv[0].cy) * vzz / P.lv[0].fy;
    }
    vxy[y][x] = make_float2(vx, vy);
    vz[y][x] = vzz;
  }
  __syncthreads();
  /* level-0 normals + all three map planes, two pixels (128 bits per plane) per thread and pass */
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int xo = 2 * (tid & 31), y = (tid >> 5) + 8 * pass;
    const int gx = x0 + xo, gy = y0 + y;
    float nx[2], ny[2], nz[2];
    /* vertices of the pair, its right neighbour and the pair below: columns xo, xo+1 (128 bits), xo+2; row y+1 */
    const float4 v01 = *reinterpret_cast<const float4*>(&vxy[y][xo]);
    const float2 v2 = vxy[y][xo + 2];
    const float4 vd = *reinterpret_cast<const float4*>(&vxy[y + 1][xo]);
    const float2 z01 = *reinterpret_cast<const float2*>(&vz[y][xo]);
    const float z2 = vz[y][xo + 2];
    const float2 zd = *reinterpret_cast<const float2*>(&vz[y + 1][xo]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      nx[e] = YK_N_INVALID;
      ny[e] = nz[e] = 0.0f;
      const float z0 = e ? z01.y : z01.x, zx = e ? z2 : z01.y, zy = e ? zd.y : zd.x;
      if (z0 > 0.0f && zx > 0.0f && zy > 0.0f) {
        const float2 a0 = e ? make_float2(v01.z, v01.w) : make_float2(v01.x, v01.y);
        const float2 ax = e ? v2 : make_float2(v01.z, v01.w);
        const float2 ay = e ? make_float2(vd.z, vd.w) : make_float2(vd.x, vd.y);
        const float ex = ax.x - a0.x, ey = ax.y - a0.y, ez = zx - z0;
        const float fx = ay.x - a0.x, fy = ay.y - a0.y, fz = zy - z0;
        const float cx = ey * fz - ez * fy;
        const float cy = ez * fx - ex * fz;
        const float cz = ex * fy - ey * fx;
        const float len2 = (cx * cx + cy * cy) + cz * cz;
        if (len2 > 1e-24f) {
          const float inv = FD ? inv_len(len2) : 1.0f / sqrtf(len2);
          nx[e] = cx * inv;
          ny[e] = cy * inv;
          nz[e] = cz * inv;
        }
      }
    }
    if (gx < W && gy < H) { /* W is even, so the pair is inside or outside together */
      const size_t np0 = (size_t)W * H, o = (size_t)gy * W + gx;
      if (DBG) *reinterpret_cast<float2*>(P.depth[0] + slot_idx * np0 + o) = make_float2(d0s[y][xo], d0s[y][xo + 1]);
      float2* base = P.maps[0] + slot_idx * 3 * np0;
      *reinterpret_cast<float4*>(base + o) = v01;
      *reinterpret_cast<float4*>(base + np0 + o) = make_float4(z01.x, nx[0], z01.y, nx[1]);
      *reinterpret_cast<float4*>(base + 2 * np0 + o) = make_float4(ny[0], nz[0], ny[1], nz[1]);
    }
  }
  if (P.levels < 2) return;
  /* level 1: 32x8 pixels per tile, one per thread */
  {
    const int lx = tid & 31, ly = tid >> 5;
    int n;
    const float d = pyr_combine(d0s[2 * ly][2 * lx], d0s[2 * ly][2 * lx + 1], d0s[2 * ly + 1][2 * lx],
                                d0s[2 * ly + 1][2 * lx + 1], P.pyr_thr, &n);
    d1s[ly][lx] = d;
    const int w1 = P.lv[1].w, h1 = P.lv[1].h;
    const int gx = (x0 >> 1) + lx, gy = (y0 >> 1) + ly;
    if (gx < w1 && gy < h1) {
      const size_t npl = (size_t)w1 * h1, o = (size_t)gy * w1 + gx;
      if (DBG) {
        P.depth[1][slot_idx * npl + o] = d;
        P.pyrcnt[1][slot_idx * npl + o] = (uint8_t)n;
      }
      store_vertex(P.maps[1] + slot_idx * 3 * npl, npl, o, d, gx, gy, P.lv[1], P.depth_factor, P.r_df, P.r_fx[1], P.r_fy[1], FD);
    }
  }
  if (P.levels < 3) return;
  __syncthreads();
  if (tid < 64) {
    const int lx = tid & 15, ly = tid >> 4;
    int n;
    const float d = pyr_combine(d1s[2 * ly][2 * lx], d1s[2 * ly][2 * lx + 1], d1s[2 * ly + 1][2 * lx],
                                d1s[2 * ly + 1][2 * lx + 1], P.pyr_thr, &n);
    d2s[ly][lx] = d;
    const int w2 = P.lv[2].w, h2 = P.lv[2].h;
    const int gx = (x0 >> 2) + lx, gy = (y0 >> 2) + ly;
    if (gx < w2 && gy < h2) {
      const size_t npl = (size_t)w2 * h2, o = (size_t)gy * w2 + gx;
      if (DBG) {
        P.depth[2][slot_idx * npl + o] = d;
        P.pyrcnt[2][slot_idx * npl + o] = (uint8_t)n;
      }
      store_vertex(P.maps[2] + slot_idx * 3 * npl, npl, o, d, gx, gy, P.lv[2], P.depth_factor, P.r_df, P.r_fx[2], P.r_fy[2], FD);
    }
  }
  if (P.levels < 4) return;
  __syncthreads();
  if (tid < 16) {
    const int lx = tid & 7, ly = tid >> 3;
    int n;
    const float d = pyr_combine(d2s[2 * ly][2 * lx], d2s[2 * ly][2 * lx + 1], d2s[2 * ly + 1][2 * lx],
                                d2s[2 * ly + 1][2 * lx + 1], P.pyr_thr, &n);
    const int w3 = P.lv[3].w, h3 = P.lv[3].h;
    const int gx = (x0 >> 3) + lx, gy = (y0 >> 3) + ly;
    if (gx < w3 && gy < h3) {
      const size_t npl = (size_t)w3 * h3, o = (size_t)gy * w3 + gx;
      if (DBG) {
        P.depth[3][slot_idx * npl + o] = d;
        P.pyrcnt[3][slot_idx * npl + o] = (uint8_t)n;
      }
      store_vertex(P.maps[3] + slot_idx * 3 * npl, npl, o, d, gx, gy, P.lv[3], P.depth_factor, P.r_df, P.r_fx[3], P.r_fy[3], FD);
    }
  }
}

/* ------------------------------------------------------------------ k_normals */

template <bool FD>
__global__ void __launch_bounds__(256) k_normals(const __grid_constant__ NormalParams P) {
  int p = blockIdx.x * 256 + threadIdx.x;
  const int s = P.ring.S == 1 ? 0 : blockIdx.y / P.chunk_n, i = P.frame0 + (blockIdx.y - s * P.chunk_n);
  const size_t slot_idx = (size_t)s * P.ring.R + ring_slot(P.ring, i);
  int level = P.first_level;
  for (; level < P.levels; ++level) {
    const int np = P.lv[level].w * P.lv[level].h;
    if (p < np) break;
    p -= np;
  }
  if (level >= P.levels) return;
  const int W = P.lv[level].w, H = P.lv[level].h;
  const size_t npix = (size_t)W * H;
  float2* base = P.maps[level] + slot_idx * 3 * npix;
  const float2* XY = base;
  const float2* ZN = base + npix;
  const int v = p / W, u = p - v * W;
  const float z0 = ZN[p].x;
  float ox = YK_N_INVALID, oy = 0.0f, oz = 0.0f; /* an invalid normal is stored as (2,0,0) */
  if (u + 1 < W && v + 1 < H) {
    const float zx = ZN[p + 1].x, zy = ZN[p + W].x;
    if (z0 > 0.0f && zx > 0.0f && zy > 0.0f) {
      const float2 a0 = XY[p], ax = XY[p + 1], ay = XY[p + W];
      const float ex = ax.x - a0.x, ey = ax.y - a0.y, ez = zx - z0;
      const float fx = ay.x - a0.x, fy = ay.y - a0.y, fz = zy - z0;
      const float nx = ey * fz - ez * fy;
      const float ny = ez * fx - ex * fz;
      const float nz = ex * fy - ey * fx;
      const float len2 = (nx * nx + ny * ny) + nz * nz;
      if (len2 > 1e-24f) {
        const float inv = FD ? inv_len(len2) : 1.0f / sqrtf(len2);
        ox = nx * inv;
        oy = ny * inv;
        oz = nz * inv;
      }
    }
  }
  base[npix + p] = make_float2(z0, ox); /* completes the (vz,nx) plane with a full 8-byte store */
  base[2 * npix + p] = make_float2(oy, oz);
}

/* ------------------------------------------------------------------ launchers */

template <int MODE>
static void launch_ingest_mode(bool fd, bool dbg, dim3 grid, cudaStream_t st, const IngestParams& p) {
  if (fd) {
    if (dbg) k_ingest<MODE, true, true><<<grid, 256, 0, st>>>(p);
    else k_ingest<MODE, true, false><<<grid, 256, 0, st>>>(p);
  } else {
    if (dbg) k_ingest<MODE, false, true><<<grid, 256, 0, st>>>(p);
    else k_ingest<MODE, false, false><<<grid, 256, 0, st>>>(p);
  }
}

void yk_launch_ingest(int mode, bool fast_div, bool debug_maps, dim3 grid, cudaStream_t st, const IngestParams& p) {
  if (mode == YK_INGEST_RAW) launch_ingest_mode<YK_INGEST_RAW>(fast_div, debug_maps, grid, st, p);
  else if (mode == YK_INGEST_BILATERAL_WT) launch_ingest_mode<YK_INGEST_BILATERAL_WT>(fast_div, debug_maps, grid, st, p);
  else launch_ingest_mode<YK_INGEST_BILATERAL>(fast_div, debug_maps, grid, st, p);
}

void yk_launch_normals(bool fast_div, dim3 grid, cudaStream_t st, const NormalParams& p) {
  if (fast_div) k_normals<true><<<grid, 256, 0, st>>>(p);
  else k_normals<false><<<grid, 256, 0, st>>>(p);
}

void yk_launch_div_check(float b, float r, unsigned long long* d_mismatches) { k_div_check<<<148 * 8, 256>>>(b, r, d_mismatches); }
