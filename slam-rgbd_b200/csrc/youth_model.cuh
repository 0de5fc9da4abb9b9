/*
 * youth_model.cuh -- sm_100a kernels of frame-to-model tracking (include/youth_model.h):
 *
 *   k_tsdf_integrate  one thread per voxel column segment (32 x 8 threads in x, y; 16 voxels in z,
 *                     independent iterations so the depth gather and the voxel read of several
 *                     voxels are in flight together): project the voxel centre into the frame,
 *                     projective signed distance, truncated running average in (int16, int16)
 *                     voxels.  A voxel is READ ONLY IF it is updated, so the traffic is the
 *                     visible truncation band, not the volume.
 *   k_tsdf_raycast    one thread per pixel of every pyramid level: march the ray in steps of
 *                     0.8 mu (shrinking inside the truncation band) on nearest-voxel samples, starting
 *                     just in front of the depth the fused frame measured there; refine the zero crossing with two
 *                     trilinear samples; writes the model vertex map in the three-float2-plane layout
 *                     k_icp gathers from.  The model normals are the cross-product normals of that
 *                     vertex map: k_normals (stage 2b) runs over the model maps right after.
 *   k_fill_u32        volume reset
 *
 * Same arithmetic contract as youth_common.cuh (--fmad=false, explicit fma where specified,
 * IEEE division / sqrt); bit-identical to the CPU statement.  HBM/L2-bound gather work: no
 * tensor cores.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "youth_common.cuh"
#include "youth_model.h"

#define YM_SCALE 32767.0f
#define YM_INV_SCALE (1.0f / 32767.0f)
#define YM_UNKNOWN 2.0f
#define YM_ZCHUNK 16
#ifndef YM_RAYCAST_MIN_BLOCKS
#define YM_RAYCAST_MIN_BLOCKS 6 /* 40 registers: measured 8 % faster than 64 registers at 4 CTAs per SM */
#endif

struct TsdfGeom {
  int dx, dy, dz;
  float vs, inv_vs;
  float ox, oy, oz;
  float mu, step;
  int maxw;
  float near_m, far_m;
};

struct IntegrateParams {
  short2* vol;                 /* [S][dz][dy][dx] */
  TsdfGeom t;
  const float2* maps0;         /* [S][R][3][npix0] level-0 map planes of the ring: plane 1 .x is z = D / depth_factor,
                                  computed by stage 2 with the very expression the specification uses here */
  LevelGeom g;                 /* level 0 */
  RingGeom ring;
  const float* world_f;        /* [S][12] camera-to-world */
  const uint32_t* last_status; /* [S]: a frame flagged LOST is not fused; NULL = always fuse */
  float depth_factor;
  int slot;                    /* >= 0: this ring slot (debug); < 0: the newest frame, head - 1 */
  int stream0;                 /* first sequence handled by blockIdx.z / zchunks */
};

struct RaycastParams {
  const short2* vol;
  TsdfGeom t;
  float2* model[YOUTH_MAX_LEVELS]; /* [S][3][npix_l] planes (vx,vy) (vz,nx) (ny,nz) */
  LevelGeom lv[YOUTH_MAX_LEVELS];
  int levels;
  const float* world_f;
  int stream0;
  /* march-start hint: the depth pyramid of a resident frame taken from the same pose (the frame just fused) */
  const float* depth[YOUTH_MAX_LEVELS]; /* [S][R][npix_l]; used when hint != 0 */
  RingGeom ring;
  int hint;      /* 0: start at the near plane */
  int hint_slot; /* >= 0: this ring slot; < 0: the newest frame (head - 1) */
  float depth_factor;
};

__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = v;
}

/* size of the map: observed voxels whose tsdf changes sign towards an observed +x / +y / +z neighbour
 * (integer count, so the atomic accumulation is order independent) */
__global__ void __launch_bounds__(256) k_tsdf_surface_count(const short2* __restrict__ vol, int dx, int dy, int dz,
                                                            unsigned long long* out) {
  const size_t n = (size_t)dx * dy * dz;
  unsigned int local = 0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const short2 v = __ldg(vol + i);
    if (v.y <= 0) continue;
    const int ix = (int)(i % dx), iy = (int)((i / dx) % dy), iz = (int)(i / ((size_t)dx * dy));
    const bool neg = v.x < 0;
    bool hit = false;
    if (ix + 1 < dx) {
      const short2 w = __ldg(vol + i + 1);
      hit |= w.y > 0 && (w.x < 0) != neg;
    }
    if (iy + 1 < dy) {
      const short2 w = __ldg(vol + i + dx);
      hit |= w.y > 0 && (w.x < 0) != neg;
    }
    if (iz + 1 < dz) {
      const short2 w = __ldg(vol + i + (size_t)dx * dy);
      hit |= w.y > 0 && (w.x < 0) != neg;
    }
    local += hit ? 1u : 0u;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, (unsigned long long)local);
}

__global__ void __launch_bounds__(256) k_tsdf_integrate(const __grid_constant__ IntegrateParams P) {
  const int zchunks = (P.t.dz + YM_ZCHUNK - 1) / YM_ZCHUNK;
  const int s = P.stream0 + blockIdx.z / zchunks, zc = blockIdx.z % zchunks;
  if (P.last_status != nullptr && (P.last_status[s] & YOUTH_STATUS_LOST)) return;
  const int ix = blockIdx.x * 32 + (threadIdx.x & 31), iy = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int slot = P.slot >= 0 ? P.slot : (__ldg(P.ring.head) + P.ring.R - 1) % P.ring.R;
  const size_t npix0 = (size_t)(P.g.w * P.g.h);
  const float2* __restrict__ zplane = P.maps0 + (((size_t)s * P.ring.R + slot) * 3 + 1) * npix0;
  const float* T = P.world_f + s * 12;
  /* world -> camera: R^T and -R^T t */
  float Ri[9], ti[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) Ri[3 * i + j] = __ldg(T + 4 * j + i);
    ti[i] = -__fmaf_rn(__ldg(T + i), __ldg(T + 3), __fmaf_rn(__ldg(T + 4 + i), __ldg(T + 7), __ldg(T + 8 + i) * __ldg(T + 11)));
  }
  /* Conservative frustum culling of the CTA's box of voxels (32 x 8 x 16): its eight corners lie half a
   * voxel outside the outermost voxel centres; when all eight are on the outer side of one frustum plane
   * (the side planes pass through the camera centre: n . p < 0 with n = (fx, 0, cx+1/2) etc.) no voxel of
   * the box can pass the per-voxel image test below, so skipping the box changes nothing.  The half
   * voxel (>= 1 pixel at any depth inside the volume) dwarfs the rounding differences between this test and
   * the per-voxel arithmetic. */
  {
    __shared__ int s_keep;
    if (threadIdx.x == 0) s_keep = 0;
    __syncthreads();
    if (threadIdx.x < 8) {
      const int cx0 = blockIdx.x * 32, cy0 = blockIdx.y * 8, cz0 = zc * YM_ZCHUNK;
      const float bx = P.t.ox + (float)(cx0 + ((threadIdx.x & 1) ? 32 : 0)) * P.t.vs;
      const float by = P.t.oy + (float)(cy0 + ((threadIdx.x & 2) ? 8 : 0)) * P.t.vs;
      const float bz = P.t.oz + (float)(cz0 + ((threadIdx.x & 4) ? YM_ZCHUNK : 0)) * P.t.vs;
      const float qx = Ri[0] * bx + Ri[1] * by + Ri[2] * bz + ti[0];
      const float qy = Ri[3] * bx + Ri[4] * by + Ri[5] * bz + ti[1];
      const float qz = Ri[6] * bx + Ri[7] * by + Ri[8] * bz + ti[2];
      const float hu = P.g.fx * qx + P.g.cxh * qz, hv = P.g.fy * qy + P.g.cyh * qz; /* u * z, v * z */
      const unsigned m = 0xffu;
      const bool out_near = __all_sync(m, qz <= 0.0f);
      const bool out_left = __all_sync(m, hu < 0.0f), out_right = __all_sync(m, hu >= (float)P.g.w * qz && qz > 0.0f);
      const bool out_top = __all_sync(m, hv < 0.0f), out_bottom = __all_sync(m, hv >= (float)P.g.h * qz && qz > 0.0f);
      if (threadIdx.x == 0 && !(out_near || out_left || out_right || out_top || out_bottom)) s_keep = 1;
    }
    __syncthreads();
    if (!s_keep) return;
  }
  if (ix >= P.t.dx || iy >= P.t.dy) return;
  const float wx = __fmaf_rn((float)ix + 0.5f, P.t.vs, P.t.ox);
  const float wy = __fmaf_rn((float)iy + 0.5f, P.t.vs, P.t.oy);
  /* the x / y part of the transform is shared by the column; the nesting order of the specification is
   * fma(R0, wx, fma(R1, wy, fma(R2, wz, t))), so only the innermost term depends on z */
  short2* col = P.vol + (size_t)s * P.t.dx * P.t.dy * P.t.dz + (size_t)iy * P.t.dx + ix;
  const size_t zstride = (size_t)P.t.dx * P.t.dy;
  const int z0 = zc * YM_ZCHUNK;
#pragma unroll 4
  for (int k = 0; k < YM_ZCHUNK; ++k) {
    const int iz = z0 + k;
    if (iz >= P.t.dz) break;
    const float wz = __fmaf_rn((float)iz + 0.5f, P.t.vs, P.t.oz);
    const float px = __fmaf_rn(Ri[0], wx, __fmaf_rn(Ri[1], wy, __fmaf_rn(Ri[2], wz, ti[0])));
    const float py = __fmaf_rn(Ri[3], wx, __fmaf_rn(Ri[4], wy, __fmaf_rn(Ri[5], wz, ti[1])));
    const float pz = __fmaf_rn(Ri[6], wx, __fmaf_rn(Ri[7], wy, __fmaf_rn(Ri[8], wz, ti[2])));
    if (!(pz > 0.0f)) continue;
    const float iz1 = 1.0f / pz;
    const float ur = __fmaf_rn(px * P.g.fx, iz1, P.g.cxh);
    const float vr = __fmaf_rn(py * P.g.fy, iz1, P.g.cyh);
    if (!(ur >= 0.0f && ur < (float)P.g.w && vr >= 0.0f && vr < (float)P.g.h)) continue;
    /* z of the measured vertex = D / depth_factor (0 where the pixel has no reading) */
    const float zm = __ldg(reinterpret_cast<const float*>(zplane + __float2int_rz(vr) * P.g.w + __float2int_rz(ur)));
    if (!(zm > 0.0f)) continue;
    const float sdf = zm - pz;
    if (!(sdf >= -P.t.mu)) continue;
    float f = sdf / P.t.mu;
    if (f > 1.0f) f = 1.0f;
    short2* vp = col + (size_t)iz * zstride;
    const short2 v = *vp;
    if (f == 1.0f && v.x == 32767) {
      /* free space seen as free space again: (F W + 1) / (W + 1) with F = 32767 * fl(1/32767) differs from 1 by
       * less than 1e-6, so the stored value stays 32767 -- only the weight moves (and not at all once it
       * has reached the cap): the same result as the arithmetic below, without it */
      if (v.y < P.t.maxw) *vp = make_short2((short)32767, (short)(v.y + 1));
      continue;
    }
    const float F = (float)v.x * YM_INV_SCALE, W = (float)v.y;
    const float Fn = (F * W + f) / (W + 1.0f);
    const int wn = v.y + 1 > P.t.maxw ? P.t.maxw : v.y + 1;
    *vp = make_short2((short)__float2int_rn(Fn * YM_SCALE), (short)wn);
  }
}

__device__ __forceinline__ float tsdf_nearest(const short2* __restrict__ vol, const TsdfGeom& t, float gx, float gy, float gz) {
  int ix = __float2int_rz(gx + 0.5f), iy = __float2int_rz(gy + 0.5f), iz = __float2int_rz(gz + 0.5f);
  ix = max(0, min(t.dx - 1, ix));
  iy = max(0, min(t.dy - 1, iy));
  iz = max(0, min(t.dz - 1, iz));
  const short2 v = __ldg(vol + ((size_t)iz * t.dy + iy) * t.dx + ix);
  return v.y > 0 ? (float)v.x * YM_INV_SCALE : YM_UNKNOWN;
}

__device__ __forceinline__ bool tsdf_trilinear(const short2* __restrict__ vol, const TsdfGeom& t, float gx, float gy, float gz,
                                               float* out) {
  if (!(gx >= 0.0f && gy >= 0.0f && gz >= 0.0f)) return false;
  if (!(gx < (float)(t.dx - 1) && gy < (float)(t.dy - 1) && gz < (float)(t.dz - 1))) return false;
  const int ix = __float2int_rz(gx), iy = __float2int_rz(gy), iz = __float2int_rz(gz);
  const float fx = gx - (float)ix, fy = gy - (float)iy, fz = gz - (float)iz;
  const short2* b = vol + ((size_t)iz * t.dy + iy) * t.dx + ix;
  const size_t sy = (size_t)t.dx, sz = (size_t)t.dx * t.dy;
  short2 v[8];
  v[0] = __ldg(b);
  v[1] = __ldg(b + 1);
  v[2] = __ldg(b + sy);
  v[3] = __ldg(b + sy + 1);
  v[4] = __ldg(b + sz);
  v[5] = __ldg(b + sz + 1);
  v[6] = __ldg(b + sz + sy);
  v[7] = __ldg(b + sz + sy + 1);
  bool ok = true;
  float c[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ok = ok && (v[k].y > 0);
    c[k] = (float)v[k].x * YM_INV_SCALE;
  }
  if (!ok) return false;
  const float c00 = c[0] + fx * (c[1] - c[0]);
  const float c10 = c[2] + fx * (c[3] - c[2]);
  const float c01 = c[4] + fx * (c[5] - c[4]);
  const float c11 = c[6] + fx * (c[7] - c[6]);
  const float c0 = c00 + fy * (c10 - c00);
  const float c1 = c01 + fy * (c11 - c01);
  *out = c0 + fz * (c1 - c0);
  return true;
}

__global__ void __launch_bounds__(256, YM_RAYCAST_MIN_BLOCKS) k_tsdf_raycast(const __grid_constant__ RaycastParams P) {
  int p = blockIdx.x * 256 + threadIdx.x;
  const int s = P.stream0 + blockIdx.y;
  /* levels are laid out one after the other, each padded to a multiple of 32 pixels, so a warp never
   * straddles two levels and its 32 lanes are the specification's group of 32 consecutive pixels */
  int level = 0;
  for (; level < P.levels; ++level) {
    const int np_pad = (P.lv[level].w * P.lv[level].h + 31) & ~31;
    if (p < np_pad) break;
    p -= np_pad;
  }
  if (level >= P.levels) return; /* warp-uniform */
  const LevelGeom g = P.lv[level];
  const size_t npix = (size_t)g.w * g.h;
  const bool in_level = (size_t)p < npix; /* padding lanes still take part in the warp reduction below */
  if (!in_level) p = (int)npix - 1;
  const int v = p / g.w, u = p - v * g.w;
  const TsdfGeom& t = P.t;
  const short2* __restrict__ vol = P.vol + (size_t)s * t.dx * t.dy * t.dz;
  const float* T = P.world_f + s * 12;
  float R[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) R[k] = __ldg(T + k);

  float vx = 0.0f, vy = 0.0f, vz = 0.0f;
  const float ogx = (R[3] - t.ox) * t.inv_vs - 0.5f, ogy = (R[7] - t.oy) * t.inv_vs - 0.5f, ogz = (R[11] - t.oz) * t.inv_vs - 0.5f;
  const float dcx = ((float)u - g.cx) / g.fx, dcy = ((float)v - g.cy) / g.fy;
  const float dgx = __fmaf_rn(R[0], dcx, __fmaf_rn(R[1], dcy, R[2])) * t.inv_vs;
  const float dgy = __fmaf_rn(R[4], dcx, __fmaf_rn(R[5], dcy, R[6])) * t.inv_vs;
  const float dgz = __fmaf_rn(R[8], dcx, __fmaf_rn(R[9], dcy, R[10])) * t.inv_vs;
  float lmin = t.near_m, lmax = t.far_m;
  bool miss = false;
  {
    const float og[3] = {ogx, ogy, ogz}, dg[3] = {dgx, dgy, dgz};
    const float hi[3] = {(float)(t.dx - 1), (float)(t.dy - 1), (float)(t.dz - 1)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (dg[a] != 0.0f) {
        float l0 = (0.0f - og[a]) / dg[a], l1 = (hi[a] - og[a]) / dg[a];
        if (l0 > l1) {
          const float sw = l0;
          l0 = l1;
          l1 = sw;
        }
        if (l0 > lmin) lmin = l0;
        if (l1 < lmax) lmax = l1;
      } else if (!(og[a] >= 0.0f && og[a] <= hi[a])) {
        miss = true;
      }
    }
  }
  /* march-start hint: own reading, else the nearest reading of the warp's 32 consecutive pixels (level
   * sizes are multiples of 32 pixels or the group is cut at the end of the level, as in the specification);
   * evaluated by every lane before any lane leaves, so the warp reduction is complete */
  float hintD = 0.0f;
  if (P.hint) {
    const int slot = P.hint_slot >= 0 ? P.hint_slot : (__ldg(P.ring.head) + P.ring.R - 1) % P.ring.R;
    hintD = __ldg(P.depth[level] + ((size_t)s * P.ring.R + slot) * npix + p);
    /* positive floats order like their bit patterns; 0x7f800000 (inf) stands for "no reading" */
    const unsigned mine = (in_level && hintD > 0.0f) ? __float_as_uint(hintD) : 0x7f800000u;
    const unsigned gmin = __reduce_min_sync(0xffffffffu, mine);
    if (!(hintD > 0.0f) && gmin != 0x7f800000u) hintD = __uint_as_float(gmin);
  }
  if (!in_level) return;
  if (!miss && lmin < lmax) {
    float lam = lmin;
    if (hintD > 0.0f) {
      const float zh = hintD / P.depth_factor;
      const float start = zh - t.mu * 1.25f, stop = zh + t.mu * 2.0f;
      if (start > lam) lam = start;
      if (stop < lmax) lmax = stop; /* the surface fused from this reading is inside this window or nowhere */
    }
    float fprev = tsdf_nearest(vol, t, __fmaf_rn(lam, dgx, ogx), __fmaf_rn(lam, dgy, ogy), __fmaf_rn(lam, dgz, ogz));
    for (;;) {
      /* inside the truncation band in front of a surface the step shrinks with the distance (never below
       * 0.8 voxel), so the first negative sample lies close behind the surface, where every voxel is observed */
      float st = t.step;
      if (fprev > 0.0f && fprev <= 1.0f) {
        float a = fprev * t.mu;
        if (a < t.vs) a = t.vs;
        st = a * 0.8f;
      }
      const float lamn = lam + st;
      if (!(lamn < lmax)) break;
      const float f = tsdf_nearest(vol, t, __fmaf_rn(lamn, dgx, ogx), __fmaf_rn(lamn, dgy, ogy), __fmaf_rn(lamn, dgz, ogz));
      if (fprev < 0.0f && f > 0.0f) break;
      if (fprev > 0.0f && fprev <= 1.0f && f < 0.0f) {
        float Ft = 0.0f, Ftn = 0.0f;
        /* `&` not `&&`: both samples are fetched together (the functions are pure, the result is the same) */
        if ((tsdf_trilinear(vol, t, __fmaf_rn(lam, dgx, ogx), __fmaf_rn(lam, dgy, ogy), __fmaf_rn(lam, dgz, ogz), &Ft) &
             tsdf_trilinear(vol, t, __fmaf_rn(lamn, dgx, ogx), __fmaf_rn(lamn, dgy, ogy), __fmaf_rn(lamn, dgz, ogz), &Ftn)) &&
            Ft >= 0.0f && Ftn < 0.0f) {
          const float ls = lam - st * Ft / (Ftn - Ft);
          vz = ls; /* reference viewerModule.c:343-345 with z = the ray parameter */
          vx = ((float)u - g.cx) * ls / g.fx;
          vy = ((float)v - g.cy) * ls / g.fy;
          break;
        }
      }
      fprev = f;
      lam = lamn;
    }
  }
  /* planes 1.y and 2 (the normal) are completed by k_normals, run over the model maps right after */
  float2* base = P.model[level] + (size_t)s * 3 * npix;
  base[p] = make_float2(vx, vy);
  base[npix + p] = make_float2(vz, YK_N_INVALID);
}
