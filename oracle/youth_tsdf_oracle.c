/*
 * youth_tsdf_oracle.c -- CPU statement of frame-to-MODEL tracking:
 * TSDF volume integration, ray casting of the model into vertex/normal maps, and the sequence
 * tracker that aligns every frame to the ray-cast model instead of the previous frame.
 * TEST INFRASTRUCTURE ONLY (same rules as youth_oracle.c).
 *
 * PARITY UNPINNED: the reference has no volumetric model at all; this
 * file DEFINES the arithmetic, the device (slam-rgbd_b200/csrc/youth_model.cuh) mirrors it bit for
 * bit.  Stages 3-5 (association, reduction, solve) are the ones of youth_oracle.c, unchanged:
 * the model maps simply take the place of the previous frame's maps.  What the reference pins is
 * still followed: depth unit and validity (viewerModule.c:341-343), the pinhole model
 * (viewerModule.c:344-345: x = (u - cx) z / fx), used both to project voxels and to place the
 * ray-cast vertices.
 *
 * Arithmetic rules as in youth_oracle.h: float, one rounding per written operation
 * (-ffp-contract=off), fmaf() exactly where written, IEEE division / sqrt, lrintf() = round to
 * nearest even (cvt.rni on the device), no other libm.
 *
 * Volume: dim[0] x dim[1] x dim[2] voxels, x fastest; voxel (ix,iy,iz) has its centre at
 * origin + (i + 0.5) * voxel_m in the world (= first camera) frame and stores
 * (int16 tsdf * 32767, int16 weight); a fresh volume is (32767, 0) everywhere.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "youth_oracle.h"

#define YT_SCALE 32767.0f
#define YT_INV_SCALE (1.0f / 32767.0f)
#define YT_UNKNOWN 2.0f /* "no observation" in the marching loop (a tsdf is within [-1, 1]) */

typedef struct yo_tsdf_config {
  int32_t dim[3];
  float voxel_m;
  float origin[3];
  float trunc_m;
  int32_t max_weight;
  float near_m, far_m;
} yo_tsdf_config;

void yo_tsdf_default_config(yo_tsdf_config* t) {
  memset(t, 0, sizeof(*t));
  t->dim[0] = 256;
  t->dim[1] = 128;
  t->dim[2] = 256;
  t->voxel_m = 0.025f;
  t->origin[0] = -3.2f;
  t->origin[1] = -1.6f;
  t->origin[2] = -1.2f;
  t->trunc_m = 0.1f;
  t->max_weight = 64;
  t->near_m = 0.4f;
  t->far_m = 8.0f;
}

size_t yo_tsdf_voxels(const yo_tsdf_config* t) { return (size_t)t->dim[0] * t->dim[1] * t->dim[2]; }

void yo_tsdf_clear(const yo_tsdf_config* t, int16_t* vox) {
  const size_t n = yo_tsdf_voxels(t);
  for (size_t i = 0; i < n; ++i) {
    vox[2 * i] = 32767;
    vox[2 * i + 1] = 0;
  }
}

/* Fuse one frame: depth0 = level-0 depth of the frame in raw units (0 = invalid, the output of
 * stage 1), pose = camera-to-world float 3x4.  Projective signed distance along the optical axis,
 * truncated at trunc_m, running average with unit weights. */
void yo_tsdf_integrate(const yo_config* c, const yo_tsdf_config* t, int16_t* vox, const float* depth0,
                       const float pose[12]) {
  yo_level g;
  yo_level_geometry(c, 0, &g);
  const float cxh = g.cx + 0.5f, cyh = g.cy + 0.5f;
  /* world -> camera: R^T and -R^T t */
  float Ri[9], ti[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Ri[3 * i + j] = pose[4 * j + i];
    ti[i] = -fmaf(pose[i], pose[3], fmaf(pose[4 + i], pose[7], pose[8 + i] * pose[11]));
  }
  const float vs = t->voxel_m, mu = t->trunc_m;
  for (int iz = 0; iz < t->dim[2]; ++iz) {
    const float wz = fmaf((float)iz + 0.5f, vs, t->origin[2]);
    for (int iy = 0; iy < t->dim[1]; ++iy) {
      const float wy = fmaf((float)iy + 0.5f, vs, t->origin[1]);
      for (int ix = 0; ix < t->dim[0]; ++ix) {
        const float wx = fmaf((float)ix + 0.5f, vs, t->origin[0]);
        const float px = fmaf(Ri[0], wx, fmaf(Ri[1], wy, fmaf(Ri[2], wz, ti[0])));
        const float py = fmaf(Ri[3], wx, fmaf(Ri[4], wy, fmaf(Ri[5], wz, ti[1])));
        const float pz = fmaf(Ri[6], wx, fmaf(Ri[7], wy, fmaf(Ri[8], wz, ti[2])));
        if (!(pz > 0.0f)) continue;
        const float iz1 = 1.0f / pz;
        const float ur = fmaf(px * g.fx, iz1, cxh);
        const float vr = fmaf(py * g.fy, iz1, cyh);
        if (!(ur >= 0.0f && ur < (float)g.w && vr >= 0.0f && vr < (float)g.h)) continue;
        const float D = depth0[(int)vr * g.w + (int)ur];
        if (!(D > 0.0f)) continue;
        const float sdf = D / c->depth_factor - pz;
        if (!(sdf >= -mu)) continue;
        float f = sdf / mu;
        if (f > 1.0f) f = 1.0f;
        int16_t* v = vox + 2 * (((size_t)iz * t->dim[1] + iy) * t->dim[0] + ix);
        const float F = (float)v[0] * YT_INV_SCALE, W = (float)v[1];
        const float Fn = (F * W + f) / (W + 1.0f);
        v[0] = (int16_t)lrintf(Fn * YT_SCALE);
        v[1] = (int16_t)(v[1] + 1 > t->max_weight ? t->max_weight : v[1] + 1);
      }
    }
  }
}

/* Size of the map: observed voxels (weight > 0) whose tsdf changes sign (>= 0 against < 0) towards an observed
 * +x, +y or +z neighbour -- the voxels the surface passes through.  The dense counterpart of
 * GetAllMapPoints().size() (reference SLAM.cpp:212-217). */
int64_t yo_tsdf_surface_voxels(const yo_tsdf_config* t, const int16_t* vox) {
  const int dx = t->dim[0], dy = t->dim[1], dz = t->dim[2];
  int64_t n = 0;
  for (int iz = 0; iz < dz; ++iz)
    for (int iy = 0; iy < dy; ++iy)
      for (int ix = 0; ix < dx; ++ix) {
        const int16_t* v = vox + 2 * (((size_t)iz * dy + iy) * dx + ix);
        if (v[1] <= 0) continue;
        const int neg = v[0] < 0;
        int hit = 0;
        if (ix + 1 < dx && v[3] > 0 && (v[2] < 0) != neg) hit = 1;
        if (iy + 1 < dy && v[2 * dx + 1] > 0 && (v[2 * dx] < 0) != neg) hit = 1;
        if (iz + 1 < dz && v[2 * (size_t)dx * dy + 1] > 0 && (v[2 * (size_t)dx * dy] < 0) != neg) hit = 1;
        n += hit;
      }
  return n;
}

/* nearest-voxel sample at grid coordinates (voxel centres at integers); YT_UNKNOWN where unobserved */
static float tsdf_nearest(const yo_tsdf_config* t, const int16_t* vox, float gx, float gy, float gz) {
  int ix = (int)(gx + 0.5f), iy = (int)(gy + 0.5f), iz = (int)(gz + 0.5f);
  ix = ix < 0 ? 0 : (ix > t->dim[0] - 1 ? t->dim[0] - 1 : ix);
  iy = iy < 0 ? 0 : (iy > t->dim[1] - 1 ? t->dim[1] - 1 : iy);
  iz = iz < 0 ? 0 : (iz > t->dim[2] - 1 ? t->dim[2] - 1 : iz);
  const int16_t* v = vox + 2 * (((size_t)iz * t->dim[1] + iy) * t->dim[0] + ix);
  return v[1] > 0 ? (float)v[0] * YT_INV_SCALE : YT_UNKNOWN;
}

/* trilinear sample; returns 0 when a corner is outside the volume or unobserved */
static int tsdf_trilinear(const yo_tsdf_config* t, const int16_t* vox, float gx, float gy, float gz, float* out) {
  if (!(gx >= 0.0f && gy >= 0.0f && gz >= 0.0f)) return 0;
  if (!(gx < (float)(t->dim[0] - 1) && gy < (float)(t->dim[1] - 1) && gz < (float)(t->dim[2] - 1))) return 0;
  const int ix = (int)gx, iy = (int)gy, iz = (int)gz;
  const float fx = gx - (float)ix, fy = gy - (float)iy, fz = gz - (float)iz;
  float c[8];
  for (int k = 0; k < 8; ++k) { /* k = dz*4 + dy*2 + dx */
    const int16_t* v = vox + 2 * (((size_t)(iz + (k >> 2)) * t->dim[1] + (iy + ((k >> 1) & 1))) * t->dim[0] + (ix + (k & 1)));
    if (v[1] <= 0) return 0;
    c[k] = (float)v[0] * YT_INV_SCALE;
  }
  const float c00 = c[0] + fx * (c[1] - c[0]);
  const float c10 = c[2] + fx * (c[3] - c[2]);
  const float c01 = c[4] + fx * (c[5] - c[4]);
  const float c11 = c[6] + fx * (c[7] - c[6]);
  const float c0 = c00 + fy * (c10 - c00);
  const float c1 = c01 + fy * (c11 - c01);
  *out = c0 + fz * (c1 - c0);
  return 1;
}

/* Ray-cast the model from `pose` (camera-to-world) at pyramid level `level` into float4 vertex and
 * normal maps (x, y, z, valid) expressed in THAT camera's frame, i.e. the maps a frame taken from
 * `pose` would have -- which is what stage 3 expects of a "previous frame".  The ray parameter is
 * the depth z: point(l) = l * ((u - cx)/fx, (v - cy)/fy, 1) in the camera frame.
 * hint (nullable): a depth map of this level (raw units, 0 = invalid) taken from `pose` -- the tracker
 * passes the frame it has just fused.  Where it has a reading (the pixel's own, else the nearest one in
 * its group of 32 consecutive pixels) the march starts 1.25 mu in front of it and gives up 2 mu behind it,
 * instead of walking from the near plane to the far end of the volume (the surface seen from this pose is there), which removes the walk through
 * empty space; everything else is unchanged. */
void yo_tsdf_raycast(const yo_config* c, const yo_tsdf_config* t, const int16_t* vox, const float pose[12], int level,
                     const float* hint, float* vmap, float* nmap) {
  yo_level g;
  yo_level_geometry(c, level, &g);
  const float inv_vs = 1.0f / t->voxel_m;
  const float step = t->trunc_m * 0.8f;
  const float og[3] = {(pose[3] - t->origin[0]) * inv_vs - 0.5f, (pose[7] - t->origin[1]) * inv_vs - 0.5f,
                       (pose[11] - t->origin[2]) * inv_vs - 0.5f};
  for (int v = 0; v < g.h; ++v) {
    for (int u = 0; u < g.w; ++u) {
      float* vo = vmap + 4 * ((size_t)v * g.w + u);
      float* no = nmap + 4 * ((size_t)v * g.w + u);
      vo[0] = vo[1] = vo[2] = vo[3] = 0.0f;
      no[0] = no[1] = no[2] = no[3] = 0.0f;
      const float dcx = ((float)u - g.cx) / g.fx, dcy = ((float)v - g.cy) / g.fy;
      float dg[3];
      dg[0] = fmaf(pose[0], dcx, fmaf(pose[1], dcy, pose[2])) * inv_vs;
      dg[1] = fmaf(pose[4], dcx, fmaf(pose[5], dcy, pose[6])) * inv_vs;
      dg[2] = fmaf(pose[8], dcx, fmaf(pose[9], dcy, pose[10])) * inv_vs;
      float lmin = t->near_m, lmax = t->far_m;
      int miss = 0;
      for (int a = 0; a < 3; ++a) {
        const float hi = (float)(t->dim[a] - 1);
        if (dg[a] != 0.0f) {
          float l0 = (0.0f - og[a]) / dg[a], l1 = (hi - og[a]) / dg[a];
          if (l0 > l1) {
            const float s = l0;
            l0 = l1;
            l1 = s;
          }
          if (l0 > lmin) lmin = l0;
          if (l1 < lmax) lmax = l1;
        } else if (!(og[a] >= 0.0f && og[a] <= hi)) {
          miss = 1;
        }
      }
      if (miss || !(lmin < lmax)) continue;
      float lam = lmin;
      if (hint != NULL) {
        /* own reading, else the nearest reading among the 32 consecutive pixels (row-major index) this pixel
         * belongs to -- the group one warp ray-casts on the device, so no lane walks alone from the near plane */
        float D = hint[(size_t)v * g.w + u];
        if (!(D > 0.0f)) {
          const size_t p = (size_t)v * g.w + u, g0 = p & ~(size_t)31, np = (size_t)g.w * g.h;
          D = 0.0f;
          for (size_t q = g0; q < g0 + 32 && q < np; ++q)
            if (hint[q] > 0.0f && (D == 0.0f || hint[q] < D)) D = hint[q];
        }
        if (D > 0.0f) {
          const float zh = D / c->depth_factor;
          const float start = zh - t->trunc_m * 1.25f, stop = zh + t->trunc_m * 2.0f;
          if (start > lam) lam = start;
          if (stop < lmax) lmax = stop; /* the surface fused from this reading is inside this window or nowhere */
        }
      }
      float fprev = tsdf_nearest(t, vox, fmaf(lam, dg[0], og[0]), fmaf(lam, dg[1], og[1]), fmaf(lam, dg[2], og[2]));
      for (;;) {
        /* inside the truncation band in front of a surface the step shrinks with the distance (never below
         * 0.8 voxel), so the first negative sample lies close behind the surface, where every voxel is observed */
        float st = step;
        if (fprev > 0.0f && fprev <= 1.0f) {
          float a = fprev * t->trunc_m;
          if (a < t->voxel_m) a = t->voxel_m;
          st = a * 0.8f;
        }
        const float lamn = lam + st;
        if (!(lamn < lmax)) break;
        const float f = tsdf_nearest(t, vox, fmaf(lamn, dg[0], og[0]), fmaf(lamn, dg[1], og[1]), fmaf(lamn, dg[2], og[2]));
        if (fprev < 0.0f && f > 0.0f) break; /* leaving a surface from behind (or into the unknown) */
        if (fprev > 0.0f && fprev <= 1.0f && f < 0.0f) {
          float Ft, Ftn;
          if (tsdf_trilinear(t, vox, fmaf(lam, dg[0], og[0]), fmaf(lam, dg[1], og[1]), fmaf(lam, dg[2], og[2]), &Ft) &&
              tsdf_trilinear(t, vox, fmaf(lamn, dg[0], og[0]), fmaf(lamn, dg[1], og[1]), fmaf(lamn, dg[2], og[2]), &Ftn) &&
              Ft >= 0.0f && Ftn < 0.0f) {
            const float ls = lam - st * Ft / (Ftn - Ft);
            vo[2] = ls; /* viewerModule.c:343-345 with z = the ray parameter */
            vo[0] = ((float)u - g.cx) * ls / g.fx;
            vo[1] = ((float)v - g.cy) * ls / g.fy;
            vo[3] = 1.0f;
            break;
          }
        }
        fprev = f;
        lam = lamn;
      }
    }
  }
  /* model normals: the normal half of stage 2 applied to the ray-cast vertex map (same convention as the
   * frame maps by construction, and a third of the volume reads of a tsdf-gradient normal) */
  yo_normals_from_vertices(g.w, g.h, vmap, nmap);
}

/* ------------------------------------------------------------------ frame-to-model sequence tracker */

typedef struct yo_model_tracker {
  yo_config cfg;
  yo_tsdf_config tcfg;
  int16_t* vox;
  yo_frame* cur;
  yo_frame* model; /* vmap / nmap only */
  int count;
  double world[12];
  int32_t inliers;
} yo_model_tracker;

yo_model_tracker* yo_model_tracker_create(const yo_config* c, const yo_tsdf_config* t) {
  yo_model_tracker* m = (yo_model_tracker*)calloc(1, sizeof(*m));
  m->cfg = *c;
  m->tcfg = *t;
  m->vox = (int16_t*)malloc(sizeof(int16_t) * 2 * yo_tsdf_voxels(t));
  m->cur = yo_frame_alloc(c);
  m->model = yo_frame_alloc(c);
  yo_tsdf_clear(t, m->vox);
  for (int i = 0; i < 12; ++i) m->world[i] = (i == 0 || i == 5 || i == 10) ? 1.0 : 0.0;
  return m;
}

void yo_model_tracker_destroy(yo_model_tracker* m) {
  if (!m) return;
  free(m->vox);
  yo_frame_free(m->cur);
  yo_frame_free(m->model);
  free(m);
}

const int16_t* yo_model_tracker_volume(const yo_model_tracker* m) { return m->vox; }
const yo_frame* yo_model_tracker_model(const yo_model_tracker* m) { return m->model; }

/* One frame: stages 1-2, then stages 3-5 against the ray-cast model (yo_track_pair, unchanged),
 * pose chain, fusion of the frame at its new pose (skipped when tracking was flagged lost), and
 * the ray cast of every pyramid level from the new pose for the next frame. */
uint32_t yo_model_tracker_track(yo_model_tracker* m, const uint16_t* raw, float pose_out[12]) {
  yo_preprocess(&m->cfg, raw, m->cur);
  uint32_t status;
  if (m->count == 0) {
    status = YO_STATUS_FIRST;
    m->inliers = 0;
  } else {
    double rel[12];
    status = yo_track_pair(&m->cfg, m->cur, m->model, rel, &m->inliers);
    yo_compose(m->world, rel, m->world);
  }
  float pf[12];
  for (int i = 0; i < 12; ++i) pf[i] = (float)m->world[i];
  if (!(status & YO_STATUS_LOST)) yo_tsdf_integrate(&m->cfg, &m->tcfg, m->vox, m->cur->depth[0], pf);
  for (int l = 0; l < m->cfg.levels; ++l)
    yo_tsdf_raycast(&m->cfg, &m->tcfg, m->vox, pf, l, m->cur->depth[l], m->model->vmap[l], m->model->nmap[l]);
  m->count++;
  if (pose_out) memcpy(pose_out, pf, sizeof(pf));
  return status;
}

int32_t yo_model_tracker_last_inliers(const yo_model_tracker* m) { return m->inliers; }

void yo_track_sequence_model(const yo_config* c, const yo_tsdf_config* t, const uint16_t* frames, int n, float* poses_out,
                             uint32_t* status_out) {
  yo_model_tracker* m = yo_model_tracker_create(c, t);
  const size_t stride = (size_t)c->width * c->height;
  for (int i = 0; i < n; ++i) {
    const uint32_t st = yo_model_tracker_track(m, frames + stride * i, poses_out ? poses_out + 12 * (size_t)i : NULL);
    if (status_out) status_out[i] = st;
  }
  yo_model_tracker_destroy(m);
}
