/*
 * youth_synth.c -- synthetic Astra-shaped depth sequences with ground-truth poses
 * (SURVEY.md section 8(d)).  Stands in for the SensorModule (reference
 * Youth.Source/SensorModule/sensorModule.c:112-244 needs a physical camera and the
 * proprietary Astra SDK).
 *
 * Scene: closed room 5 x 3 x 6 m (x in [-2.5,2.5], y in [-1.5,1.5], z in [-1,5]; the 6 m axis runs
 * along the initial viewing direction so side walls, floor and ceiling are in view) with two
 * spheres and a tilted disc, camera inside, so every ray hits and all six pose
 * directions are constrained.  Pinhole intrinsics as in reference
 * config/astra_orb_slam3_rgbd.yaml:9-12.  Depth is z (not range) in mm, rounded to
 * nearest; 0 outside [dmin, dmax] and where the lowbias32 hash drops the pixel.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "youth_host.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

void youth_synth_default(youth_synth_config* c, int width, int height, int sequence) {
  memset(c, 0, sizeof(*c));
  c->width = width;
  c->height = height;
  const double s = (double)width / 640.0;
  c->fx = 570.3 * s;
  c->fy = 570.3 * s;
  c->cx = 320.0 * s;
  c->cy = 240.0 * s;
  c->seed = 20261018u + (uint32_t)sequence;
  c->period = 300;
  c->phase = 2.0 * M_PI * (double)sequence / 64.0;
  c->dropout = 0.02;
  c->noise = 0;
  c->dmin_mm = 600;
  c->dmax_mm = 8000;
}

void youth_synth_pose(const youth_synth_config* c, int frame, double T[12]) {
  const double a = 2.0 * M_PI * (double)frame / (double)c->period;
  const double p = c->phase;
  const double tx = 0.30 * sin(a + p) - 0.30 * sin(p);
  const double ty = 0.10 * sin(2.0 * a + p) - 0.10 * sin(p);
  const double tz = 0.20 * (1.0 - cos(a + p)) - 0.20 * (1.0 - cos(p));
  const double d2r = M_PI / 180.0;
  const double yaw = 10.0 * d2r * sin(a + p);
  const double pitch = 5.0 * d2r * sin(2.0 * a + p);
  const double roll = 3.0 * d2r * sin(3.0 * a + p);
  const double cyw = cos(yaw), syw = sin(yaw), cp = cos(pitch), sp = sin(pitch), cr = cos(roll), sr = sin(roll);
  /* R = Ry(yaw) * Rx(pitch) * Rz(roll) */
  const double Ry[9] = {cyw, 0, syw, 0, 1, 0, -syw, 0, cyw};
  const double Rx[9] = {1, 0, 0, 0, cp, -sp, 0, sp, cp};
  const double Rz[9] = {cr, -sr, 0, sr, cr, 0, 0, 0, 1};
  double M[9], R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[3 * i + j] = Ry[3 * i] * Rx[j] + Ry[3 * i + 1] * Rx[3 + j] + Ry[3 * i + 2] * Rx[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = M[3 * i] * Rz[j] + M[3 * i + 1] * Rz[3 + j] + M[3 * i + 2] * Rz[6 + j];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
  }
  T[3] = tx;
  T[7] = ty;
  T[11] = tz;
}

void youth_synth_gt(const youth_synth_config* c, int frame, double G[12]) {
  double A[12], B[12];
  youth_synth_pose(c, 0, A);
  youth_synth_pose(c, frame, B);
  /* G = A^-1 * B,  A^-1 = [R^T | -R^T t] */
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) G[4 * i + j] = A[i] * B[j] + A[4 + i] * B[4 + j] + A[8 + i] * B[8 + j];
    const double dx = B[3] - A[3], dy = B[7] - A[7], dz = B[11] - A[11];
    G[4 * i + 3] = A[i] * dx + A[4 + i] * dy + A[8 + i] * dz;
  }
}

static double hit_scene(const double o[3], const double d[3]) {
  double best = 1e30;
  /* room: x = +-2.5, y = +-1.5, z = -1 / 5 */
  const double lo[3] = {-2.5, -1.5, -1.0}, hi[3] = {2.5, 1.5, 5.0};
  for (int a = 0; a < 3; ++a) {
    if (d[a] > 1e-12) {
      const double s = (hi[a] - o[a]) / d[a];
      if (s > 0 && s < best) best = s;
    } else if (d[a] < -1e-12) {
      const double s = (lo[a] - o[a]) / d[a];
      if (s > 0 && s < best) best = s;
    }
  }
  /* spheres */
  static const double sph[2][4] = {{-1.0, 0.4, 2.6, 0.5}, {1.1, -0.5, 2.0, 0.3}};
  const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  for (int k = 0; k < 2; ++k) {
    const double ox = o[0] - sph[k][0], oy = o[1] - sph[k][1], oz = o[2] - sph[k][2];
    const double b = ox * d[0] + oy * d[1] + oz * d[2];
    const double cc = ox * ox + oy * oy + oz * oz - sph[k][3] * sph[k][3];
    const double disc = b * b - dd * cc;
    if (disc > 0) {
      const double s = (-b - sqrt(disc)) / dd;
      if (s > 0 && s < best) best = s;
    }
  }
  /* tilted disc */
  {
    static const double p0[3] = {0.2, 0.8, 2.9};
    const double nl = sqrt(0.25 * 0.25 + 0.85 * 0.85 + 0.46 * 0.46);
    const double n[3] = {0.25 / nl, -0.85 / nl, -0.46 / nl};
    const double den = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
    if (fabs(den) > 1e-12) {
      const double s = (n[0] * (p0[0] - o[0]) + n[1] * (p0[1] - o[1]) + n[2] * (p0[2] - o[2])) / den;
      if (s > 0 && s < best) {
        const double hx = o[0] + s * d[0] - p0[0], hy = o[1] + s * d[1] - p0[1], hz = o[2] + s * d[2] - p0[2];
        if (hx * hx + hy * hy + hz * hz < 0.9 * 0.9) best = s;
      }
    }
  }
  return best;
}

struct synth_job {
  const youth_synth_config* c;
  double T[12];
  uint32_t fseed, drop_thr;
  int v0, v1;
  uint16_t* out;
};

static void* synth_rows(void* arg) {
  const struct synth_job* j = (const struct synth_job*)arg;
  const youth_synth_config* c = j->c;
  const double* T = j->T;
  const double o[3] = {T[3], T[7], T[11]};
  const int W = c->width;
  for (int v = j->v0; v < j->v1; ++v) {
    for (int u = 0; u < W; ++u) {
      const double xc = ((double)u - c->cx) / c->fx, yc = ((double)v - c->cy) / c->fy;
      double d[3];
      for (int i = 0; i < 3; ++i) d[i] = T[4 * i] * xc + T[4 * i + 1] * yc + T[4 * i + 2];
      const double s = hit_scene(o, d); /* camera-space direction has z = 1, so s is the z depth */
      long mm = lrint(s * 1000.0);
      const uint32_t h = lowbias32(j->fseed ^ (uint32_t)(v * W + u));
      if (c->noise) mm += (long)((h >> 8) % 5u) - 2;
      uint16_t val = 0;
      if (mm >= c->dmin_mm && mm <= c->dmax_mm && h >= j->drop_thr) val = (uint16_t)mm;
      j->out[(size_t)v * W + u] = val;
    }
  }
  return NULL;
}

void youth_synth_frame(const youth_synth_config* c, int frame, uint16_t* out) {
  enum { MAX_T = 32 };
  struct synth_job jobs[MAX_T];
  pthread_t th[MAX_T];
  int nt = 0;
  const char* e = getenv("YOUTH_SYNTH_THREADS");
  if (e) nt = atoi(e);
  if (nt <= 0) nt = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nt > MAX_T) nt = MAX_T;
  if (nt > c->height) nt = c->height;
  if (nt < 1) nt = 1;
  for (int t = 0; t < nt; ++t) {
    jobs[t].c = c;
    youth_synth_pose(c, frame, jobs[t].T);
    jobs[t].fseed = lowbias32(c->seed + (uint32_t)frame * 0x9E3779B9u);
    jobs[t].drop_thr = (uint32_t)(c->dropout * 4294967296.0);
    jobs[t].v0 = (int)((long)c->height * t / nt);
    jobs[t].v1 = (int)((long)c->height * (t + 1) / nt);
    jobs[t].out = out;
  }
  for (int t = 1; t < nt; ++t)
    if (pthread_create(&th[t], NULL, synth_rows, &jobs[t]) != 0) {
      synth_rows(&jobs[t]);
      th[t] = 0;
    }
  synth_rows(&jobs[0]);
  for (int t = 1; t < nt; ++t)
    if (th[t]) pthread_join(th[t], NULL);
}

void youth_synth_sequence(const youth_synth_config* c, int first, int n, uint16_t* out) {
  const size_t stride = (size_t)c->width * c->height;
  for (int i = 0; i < n; ++i) youth_synth_frame(c, first + i, out + stride * (size_t)i);
}
