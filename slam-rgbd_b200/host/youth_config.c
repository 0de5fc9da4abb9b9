/*
 * youth_config.c -- YAML camera config and TUM trajectory egress.
 *
 * The reference hands its YAML to ORB-SLAM3 (Youth.Source/AlgorithmModule/SLAM.cpp:78-83);
 * the keys that matter to a depth tracker are Camera.fx/fy/cx/cy/width/height and
 * DepthMapFactor (config/astra_orb_slam3_rgbd.yaml:9-20,35).  Pose egress is TUM text
 * ("timestamp tx ty tz qx qy qz qw", SLAM.cpp:187-188).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "youth_host.h"

static int parse_key(const char* line, const char* key, double* out) {
  while (*line == ' ' || *line == '\t') ++line;
  size_t n = strlen(key);
  if (strncmp(line, key, n) != 0) return 0;
  line += n;
  while (*line == ' ' || *line == '\t') ++line;
  if (*line != ':') return 0;
  ++line;
  char* end = NULL;
  double v = strtod(line, &end);
  if (end == line) return 0;
  *out = v;
  return 1;
}

int youth_config_from_yaml(const char* path, youth_cuda_config* cfg) {
  if (!cfg) return 0;
  youth_cuda_default_config(cfg);
  if (!path || !*path) return 1;
  FILE* f = fopen(path, "r");
  if (!f) return 0;
  char line[512];
  double v;
  int ok = 1;
  /* a value the tracker cannot mean anything by (not finite, beyond float / image-size range) makes the file
   * invalid rather than being converted (out-of-range conversions are undefined behaviour in C) */
  while (fgets(line, sizeof(line), f)) {
    if (line[0] == '#' || line[0] == '%') continue;
    float* fdst = NULL;
    int32_t* idst = NULL;
    if (parse_key(line, "Camera.fx", &v)) fdst = &cfg->fx;
    else if (parse_key(line, "Camera.fy", &v)) fdst = &cfg->fy;
    else if (parse_key(line, "Camera.cx", &v)) fdst = &cfg->cx;
    else if (parse_key(line, "Camera.cy", &v)) fdst = &cfg->cy;
    else if (parse_key(line, "Camera.width", &v)) idst = &cfg->width;
    else if (parse_key(line, "Camera.height", &v)) idst = &cfg->height;
    else if (parse_key(line, "DepthMapFactor", &v)) {
      if (v >= 1e-30 && v < 1e30) cfg->depth_factor = (float)v; /* ORB-SLAM3 ignores a non-positive factor too */
      else if (!(v <= 0)) ok = 0;
      continue;
    }
    if (fdst) {
      if (v > -1e30 && v < 1e30) *fdst = (float)v;
      else ok = 0;
    } else if (idst) {
      if (v >= 1 && v <= 65535) *idst = (int32_t)v;
      else ok = 0;
    }
  }
  fclose(f);
  if (!ok) youth_cuda_default_config(cfg);
  return ok;
}

void youth_pose_to_quat(const float P[12], double q[4]) {
  const double m00 = P[0], m01 = P[1], m02 = P[2], m10 = P[4], m11 = P[5], m12 = P[6], m20 = P[8], m21 = P[9],
               m22 = P[10];
  const double tr = m00 + m11 + m22;
  double x, y, z, w;
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0) * 2.0;
    w = 0.25 * s;
    x = (m21 - m12) / s;
    y = (m02 - m20) / s;
    z = (m10 - m01) / s;
  } else if (m00 > m11 && m00 > m22) {
    const double s = sqrt(1.0 + m00 - m11 - m22) * 2.0;
    w = (m21 - m12) / s;
    x = 0.25 * s;
    y = (m01 + m10) / s;
    z = (m02 + m20) / s;
  } else if (m11 > m22) {
    const double s = sqrt(1.0 + m11 - m00 - m22) * 2.0;
    w = (m02 - m20) / s;
    x = (m01 + m10) / s;
    y = 0.25 * s;
    z = (m12 + m21) / s;
  } else {
    const double s = sqrt(1.0 + m22 - m00 - m11) * 2.0;
    w = (m10 - m01) / s;
    x = (m02 + m20) / s;
    y = (m12 + m21) / s;
    z = 0.25 * s;
  }
  const double n = sqrt(x * x + y * y + z * z + w * w);
  q[0] = x / n;
  q[1] = y / n;
  q[2] = z / n;
  q[3] = w / n;
}

int youth_tum_write(const char* path, const float* poses, const uint32_t* ts, int n) {
  if (!path || (!poses && n > 0)) return 0;
  FILE* f = fopen(path, "w");
  if (!f) return 0;
  for (int i = 0; i < n; ++i) {
    const float* P = poses + 12 * (size_t)i;
    double q[4];
    youth_pose_to_quat(P, q);
    const double sec = ts ? (double)ts[i] / 1000.0 : (double)i / 30.0;
    fprintf(f, "%.6f %.7f %.7f %.7f %.7f %.7f %.7f %.7f\n", sec, (double)P[3], (double)P[7], (double)P[11], q[0], q[1],
            q[2], q[3]);
  }
  return fclose(f) == 0;
}
