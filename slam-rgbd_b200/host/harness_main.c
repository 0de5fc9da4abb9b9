/*
 * harness_main.c -- headless driver replacing the interactive orchestrator (reference
 * Youth.Source/main.c:247-351 exits without a camera, main.c:273-277).
 *
 *   youth_harness gen  <out.bin> <frames> [sequence] [width height]   write a synthetic recording + <out.bin>.gt.txt
 *   youth_harness run  <in.bin> <out_prefix>                          replay through algorithmModule(), write TUM files
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "algorithmModule.h"
#include "youth_host.h"

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static int cmd_gen(int argc, char** argv) {
  if (argc < 4) return 2;
  const char* path = argv[2];
  const int n = atoi(argv[3]);
  const int seq = argc > 4 ? atoi(argv[4]) : 0;
  const int w = argc > 6 ? atoi(argv[5]) : 640, h = argc > 6 ? atoi(argv[6]) : 480;
  youth_synth_config sc;
  youth_synth_default(&sc, w, h, seq);
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  uint16_t* d = (uint16_t*)malloc((size_t)w * h * 2);
  float* gt = (float*)malloc(sizeof(float) * 12 * (size_t)n);
  uint32_t* ts = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    youth_synth_frame(&sc, i, d);
    ts[i] = (uint32_t)(33 * i);
    if (!youth_bin_write_frame(f, (uint32_t)i, ts[i], w, h, d, NULL)) return 1;
    double G[12];
    youth_synth_gt(&sc, i, G);
    for (int k = 0; k < 12; ++k) gt[12 * i + k] = (float)G[k];
  }
  youth_bin_write_eof(f);
  fclose(f);
  char gp[1024];
  snprintf(gp, sizeof(gp), "%s.gt.txt", path);
  youth_tum_write(gp, gt, ts, n);
  printf("{\"wrote\": \"%s\", \"frames\": %d, \"width\": %d, \"height\": %d}\n", path, n, w, h);
  free(d);
  free(gt);
  free(ts);
  return 0;
}

static int cmd_run(int argc, char** argv) {
  if (argc < 4) return 2;
  setenv("YOUTH_SLAM_OUT", argv[3], 1);
  pthread_t th;
  const double t0 = now_s();
  if (pthread_create(&th, NULL, algorithmModule, argv[2]) != 0) return 1;
  pthread_join(th, NULL);
  printf("{\"replayed\": \"%s\", \"seconds\": %.3f, \"trajectory\": \"%s_trajectory.txt\"}\n", argv[2], now_s() - t0,
         argv[3]);
  return 0;
}

int main(int argc, char** argv) {
  int rc = 2;
  if (argc >= 2 && !strcmp(argv[1], "gen")) rc = cmd_gen(argc, argv);
  else if (argc >= 2 && !strcmp(argv[1], "run")) rc = cmd_run(argc, argv);
  if (rc == 2) fprintf(stderr, "usage: %s gen <out.bin> <frames> [sequence] [w h] | run <in.bin> <out_prefix>\n", argv[0]);
  return rc;
}
