/*
 * harness_main.c -- headless driver replacing the interactive orchestrator (reference
 * Youth.Source/main.c:247-351 exits without a camera, main.c:273-277).
 *
 *   youth_harness gen  <out.bin> <frames> [sequence] [width height]   write a synthetic recording + <out.bin>.gt.txt
 *   youth_harness run  <in.bin> <out_prefix>                          replay through algorithmModule(), write TUM files
 *   youth_harness run  mq:/logger_viewer_queue <out_prefix>           track what the reference's logger / playbackThread put on
 *                                                                     the viewer queue (until YOUTH_SLAM_MQ_IDLE_MS of silence)
 *   youth_harness pack <in.bin> <out.bin>                             transcode raw depth records to YD16-packed records (GPU codec)
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
  if (n < 1 || w < 1 || h < 1) return 2;
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  uint16_t* d = (uint16_t*)malloc((size_t)w * h * 2);
  float* gt = (float*)malloc(sizeof(float) * 12 * (size_t)n);
  uint32_t* ts = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
  int failed = !d || !gt || !ts;
  for (int i = 0; i < n && !failed; ++i) {
    youth_synth_frame(&sc, i, d);
    ts[i] = (uint32_t)(33 * i);
    if (!youth_bin_write_frame(f, (uint32_t)i, ts[i], w, h, d, NULL)) {
      failed = 1;
      break;
    }
    double G[12];
    youth_synth_gt(&sc, i, G);
    for (int k = 0; k < 12; ++k) gt[12 * i + k] = (float)G[k];
  }
  if (!failed) youth_bin_write_eof(f);
  fclose(f);
  if (failed) {
    fprintf(stderr, "gen: cannot write %s\n", path);
    free(d);
    free(gt);
    free(ts);
    return 1;
  }
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

/* raw recording -> packed recording: depth payloads through youth_codec_encode in runs of 32 frames;
 * colour payloads are carried over unchanged */
static int cmd_pack(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* in = fopen(argv[2], "rb");
  FILE* out = in ? fopen(argv[3], "wb") : NULL;
  if (!in || !out) {
    fprintf(stderr, "pack: cannot open files\n");
    if (in) fclose(in);
    return 1;
  }
  enum { RUN = 32 };
  youth_codec* cd = NULL;
  uint16_t* depth = NULL;
  uint8_t *color = NULL, *packed = NULL;
  uint64_t offs[RUN + 1];
  FrameHeader hdr[RUN];
  size_t fpx = 0, max_b = 0;
  long frames = 0;
  unsigned long long raw_bytes = 0, packed_bytes = 0;
  int rc = 0, more = 1;
  while (more && rc == 0) {
    int n = 0;
    while (n < RUN) {
      FrameHeader h;
      long pos = ftell(in);
      if (fread(&h, sizeof(h), 1, in) != 1 || h.frameType == FRAME_TYPE_END_OF_FILE) {
        more = 0;
        break;
      }
      if (h.frameType != FRAME_TYPE_DEPTH_COLOR || h.depthDataSize != (uint32_t)h.width * h.height * 2u) {
        fprintf(stderr, "pack: record %ld is not a raw depth record\n", frames + n);
        rc = 1;
        break;
      }
      if (!cd) {
        fpx = (size_t)h.width * h.height;
        max_b = youth_codec_max_bytes(h.width, h.height);
        if (!youth_codec_create(h.width, h.height, RUN, 0, &cd)) {
          fprintf(stderr, "pack: %s\n", youth_cuda_last_error());
          rc = 1;
          break;
        }
        depth = (uint16_t*)malloc(fpx * 2 * RUN);
        color = (uint8_t*)malloc(fpx * 3 * RUN);
        packed = (uint8_t*)malloc(max_b * RUN);
        if (!depth || !color || !packed) {
          fprintf(stderr, "pack: out of memory\n");
          rc = 1;
          break;
        }
      } else if (fpx != (size_t)h.width * h.height) {
        rc = 1;
        break;
      }
      (void)pos;
      if (h.depthDataSize != fpx * 2) { /* raw 16-bit records only (an already packed recording has nothing to gain) */
        fprintf(stderr, "pack: record %u is not a raw depth frame (%u payload bytes)\n", h.frameId, h.depthDataSize);
        rc = 1;
        break;
      }
      if (fread(depth + fpx * n, 1, h.depthDataSize, in) != h.depthDataSize) { more = 0; break; }
      if (h.colorDataSize > fpx * 3 || fread(color + fpx * 3 * n, 1, h.colorDataSize, in) != h.colorDataSize) { more = 0; break; }
      hdr[n++] = h;
    }
    if (n == 0 || rc) break;
    if (!youth_codec_encode(cd, depth, YOUTH_MEM_HOST, n, packed, max_b * RUN, offs)) {
      fprintf(stderr, "pack: %s\n", youth_cuda_last_error());
      rc = 1;
      break;
    }
    for (int i = 0; i < n; ++i) {
      if (!youth_bin_write_packed_frame(out, hdr[i].frameId, hdr[i].timestamp, hdr[i].width, hdr[i].height,
                                        packed + offs[i], (uint32_t)(offs[i + 1] - offs[i]),
                                        hdr[i].colorDataSize == fpx * 3 ? color + fpx * 3 * i : NULL))
        rc = 1;
      raw_bytes += hdr[i].depthDataSize;
      packed_bytes += offs[i + 1] - offs[i];
    }
    frames += n;
  }
  if (rc == 0) youth_bin_write_eof(out);
  fclose(in);
  fclose(out);
  youth_codec_destroy(cd);
  free(depth);
  free(color);
  free(packed);
  if (rc == 0)
    printf("{\"packed\": \"%s\", \"frames\": %ld, \"depth_bytes_raw\": %llu, \"depth_bytes_packed\": %llu}\n", argv[3], frames,
           raw_bytes, packed_bytes);
  return rc;
}

int main(int argc, char** argv) {
  int rc = 2;
  if (argc >= 2 && !strcmp(argv[1], "gen")) rc = cmd_gen(argc, argv);
  else if (argc >= 2 && !strcmp(argv[1], "run")) rc = cmd_run(argc, argv);
  else if (argc >= 2 && !strcmp(argv[1], "pack")) rc = cmd_pack(argc, argv);
  if (rc == 2)
    fprintf(stderr, "usage: %s gen <out.bin> <frames> [sequence] [w h] | run <in.bin | mq:/queue> <out_prefix> | pack <in.bin> <out.bin>\n",
            argv[0]);
  return rc;
}
