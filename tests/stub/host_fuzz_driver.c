/*
 * host_fuzz_driver.c -- TEST INFRASTRUCTURE: random and mutated inputs against the host-side parsers that face
 * data from outside the process -- mq messages (youth_reasm_feed, youth_pose_msg_parse), .bin recordings
 * (youth_bin_read_frame) and the camera YAML (youth_config_from_yaml) -- built with -fsanitize=address,undefined.
 * The parsers must reject or accept, never touch memory they do not own.  usage: host_fuzz_driver <scratch dir>
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "youth_host.h"

static uint32_t rs = 12345u;
static uint32_t rnd(void) {
  rs ^= rs << 13;
  rs ^= rs >> 17;
  rs ^= rs << 5;
  return rs;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  char path[1024];
  /* 1. message reassembly: valid chunk streams with mutated headers, interleaved sizes, truncated messages */
  youth_reasm* r = youth_reasm_create();
  unsigned char* msg = (unsigned char*)malloc(MAX_MSG_SIZE);
  uint16_t depth[48 * 64];
  uint8_t color[48 * 64 * 3];
  memset(depth, 7, sizeof(depth));
  memset(color, 9, sizeof(color));
  long frames = 0, rejected = 0;
  for (int it = 0; it < 60000; ++it) {
    const int w = 8 << (rnd() % 4), h = 6 << (rnd() % 4); /* up to 64x48 */
    const int kind = 1 + (int)(rnd() % 3);
    const size_t bytes = kind == MSG_TYPE_DEPTH_DATA ? (size_t)w * h * 2 : (size_t)w * h * 3;
    size_t len;
    if (kind == MSG_TYPE_METADATA) len = youth_chunk_build(msg, MSG_TYPE_METADATA, it, (uint32_t)it, w, h, NULL, 0, 0);
    else len = youth_chunk_build(msg, kind, it, (uint32_t)it, w, h, kind == MSG_TYPE_DEPTH_DATA ? (void*)depth : (void*)color, bytes,
                                 (int)(rnd() % (uint32_t)youth_chunk_count(bytes)));
    if (len == 0 || len > MAX_MSG_SIZE) return fprintf(stderr, "chunk_build length %zu\n", len), 1;
    const int mut = (int)(rnd() % 4);
    for (int k = 0; k < mut; ++k) msg[rnd() % sizeof(MessageHeader)] = (unsigned char)rnd(); /* header fields only */
    if (rnd() % 16 == 0) len = rnd() % (len + 1); /* truncated */
    const int rc = youth_reasm_feed(r, msg, len);
    if (rc == 1) {
      int ww, hh, id;
      uint32_t ts;
      youth_reasm_info(r, &ww, &hh, &id, &ts);
      volatile uint16_t a = youth_reasm_depth(r)[(size_t)ww * hh - 1]; /* the buffers must cover the announced size */
      volatile uint8_t b = youth_reasm_color(r)[(size_t)ww * hh * 3 - 1];
      (void)a, (void)b;
      ++frames;
    } else if (rc < 0) ++rejected;
    YouthPoseMsg pm;
    int id;
    uint32_t ts;
    (void)youth_pose_msg_parse(msg, len, &id, &ts, &pm);
  }
  youth_reasm_destroy(r);
  /* 2. recordings: a valid file with random bytes flipped in the headers / truncated at random */
  snprintf(path, sizeof(path), "%s/fuzz.bin", argv[1]);
  uint16_t* dbuf = (uint16_t*)malloc(64 * 48 * 2);
  uint8_t* cbuf = (uint8_t*)malloc(64 * 48 * 3);
  long records = 0;
  for (int it = 0; it < 300; ++it) {
    FILE* f = fopen(path, "wb");
    if (!f) return fprintf(stderr, "cannot write %s\n", path), 1;
    for (int i = 0; i < 4; ++i) youth_bin_write_frame(f, (uint32_t)i, (uint32_t)i, 64, 48, depth, color);
    youth_bin_write_eof(f);
    long size = ftell(f);
    fclose(f);
    f = fopen(path, "r+b");
    for (int k = 0; k < 3; ++k) {
      const long rec = (long)(rnd() % 4) * (28 + 64 * 48 * 5);
      fseek(f, rec + (long)(rnd() % 28), SEEK_SET);
      fputc((int)(rnd() & 0xff), f);
    }
    fclose(f);
    if (rnd() % 3 == 0 && truncate(path, (off_t)(rnd() % (uint32_t)size)) != 0) return fprintf(stderr, "truncate failed\n"), 1;
    f = fopen(path, "rb");
    FrameHeader hdr;
    while (youth_bin_read_frame(f, &hdr, dbuf, 64 * 48 * 2, cbuf, 64 * 48 * 3)) {
      if (hdr.depthDataSize > 64 * 48 * 2 || hdr.colorDataSize > 64 * 48 * 3) return fprintf(stderr, "payload caps ignored\n"), 1;
      ++records;
    }
    fclose(f);
  }
  /* 3. camera YAML: random keys, numbers out of range, junk */
  snprintf(path, sizeof(path), "%s/fuzz.yaml", argv[1]);
  static const char* keys[] = {"Camera.fx", "Camera.fy", "Camera.cx", "Camera.cy", "Camera.width", "Camera.height",
                               "DepthMapFactor", "Camera.k1", "junk"};
  static const char* vals[] = {"570.3", "-1", "0", "1e40", "-1e40", "nan", "inf", "abc", "", "99999999999999999999", "640", "3.5e-320"};
  long parsed = 0;
  for (int it = 0; it < 2000; ++it) {
    FILE* f = fopen(path, "w");
    for (int l = 0; l < 8; ++l) {
      if (rnd() % 8 == 0) {
        for (int k = 0; k < 40; ++k) fputc((int)(rnd() % 255) + 1, f);
        fputc('\n', f);
      } else {
        fprintf(f, "%s%s %s\n", keys[rnd() % 9], rnd() % 6 ? ":" : "", vals[rnd() % 12]);
      }
    }
    fclose(f);
    youth_cuda_config cfg;
    if (youth_config_from_yaml(path, &cfg)) ++parsed;
    if (!(cfg.depth_factor > 0.0f)) return fprintf(stderr, "depth factor %g\n", (double)cfg.depth_factor), 1;
  }
  free(msg);
  free(dbuf);
  free(cbuf);
  printf("ok %ld frames %ld rejected %ld records %ld configs\n", frames, rejected, records, parsed);
  return 0;
}
