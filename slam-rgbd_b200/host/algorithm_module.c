/*
 * algorithm_module.c -- pthread entry of the Algorithm stage.
 *
 * The reference body is a call to an undeclared SLAM() (Youth.Source/AlgorithmModule/
 * algorithmModule.c:3-5) and its launch site is commented out (main.c:279-281).  Here the
 * thread initialises the tracker and then either serves processSlamFrame() callers (the
 * Logging hook at loggingModule.c:354) until stopSlamModule(), or -- when `id` is a path --
 * replays a .bin recording headless, the way playbackThread does for the viewer
 * (loggingModule.c:542-594) but without the 30 fps pacing sleep, or -- when `id` is
 * "mq:<queue name>" -- takes the viewer's seat on that queue (viewerModule.c:160-250): with
 * "mq:/logger_viewer_queue" the tracker follows whatever the reference's logger passes through or its
 * playbackThread replays (loggingModule.c:284-288, :584-590), with no change to the reference at all.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "algorithmModule.h"
#include "youth_host.h"
#include "youth_slam_ext.h"

#define PACKED_RUN 64 /* packed records are gathered into runs of this many frames */

void* algorithmModule(void* id) {
  const char* replay = (const char*)id;
  if (replay) youthSlamSetOptions(1, -1); /* a recording is replayed losslessly */
  initSlamModule(getenv("YOUTH_SLAM_CONFIG"), getenv("YOUTH_SLAM_VOCAB"));
  if (!isSlamModuleRunning()) return NULL;
  if (!replay) {
    while (isSlamModuleRunning()) usleep(10000);
    return NULL;
  }
  if (strncmp(replay, "mq:", 3) == 0) {
    const char* idle = getenv("YOUTH_SLAM_MQ_IDLE_MS"); /* leave after this long without a message (default 2 s) */
    const long frames = youth_mq_consume(replay + 3, processSlamFrame, NULL, idle ? atoi(idle) : 2000);
    if (frames < 0) fprintf(stderr, "algorithmModule: cannot read queue '%s'\n", replay + 3);
    youthSlamDrain();
    const char* out = getenv("YOUTH_SLAM_OUT");
    if (out && frames > 0) saveSlamMap(out);
    fprintf(stderr, "algorithmModule: tracked %ld frames from queue %s\n", frames, replay + 3);
    stopSlamModule();
    return NULL;
  }
  FILE* f = fopen(replay, "rb");
  if (!f) {
    fprintf(stderr, "algorithmModule: cannot open '%s'\n", replay);
    stopSlamModule();
    return NULL;
  }
  const size_t cap = 64u << 20; /* large enough for 1280x960 and beyond (reference caps at 1 MiB) */
  uint16_t* depth = (uint16_t*)malloc(cap);
  FrameHeader hdr;
  long frames = 0;
  /* FRAME_TYPE_DEPTH_PACKED records (YD16, include/youth_codec.h) are gathered into runs and unpacked on the device */
  uint8_t* run = NULL;
  size_t run_cap = 0;
  uint64_t offs[PACKED_RUN + 1] = {0};
  uint32_t ts[PACKED_RUN];
  int nrun = 0, rw = 0, rh = 0, failed = 0;
  while (depth && !failed && youth_bin_read_header(f, &hdr)) {
    if (hdr.frameType != FRAME_TYPE_DEPTH_PACKED && nrun == 0 &&
        (size_t)hdr.depthDataSize == (size_t)hdr.width * hdr.height * 2) {
      /* a raw frame: read it straight into a slot of the tracker's page-locked ring (no second copy) */
      uint16_t* slot = youthSlamAcquireSlot(hdr.width, hdr.height);
      if (slot) {
        if (!youth_bin_read_payload(f, &hdr, slot, (size_t)hdr.depthDataSize, NULL, 0)) {
          youthSlamAbortSlot(); /* a torn record: nothing is published */
          break;
        }
        if (!youthSlamCommitSlot(hdr.timestamp)) break;
        ++frames;
        continue;
      }
    }
    if (!youth_bin_read_payload(f, &hdr, depth, cap, NULL, 0)) break;
    if (hdr.frameType == FRAME_TYPE_DEPTH_PACKED) {
      if (offs[nrun] + hdr.depthDataSize > run_cap) {
        run_cap = 2 * (offs[nrun] + hdr.depthDataSize) + (1u << 20);
        uint8_t* grown = (uint8_t*)realloc(run, run_cap);
        if (!grown) break;
        run = grown;
      }
      memcpy(run + offs[nrun], depth, hdr.depthDataSize);
      offs[nrun + 1] = offs[nrun] + hdr.depthDataSize;
      ts[nrun] = hdr.timestamp;
      rw = hdr.width;
      rh = hdr.height;
      if (++nrun == PACKED_RUN) {
        if (!youthSlamProcessPackedFrames(run, offs, nrun, rw, rh, ts)) failed = 1;
        else frames += nrun;
        nrun = 0;
      }
      continue;
    }
    if (nrun) { /* a raw record after packed ones: keep the order of the recording */
      if (!youthSlamProcessPackedFrames(run, offs, nrun, rw, rh, ts)) break;
      frames += nrun;
      nrun = 0;
    }
    if ((size_t)hdr.depthDataSize != (size_t)hdr.width * hdr.height * 2) {
      fprintf(stderr, "algorithmModule: record %u carries %u depth bytes for %ux%u, stopping\n", hdr.frameId,
              hdr.depthDataSize, hdr.width, hdr.height);
      break;
    }
    if (!processSlamFrame((const int16_t*)depth, NULL, hdr.width, hdr.height, hdr.timestamp)) break;
    ++frames;
  }
  if (nrun && !failed && youthSlamProcessPackedFrames(run, offs, nrun, rw, rh, ts)) frames += nrun;
  free(run);
  fclose(f);
  free(depth);
  youthSlamDrain();
  const char* out = getenv("YOUTH_SLAM_OUT");
  if (out) saveSlamMap(out);
  fprintf(stderr, "algorithmModule: replayed %ld frames from %s\n", frames, replay);
  stopSlamModule();
  return NULL;
}
