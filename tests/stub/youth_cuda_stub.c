/*
 * youth_cuda_stub.c -- TEST DOUBLE of the inner C ABI (include/youth_cuda.h), test infrastructure only.
 *
 * Linked with the facade sources into tests/_build/libfacade_under_test.so so that the HOST LOGIC of
 * slam-rgbd_b200/host/slam_facade.c and algorithm_module.c -- the frame ring, the reference's drop-oldest
 * back-pressure (SLAM.cpp:163-167), the lossless mode, runs in flight, drain / stop / reset, TUM egress --
 * can be exercised on a machine without a GPU.  It is NOT a CPU implementation of the tracker and not a
 * fallback: it performs no tracking arithmetic at all.  A "pose" is a marker of the frame it was handed
 * (identity rotation; t = (first pixel of the frame, index in the call, frames per call)), which is exactly
 * what the tests need to see which frames reached the tracker, in which order and in which groups.
 * The product never links or loads this file: libAlgorithmModule.so is linked against libyouth_cuda.so
 * (slam-rgbd_b200/Makefile), whose entry points fail loudly without a CUDA device.
 *
 * YOUTH_STUB_DELAY_US: microseconds a track_batch call takes per frame (a slow tracker for back-pressure tests).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "youth_codec.h"
#include "youth_cuda.h"
#include "youth_model.h"

struct youth_cuda_handle {
  youth_cuda_config cfg;
  int count;
  float* poses;  /* [traj_capacity][12] */
  uint32_t* ts;  /* [traj_capacity] */
  uint64_t ticket;
  int delay_us;
  int calls;
};

static youth_cuda_handle* g_last; /* for stub_* inspection from the tests */
static long g_torn; /* frames whose first / last pixel changed while track_batch held them */

static const char* g_err = "";
const char* youth_cuda_last_error(void) { return g_err; }
static int fail(const char* why) {
  g_err = why;
  return 0;
}
int youth_cuda_abi_version(void) { return YOUTH_CUDA_ABI_VERSION; }

int youth_cuda_default_config(youth_cuda_config* c) {
  if (!c) return 0;
  memset(c, 0, sizeof(*c));
  c->width = 640;
  c->height = 480;
  c->fx = c->fy = 570.3f;
  c->cx = 320.0f;
  c->cy = 240.0f;
  c->depth_factor = 1000.0f;
  c->levels = 3;
  c->iters[0] = 10;
  c->iters[1] = 5;
  c->iters[2] = c->iters[3] = 4;
  c->depth_min_mm = 1;
  c->depth_max_mm = 10000;
  c->bilateral = 1;
  c->sigma_space_px = 4.5f;
  c->sigma_range_mm = 30.0f;
  c->dist_thresh_m = 0.10f;
  c->cos_thresh = 0.93969262f;
  c->min_inliers = 100;
  c->icp_ppt = 64;
  c->n_streams = 1;
  c->batch = 8;
  c->traj_capacity = 4096;
  return 1;
}

int youth_cuda_init(const youth_cuda_config* cfg, youth_cuda_handle** out) {
  if (!cfg || !out || cfg->width <= 0 || cfg->height <= 0 || cfg->batch < 1 || cfg->traj_capacity < 1) return 0;
  youth_cuda_handle* h = (youth_cuda_handle*)calloc(1, sizeof(*h));
  h->cfg = *cfg;
  h->poses = (float*)calloc((size_t)cfg->traj_capacity * 12, sizeof(float));
  h->ts = (uint32_t*)calloc((size_t)cfg->traj_capacity, sizeof(uint32_t));
  const char* d = getenv("YOUTH_STUB_DELAY_US");
  h->delay_us = d ? atoi(d) : 0;
  g_last = h;
  *out = h;
  return 1;
}

void youth_cuda_destroy(youth_cuda_handle* h) {
  if (!h) return;
  if (g_last == h) g_last = NULL;
  free(h->poses);
  free(h->ts);
  free(h);
}

int youth_cuda_track_batch(youth_cuda_handle* h, const uint16_t* const* depth, int n_frames, int mem_kind,
                           const uint32_t* timestamps_ms, float* poses_out) {
  (void)mem_kind;
  if (!h || !depth || !depth[0] || n_frames < 1 || n_frames > h->cfg.batch) return fail("stub: bad track_batch arguments");
  if (h->count + n_frames > h->cfg.traj_capacity) return fail("stub: trajectory capacity exceeded");
  const size_t npx = (size_t)h->cfg.width * h->cfg.height;
  for (int i = 0; i < n_frames; ++i) {
    float* p = h->poses + 12 * (size_t)(h->count + i);
    memset(p, 0, 12 * sizeof(float));
    p[0] = p[5] = p[10] = 1.0f;
    p[3] = (float)depth[0][npx * (size_t)i]; /* marker the test wrote into the first pixel */
    p[7] = (float)i;
    p[11] = (float)n_frames;
    h->ts[h->count + i] = timestamps_ms ? timestamps_ms[i] : 0u;
  }
  if (h->delay_us > 0) usleep((useconds_t)h->delay_us * (useconds_t)n_frames);
  /* the frames belong to the tracker until the call returns (the H2D copy of the real library reads them
   * asynchronously): nobody may have written to them meanwhile */
  for (int i = 0; i < n_frames; ++i) {
    const float* p = h->poses + 12 * (size_t)(h->count + i);
    if ((float)depth[0][npx * (size_t)i] != p[3] || depth[0][npx * (size_t)i + npx - 1] != depth[0][npx * (size_t)i]) ++g_torn;
  }
  if (poses_out) memcpy(poses_out, h->poses + 12 * (size_t)h->count, sizeof(float) * 12 * (size_t)n_frames);
  h->count += n_frames;
  h->calls++;
  return 1;
}

int youth_cuda_track_batch_packed(youth_cuda_handle* h, const uint8_t* const* streams, const uint64_t* const* offsets,
                                  int n_frames, int mem_kind, const uint32_t* timestamps_ms, float* poses_out) {
  (void)h, (void)streams, (void)offsets, (void)n_frames, (void)mem_kind, (void)timestamps_ms, (void)poses_out;
  return 0; /* packed records need the device codec */
}

int youth_cuda_frame_count(youth_cuda_handle* h, int stream) { return h && stream == 0 ? h->count : -1; }

int youth_cuda_get_trajectory(youth_cuda_handle* h, int stream, int first, int max_frames, float* poses_out,
                              uint32_t* timestamps_out, uint32_t* status_out) {
  if (!h || stream != 0 || first < 0 || max_frames < 0) return -1;
  int n = h->count - first;
  if (n < 0) n = 0;
  if (n > max_frames) n = max_frames;
  if (poses_out) memcpy(poses_out, h->poses + 12 * (size_t)first, sizeof(float) * 12 * (size_t)n);
  if (timestamps_out) memcpy(timestamps_out, h->ts + first, sizeof(uint32_t) * (size_t)n);
  if (status_out)
    for (int i = 0; i < n; ++i) status_out[i] = first + i == 0 ? YOUTH_STATUS_FIRST : 0u;
  return n;
}

int youth_cuda_read_trajectory_async(youth_cuda_handle* h, int stream, int first, int max_frames, float* poses_out,
                                     uint32_t* status_out, uint64_t* ticket_out) {
  const int n = youth_cuda_get_trajectory(h, stream, first, max_frames, poses_out, NULL, status_out);
  if (n >= 0 && ticket_out) *ticket_out = ++h->ticket;
  return n;
}

int youth_cuda_wait_ticket(youth_cuda_handle* h, uint64_t ticket) { return h && ticket <= h->ticket; }

int youth_cuda_read_last_inliers_async(youth_cuda_handle* h, int stream, int* inliers_out) {
  if (!h || stream != 0 || !inliers_out) return 0;
  *inliers_out = 1000 + h->count;
  return 1;
}

int youth_cuda_last_inliers(youth_cuda_handle* h, int stream) { return h && stream == 0 ? 1000 + h->count : 0; }

int youth_cuda_sync(youth_cuda_handle* h) { return h != NULL; }

int youth_cuda_reset(youth_cuda_handle* h, int stream) {
  (void)stream;
  if (!h) return 0;
  h->count = 0;
  return 1;
}

void* youth_cuda_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void youth_cuda_host_free(void* p) { free(p); }

int youth_tsdf_default_config(youth_tsdf_config* cfg) {
  if (cfg) memset(cfg, 0, sizeof(*cfg));
  return 0;
}
int youth_cuda_enable_model(youth_cuda_handle* h, const youth_tsdf_config* cfg) {
  (void)h, (void)cfg;
  return 0;
}
int youth_cuda_model_enabled(const youth_cuda_handle* h) {
  (void)h;
  return 0;
}
long long youth_cuda_model_surface_voxels(youth_cuda_handle* h, int stream) {
  (void)h, (void)stream;
  return -1;
}

/* inspection for the tests */
int stub_track_calls(void) { return g_last ? g_last->calls : -1; }
long stub_torn_frames(void) { return g_torn; }
