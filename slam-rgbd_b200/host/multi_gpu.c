/*
 * multi_gpu.c -- native multi-GPU driver: one host pthread and one youth_cuda_handle per GPU, all in one process,
 * the way the reference runs its modules (one pthread each: LoggingModule/loggingModule.c:659-662, main.c:279-281).
 *
 *   youth_multi <gpus> <sequences_per_gpu> <frames> <out_prefix> [steps] [width height levels]
 *
 * The path shards over independent sequences only: GPU g tracks sequences
 * g*S .. g*S+S-1 (seed 20261018 + sequence) from page-locked host frames, the S sequences of a GPU in the same
 * launches.  No collective inside the data path; ONE ncclAllGather at the end of the run hands every GPU all
 * trajectories (G x S x frames x 12 floats, zero-copy from youth_cuda_trajectory_device_ptr, ordered behind the last
 * step on the handle's own stream).  Thread 0 then writes one TUM file
 * per sequence, <out_prefix>_seq<NNN>_trajectory.txt, from ITS copy of the gathered buffer -- so the files of the
 * sequences tracked on other GPUs prove the gather.  Timing: every thread meets a barrier, tracks `steps` steps,
 * synchronises its device and meets a barrier again; the wall clock between the barriers is the max over GPUs.
 * Prints one JSON line.  Same numbers of sequences on fewer GPUs give bit-identical files (tests/test_multi_gpu.py).
 */
#define _GNU_SOURCE
#include <nccl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "youth_host.h"

typedef struct {
  int g, G, S, n, steps, w, h, levels;
  ncclComm_t comm;
  pthread_barrier_t* bar;
  const char* prefix;
  int ok;
  double seconds; /* thread 0: wall clock of the timed steps */
  char err[256];
} Rank;

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void* rank_main(void* arg) {
  Rank* r = (Rank*)arg;
  r->ok = 0;
  youth_cuda_handle* h = NULL;
  uint16_t* frames = NULL;
  float *gathered = NULL, *host = NULL;
  const size_t fpx = (size_t)r->w * r->h, traj_floats = (size_t)r->S * r->n * 12;
  int failed = 0;
  if (!youth_cuda_set_device(r->g)) failed = 1;
  youth_cuda_config cfg;
  youth_cuda_default_config(&cfg);
  if (r->w != 640 || r->h != 480) {
    cfg.width = r->w;
    cfg.height = r->h;
    cfg.fx = cfg.fy = 570.3f * (float)r->w / 640.0f;
    cfg.cx = (float)r->w / 2.0f;
    cfg.cy = (float)r->h / 2.0f;
  }
  cfg.levels = r->levels;
  const int iters[4] = {10, 5, 4, 4};
  for (int l = 0; l < 4; ++l) cfg.iters[l] = l < r->levels ? iters[l] : 0;
  cfg.n_streams = r->S;
  cfg.batch = r->n;
  cfg.traj_capacity = r->n; /* the S trajectories are contiguous: [S][n][12] */
  cfg.icp_ppt = (r->n * r->S >= 64) ? 128 : cfg.icp_ppt;
  cfg.device = r->g;
  if (!failed && !youth_cuda_init(&cfg, &h)) failed = 1;
  if (!failed) {
    frames = (uint16_t*)youth_cuda_host_alloc(fpx * 2 * (size_t)r->S * r->n);
    gathered = (float*)youth_cuda_device_alloc(sizeof(float) * traj_floats * (size_t)r->G);
    if (!frames || !gathered) failed = 1;
  }
  const uint16_t* ptrs[64];
  /* The trajectories are gathered ONCE, behind the last step (one collective at the end of a
   * run; a "step" re-tracks the same sequences for timing).  YOUTH_MULTI_GATHER_EVERY_STEP=1 gathers behind every
   * step instead; YOUTH_MULTI_DEVICE_INPUT=1 tracks from a device-resident copy of the frames (no H2D in the loop). */
  const int no_gather = getenv("YOUTH_MULTI_GATHER_EVERY_STEP") == NULL, dev_input = getenv("YOUTH_MULTI_DEVICE_INPUT") != NULL;
  uint16_t* dframes = NULL;
  if (!failed) {
    for (int k = 0; k < r->S; ++k) {
      youth_synth_config sc;
      youth_synth_default(&sc, r->w, r->h, r->g * r->S + k);
      youth_synth_sequence(&sc, 0, r->n, frames + fpx * (size_t)r->n * k);
      ptrs[k] = frames + fpx * (size_t)r->n * k;
    }
    if (dev_input) {
      dframes = (uint16_t*)youth_cuda_device_alloc(fpx * 2 * (size_t)r->S * r->n);
      if (!dframes || !youth_cuda_copy_to_device(dframes, frames, fpx * 2 * (size_t)r->S * r->n)) failed = 1;
      for (int k = 0; k < r->S; ++k) ptrs[k] = dframes + fpx * (size_t)r->n * k;
    }
  }
  /* every rank reaches the barriers and the collective whether or not it failed locally: nobody is left waiting */
  pthread_barrier_wait(r->bar);
  const double t0 = now_s();
  for (int s = 0; s < r->steps; ++s) {
    /* no host synchronisation inside the loop: the tracker call returns after enqueueing (its H2D copies run on the
     * handle's copy stream under the kernels of the step before), and the one collective -- the per-sequence
     * trajectories of every GPU -- is ordered behind the step on the handle's own stream */
    if (!failed) {
      if (!youth_cuda_reset(h, -1) ||
          !youth_cuda_track_batch(h, ptrs, r->n, dev_input ? YOUTH_MEM_DEVICE : YOUTH_MEM_HOST_PINNED, NULL, NULL))
        failed = 1;
    }
    if (!no_gather || s == r->steps - 1) {
      const void* send = (failed || !h) ? (const void*)gathered : youth_cuda_trajectory_device_ptr(h, 0);
      if (ncclAllGather(send, gathered, traj_floats, ncclFloat, r->comm, h ? (cudaStream_t)youth_cuda_stream(h) : NULL) != ncclSuccess) failed = 1;
    }
  }
  if (h && !youth_cuda_sync(h)) failed = 1;
  if (!youth_cuda_device_sync()) failed = 1;
  pthread_barrier_wait(r->bar);
  r->seconds = now_s() - t0;
  if (failed) snprintf(r->err, sizeof(r->err), "%s", youth_cuda_last_error());
  if (!failed && r->g == 0) {
    host = (float*)malloc(sizeof(float) * traj_floats * (size_t)r->G);
    uint32_t* ts = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)r->n);
    if (!host || !ts || !youth_cuda_copy_to_host(host, gathered, sizeof(float) * traj_floats * (size_t)r->G)) failed = 1;
    for (int i = 0; i < r->n && !failed; ++i) ts[i] = (uint32_t)(33 * i);
    for (int q = 0; q < r->G * r->S && !failed; ++q) {
      char path[1024];
      snprintf(path, sizeof(path), "%s_seq%03d_trajectory.txt", r->prefix, q);
      if (!youth_tum_write(path, host + (size_t)q * r->n * 12, ts, r->n)) failed = 1;
    }
    free(ts);
  }
  free(host);
  youth_cuda_device_free(gathered);
  youth_cuda_device_free(dframes);
  youth_cuda_host_free(frames);
  youth_cuda_destroy(h);
  r->ok = !failed;
  return NULL;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <gpus> <sequences_per_gpu> <frames> <out_prefix> [steps] [width height levels]\n", argv[0]);
    return 2;
  }
  const int G = atoi(argv[1]), S = atoi(argv[2]), n = atoi(argv[3]);
  const int steps = argc > 5 ? atoi(argv[5]) : 1;
  const int w = argc > 8 ? atoi(argv[6]) : 640, h = argc > 8 ? atoi(argv[7]) : 480, levels = argc > 8 ? atoi(argv[8]) : 3;
  if (G < 1 || G > 16 || S < 1 || S > 64 || n < 1 || steps < 1 || levels < 1 || levels > 4) return 2;
  if (youth_cuda_device_count() < G) {
    fprintf(stderr, "youth_multi: %d GPUs wanted, %d visible (no CPU fallback)\n", G, youth_cuda_device_count());
    return 1;
  }
  /* threads of one process: every device launches the collective on its own, with no cross-device launch group */
  setenv("NCCL_LAUNCH_MODE", "PARALLEL", 0);
  ncclComm_t comms[16];
  int devs[16];
  for (int g = 0; g < G; ++g) devs[g] = g;
  if (ncclCommInitAll(comms, G, devs) != ncclSuccess) {
    fprintf(stderr, "youth_multi: ncclCommInitAll failed\n");
    return 1;
  }
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, (unsigned)G);
  Rank ranks[16];
  pthread_t th[16];
  memset(ranks, 0, sizeof(ranks));
  for (int g = 0; g < G; ++g) {
    Rank* r = &ranks[g];
    r->g = g; r->G = G; r->S = S; r->n = n; r->steps = steps; r->w = w; r->h = h; r->levels = levels;
    r->comm = comms[g];
    r->bar = &bar;
    r->prefix = argv[4];
    pthread_create(&th[g], NULL, rank_main, r);
  }
  int ok = 1;
  for (int g = 0; g < G; ++g) {
    pthread_join(th[g], NULL);
    if (!ranks[g].ok) {
      fprintf(stderr, "youth_multi: GPU %d failed: %s\n", g, ranks[g].err);
      ok = 0;
    }
  }
  for (int g = 0; g < G; ++g) ncclCommDestroy(comms[g]);
  pthread_barrier_destroy(&bar);
  if (!ok) return 1;
  const double secs = ranks[0].seconds;
  printf("{\"gpus\": %d, \"sequences_per_gpu\": %d, \"frames_per_sequence\": %d, \"steps\": %d, \"seconds\": %.6f, "
         "\"frames_per_sec\": %.1f, \"what\": \"one pthread + one tracker handle per GPU, pinned host frames in, no host synchronisation between steps, one "
         "stream-ordered ncclAllGather of the trajectories behind the last step (H2D and the gather inside the timed region)\", \"trajectories\": \"%s_seq<NNN>_trajectory.txt\"}\n",
         G, S, n, steps, secs, (double)G * S * n * steps / secs, argv[4]);
  return 0;
}
