/*
 * facade_race_driver.c -- TEST INFRASTRUCTURE: drives the facade (slam_facade.c) from several threads so that
 * ThreadSanitizer / AddressSanitizer builds (tests/test_facade_host_logic.py) can look at its locking.  The
 * tracker behind the facade is the test double (youth_cuda_stub.c).  Scenarios, each in both queue modes:
 *   a producer pushing frames while the main thread stops the module;
 *   a producer + a status reader (isSlamModuleRunning / getSlamMapPoints / youthSlamStats) + drain, reset,
 *   saveSlamMap from the main thread, then stop;
 *   the same with a SECOND producer that uses the zero-copy calls (youthSlamAcquireSlot / youthSlamCommitSlot)
 *   at the same time as the first one calls processSlamFrame.
 * usage: facade_race_driver <config.yaml> <out_prefix>
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "SLAM.h"
#include "youth_slam_ext.h"

long stub_torn_frames(void);

enum { W = 64, H = 48 };
static atomic_int g_quit;
static int g_pace_us; /* > 0: the producer is slower than the (stub) tracker, so that drain() meets an idle worker */

static void* producer(void* arg) {
  long* pushed = (long*)arg;
  int16_t* f = (int16_t*)calloc((size_t)W * H, sizeof(int16_t));
  for (uint32_t i = 0; !atomic_load(&g_quit); ++i) {
    f[0] = f[W * H - 1] = (int16_t)(i & 0x7fff);
    if (processSlamFrame(f, NULL, W, H, i)) ++*pushed;
    else if (!isSlamModuleRunning()) break;
    if (g_pace_us) usleep((useconds_t)g_pace_us);
  }
  free(f);
  return NULL;
}

/* fills the ring slot in place: the frame is born in the page-locked ring */
static void* producer_zero_copy(void* arg) {
  long* pushed = (long*)arg;
  for (uint32_t i = 0; !atomic_load(&g_quit); ++i) {
    uint16_t* slot = youthSlamAcquireSlot(W, H);
    if (slot) {
      memset(slot, 0, (size_t)W * H * sizeof(uint16_t));
      slot[0] = slot[W * H - 1] = (uint16_t)(i & 0x7fff);
      if (youthSlamCommitSlot(i)) ++*pushed;
    } else if (!isSlamModuleRunning()) {
      break;
    }
    if (g_pace_us) usleep((useconds_t)g_pace_us);
  }
  return NULL;
}

static void* reader(void* arg) {
  (void)arg;
  long a, d, t, sink = 0;
  while (!atomic_load(&g_quit)) {
    sink += isSlamModuleRunning() + getSlamMapPoints();
    youthSlamStats(&a, &d, &t);
    sink += a + d + t;
    usleep(50);
  }
  return (void*)(sink & 1);
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  setenv("YOUTH_STUB_DELAY_US", "100", 1);
  setenv("YOUTH_SLAM_COPY_MIN_BYTES", "1024", 1); /* the 6 KB test frames go through the parallel copy-in too */
  for (int round = 0; round < 10; ++round) {
    const int lossless = round & 1;
    youthSlamSetOptions(lossless, 4);
    initSlamModule(argv[1], NULL);
    if (!isSlamModuleRunning()) {
      fprintf(stderr, "init failed\n");
      return 1;
    }
    atomic_store(&g_quit, 0);
    g_pace_us = round >= 6 ? 400 : (round >= 2 ? 150 : 0); /* rounds 0, 1: stop under overload; two producers: each slower */
    long pushed = 0, pushed2 = 0;
    const int two = round >= 6; /* rounds 6..9: a zero-copy producer next to the copying one */
    pthread_t tp, tp2, tr;
    pthread_create(&tp, NULL, producer, &pushed);
    if (two) pthread_create(&tp2, NULL, producer_zero_copy, &pushed2);
    pthread_create(&tr, NULL, reader, NULL);
    usleep(20000);
    if (round >= 2) {
      youthSlamDrain();
      if (round >= 4 && round != 6 && round != 7) {
        resetSlam();
        usleep(5000);
        if (!saveSlamMap(argv[2])) {
          fprintf(stderr, "saveSlamMap failed\n");
          return 1;
        }
      }
    }
    stopSlamModule(); /* with the producer still pushing */
    atomic_store(&g_quit, 1);
    pthread_join(tp, NULL);
    if (two) pthread_join(tp2, NULL);
    pthread_join(tr, NULL);
    if (isSlamModuleRunning() || pushed <= 0 || (two && pushed2 <= 0)) {
      fprintf(stderr, "round %d: running=%d pushed=%ld\n", round, isSlamModuleRunning(), pushed);
      return 1;
    }
  }
  if (stub_torn_frames() != 0) {
    fprintf(stderr, "%ld frames were written while the tracker held them\n", stub_torn_frames());
    return 1;
  }
  printf("ok\n");
  return 0;
}
