/*
 * facade_host_bench.c -- host-side ceiling of the SLAM.h facade: processSlamFrame() per 640x480 frame with the
 * test double of the inner C ABI behind it (tests/stub/youth_cuda_stub.c: a tracker that takes no time), i.e.
 * what the caller's thread, the copy-in and the queue cost on their own.  No GPU needed.
 *   gcc -O2 -std=gnu11 -Iinclude -o /tmp/facade_host_bench tools/facade_host_bench.c tests/stub/youth_cuda_stub.c \
 *       slam-rgbd_b200/host/slam_facade.c slam-rgbd_b200/host/youth_frameio.c slam-rgbd_b200/host/youth_config.c \
 *       slam-rgbd_b200/host/youth_synth.c -lpthread -lrt -lm && YOUTH_SLAM_TRAJ_CAPACITY=20000 /tmp/facade_host_bench
 * This container (8-core Xeon, one producer thread): 9.7 k frames/s with memcpy copy-in and a signal per frame,
 * 17.8 k frames/s with non-temporal stores and a signal only on an empty queue (56 us per frame, 47 of them the copy).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "SLAM.h"
void youthSlamSetOptions(int lossless, int batch);
void youthSlamDrain(void);
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){
  int n=argc>1?atoi(argv[1]):3000;
  youthSlamSetOptions(1,64);
  initSlamModule(NULL,NULL);
  if(!isSlamModuleRunning())return 1;
  int16_t* f=malloc(640*480*2*8); memset(f,1,640*480*2*8);
  for(int rep=0;rep<3;++rep){
    double t0=now();
    for(int i=0;i<n;++i) processSlamFrame(f+(size_t)(i&7)*640*480,NULL,640,480,i);
    youthSlamDrain();
    double t1=now();
    printf("%d frames in %.3f s = %.0f frames/s (%.1f us/frame, %.1f GB/s copy-in)\n",n,t1-t0,n/(t1-t0),1e6*(t1-t0)/n, n*614400.0/(t1-t0)/1e9);
    resetSlam();
  }
  stopSlamModule(); return 0;}
