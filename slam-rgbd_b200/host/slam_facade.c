/*
 * slam_facade.c -- the AlgorithmModule C facade (include/SLAM.h) over libyouth_cuda.so.
 *
 * Replaces reference Youth.Source/AlgorithmModule/SLAM.cpp:15-228 (process-global
 * singleton, bounded frame queue, worker thread) with:
 *   - a page-locked host frame ring that processSlamFrame() copies into synchronously
 *     (ownership rule of SLAM.cpp:133-134: the caller may reuse its buffer on return);
 *   - a worker thread that hands runs of consecutive ring slots to
 *     youth_cuda_track_batch(), i.e. the device-resident frame ring is fed straight
 *     from the host ring with one async H2D per run, two runs in flight (the copy of a
 *     run overlaps the kernels of the run before);
 *   - the reference's lossy back-pressure (more than 10 queued -> drop oldest down to 5,
 *     SLAM.cpp:163-167) as the default, and a lossless mode (producer blocks) for
 *     replay / benchmarking, where dropping frames would make parity meaningless.
 * Shared flags are C11 atomics (the reference's plain bools at SLAM.cpp:17,29 are racy).
 * No C++ exceptions exist here; errors are reported by return value and on stderr.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <mqueue.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "SLAM.h"
#include "youth_slam_ext.h"
#include "youth_host.h"
#include "youth_model.h"

#define QUEUE_HIGH_WATER 10 /* SLAM.cpp:163 */
#define QUEUE_LOW_WATER 5   /* SLAM.cpp:165 */
#define MAX_GROUP 512       /* frames per launch group (youthSlamSetOptions / YOUTH_SLAM_BATCH) */

/* The host frame ring.  Positions are absolute frame counters (slot = position % qcap):
 *     done <= head <= tail,   tail - done < qcap
 *   [done, head)  handed to the tracker and not collected yet (or dropped while such a run was in flight)
 *   [head, tail)  published, waiting for the worker
 *   tail          the slot a producer is filling right now (`filling`), published by the commit
 * A producer reserves the slot under the mutex, fills it WITHOUT the mutex (614 KB per VGA frame: the worker
 * claims and collects runs meanwhile) and publishes it under the mutex.  Space is accounted from `done`, so the
 * write position cannot run round into frames the tracker still reads. */
static struct {
  pthread_mutex_t mu;
  pthread_cond_t nonempty, nonfull, idle, prod;
  atomic_int running;
  atomic_int stop_req;
  atomic_int failed; /* sticky: the tracker refused a run (e.g. trajectory capacity exhausted); cleared by resetSlam */
  youth_cuda_handle* h;
  youth_cuda_config cfg;
  uint16_t* ring; /* pinned, qcap frames */
  uint32_t* ts;
  int qcap;
  uint64_t done, head, tail;
  int inflight; /* frames in runs handed to the tracker and not collected */
  int filling;  /* a producer holds the slot at `tail` */
  int lossless, batch;
  pthread_t worker;
  long accepted, dropped, tracked, poses_sent;
  atomic_int last_inliers;
  atomic_int model_on; /* frame-to-model tracking (YOUTH_SLAM_MODE=model) */
  mqd_t pose_mq; /* optional pose egress to the viewer queue (MSG_TYPE_POSE), -1 = off */
} G = {.mu = PTHREAD_MUTEX_INITIALIZER,
       .nonempty = PTHREAD_COND_INITIALIZER,
       .nonfull = PTHREAD_COND_INITIALIZER,
       .idle = PTHREAD_COND_INITIALIZER,
       .prod = PTHREAD_COND_INITIALIZER,
       .lossless = -1,
       .batch = -1,
       .pose_mq = (mqd_t)-1};

static size_t frame_px(void) { return (size_t)G.cfg.width * (size_t)G.cfg.height; }
static int waiting(void) { return (int)(G.tail - G.head); }

/* The synchronous copy-in of processSlamFrame (SLAM.cpp:133-134) is what bounds the facade's frame rate: one
 * caller thread moving 614 KB per VGA frame into a ring far larger than the caches.  The ring is read next by
 * the GPU's copy engine, not by this CPU, so the frame is written with non-temporal stores: no read-for-ownership
 * of the destination lines (measured on a Xeon host: 13.0 instead of 6.5 GB/s, 47 instead of 94 us per frame).
 * The fence orders the streaming stores before the mutex release that publishes the slot. */
#if defined(__SSE2__)
#include <emmintrin.h>
static void copy_in(void* dst, const void* src, size_t bytes) {
  char* d = (char*)dst;
  const char* s = (const char*)src;
  size_t i = 0;
  if (((uintptr_t)d & 15u) == 0) {
    for (; i + 64 <= bytes; i += 64) {
      const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 16));
      const __m128i c = _mm_loadu_si128((const __m128i*)(s + i + 32)), e = _mm_loadu_si128((const __m128i*)(s + i + 48));
      _mm_stream_si128((__m128i*)(d + i), a);
      _mm_stream_si128((__m128i*)(d + i + 16), b);
      _mm_stream_si128((__m128i*)(d + i + 32), c);
      _mm_stream_si128((__m128i*)(d + i + 48), e);
    }
    _mm_sfence();
  }
  if (i < bytes) memcpy(d + i, s + i, bytes - i);
}
#else
static void copy_in(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
#endif

/* Parallel copy-in.  One thread moves a VGA frame in ~50-60 us (13 GB/s of streaming stores), i.e. at most ~18 k
 * frames/s through processSlamFrame() however fast the tracker is.  For replay and benchmarks (frames arriving back
 * to back) a few helper threads share the copy: the caller's thread copies the first stripe and waits for the
 * others, so the call is still synchronous (SLAM.cpp:133-134: the caller may reuse its buffer on return).  Helpers
 * spin briefly for the next frame and then sleep on a condition variable, so a live 30 fps producer costs nothing
 * between frames.  YOUTH_SLAM_COPY_THREADS = total threads per frame copy (default 4, 1 = off). */
#define COPY_HELPERS_MAX 7
#define COPY_MIN_BYTES (256u << 10) /* smaller frames are copied by the caller alone (YOUTH_SLAM_COPY_MIN_BYTES overrides) */
static struct {
  pthread_t th[COPY_HELPERS_MAX];
  int n; /* helpers running */
  pthread_mutex_t mu;
  pthread_cond_t go;
  atomic_ullong gen; /* job counter: a new value = a new job is posted */
  atomic_int remaining;
  atomic_int quit;
  atomic_int sleepers;
  const char* src;
  char* dst;
  size_t stripe, bytes, min_bytes;
  unsigned long long gen0; /* value of gen when the helpers of this run of the module were started */
} CP = {.mu = PTHREAD_MUTEX_INITIALIZER, .go = PTHREAD_COND_INITIALIZER};

static void* copy_helper(void* arg) {
  const int k = (int)(intptr_t)arg; /* stripe index 1 .. n */
  unsigned long long seen = CP.gen0; /* jobs of an earlier run of the module are not ours */
  for (;;) {
    unsigned long long g = atomic_load_explicit(&CP.gen, memory_order_acquire);
    int spins = 0;
    while (g == seen && !atomic_load(&CP.quit)) {
      if (++spins < 4000) {
#if defined(__SSE2__)
        _mm_pause();
#endif
      } else { /* ~50 us without a frame: sleep until the next one is posted */
        pthread_mutex_lock(&CP.mu);
        atomic_fetch_add(&CP.sleepers, 1);
        /* sleepers++ then gen (both seq_cst) against the poster's gen++ then sleepers: one of the two sees the other */
        while (atomic_load(&CP.gen) == seen && !atomic_load(&CP.quit)) pthread_cond_wait(&CP.go, &CP.mu);
        atomic_fetch_sub(&CP.sleepers, 1);
        pthread_mutex_unlock(&CP.mu);
        spins = 0;
      }
      g = atomic_load_explicit(&CP.gen, memory_order_acquire);
    }
    if (atomic_load(&CP.quit)) return NULL;
    seen = g;
    const size_t a = CP.stripe * (size_t)k;
    if (a < CP.bytes) copy_in(CP.dst + a, CP.src + a, CP.bytes - a < CP.stripe ? CP.bytes - a : CP.stripe);
    atomic_fetch_sub_explicit(&CP.remaining, 1, memory_order_release);
  }
}

static void copy_pool_start(void) {
  const char* e = getenv("YOUTH_SLAM_COPY_THREADS");
  int want = e ? atoi(e) : 4;
  if (want > COPY_HELPERS_MAX + 1) want = COPY_HELPERS_MAX + 1;
  const char* mb = getenv("YOUTH_SLAM_COPY_MIN_BYTES");
  CP.min_bytes = mb ? (size_t)atol(mb) : COPY_MIN_BYTES;
  if (CP.min_bytes < 128) CP.min_bytes = 128;
  atomic_store(&CP.quit, 0);
  atomic_store(&CP.remaining, 0);
  CP.gen0 = atomic_load(&CP.gen);
  CP.n = 0;
  for (int k = 1; k < want; ++k) {
    if (pthread_create(&CP.th[CP.n], NULL, copy_helper, (void*)(intptr_t)k) != 0) break;
    CP.n++;
  }
}

static void copy_pool_stop(void) {
  pthread_mutex_lock(&CP.mu);
  atomic_store(&CP.quit, 1);
  pthread_cond_broadcast(&CP.go);
  pthread_mutex_unlock(&CP.mu);
  for (int k = 0; k < CP.n; ++k) pthread_join(CP.th[k], NULL);
  CP.n = 0;
}

/* one producer at a time gets here (the `filling` flag of the ring) */
static void copy_in_parallel(void* dst, const void* src, size_t bytes) {
  if (CP.n == 0 || bytes < CP.min_bytes) {
    copy_in(dst, src, bytes);
    return;
  }
  const size_t parts = (size_t)CP.n + 1;
  CP.stripe = ((bytes + parts - 1) / parts + 63) & ~(size_t)63; /* whole cache lines, 16-byte aligned stripes */
  CP.bytes = bytes;
  CP.src = (const char*)src;
  CP.dst = (char*)dst;
  atomic_store_explicit(&CP.remaining, CP.n, memory_order_relaxed);
  atomic_fetch_add(&CP.gen, 1); /* seq_cst: publishes the job, and orders against the sleepers test below */
  if (atomic_load(&CP.sleepers) > 0) {
    pthread_mutex_lock(&CP.mu);
    pthread_cond_broadcast(&CP.go);
    pthread_mutex_unlock(&CP.mu);
  }
  copy_in(dst, src, CP.stripe < bytes ? CP.stripe : bytes);
  while (atomic_load_explicit(&CP.remaining, memory_order_acquire) > 0) {
#if defined(__SSE2__)
    _mm_pause();
#endif
  }
}

/* A run of ring slots handed to the tracker and not collected yet.  The worker keeps up to two in
 * flight: it submits run g+1 (asynchronous youth_cuda_track_batch: its H2D copy runs on the copy stream
 * under the kernels of run g) before it waits for run g, publishes its poses and releases its slots.
 * With an empty queue it collects at once, so a lone live frame is not delayed. */
typedef struct {
  int valid, ok, n, first, base, buf;
  uint64_t pos; /* ring position of the run's first frame */
  uint64_t ticket;
} PendingRun;

/* next: the run submitted after r (still in flight), or NULL */
static void collect_run(const PendingRun* r, const PendingRun* next, float* const poses[2], uint32_t* const status[2],
                        int* const inl[2]) {
  long sent = 0;
  int ok = r->ok;
  if (ok && !youth_cuda_wait_ticket(G.h, r->ticket)) {
    fprintf(stderr, "AlgorithmModule: tracking failed: %s\n", youth_cuda_last_error());
    ok = 0;
  }
  if (!ok) {
    /* work of this run may have been enqueued before the failure (H2D copies out of the ring slots): nothing
     * is released before the device is idle */
    youth_cuda_sync(G.h);
    atomic_store(&G.failed, 1);
  }
  if (ok) {
    atomic_store(&G.last_inliers, *inl[r->buf]);
    if (G.pose_mq != (mqd_t)-1) {
      /* pose egress: one MSG_TYPE_POSE message per tracked frame on the
       * logger->viewer queue; non-blocking, a full queue drops the pose rather than stalling tracking.
       * status = the frame's own YOUTH_STATUS_* bits; inliers = the inlier count of the run's LAST frame (the
       * tracker keeps one count per sequence, not per frame) */
      const uint32_t in = (uint32_t)*inl[r->buf];
      char msg[sizeof(MessageHeader) + sizeof(YouthPoseMsg)];
      for (int i = 0; i < r->n; ++i) {
        const size_t len = youth_pose_msg_build(msg, r->base + i, G.ts[r->first + i], poses[r->buf] + 12 * (size_t)i,
                                                status[r->buf][i], in);
        if (mq_send(G.pose_mq, msg, len, 0) == 0) ++sent;
      }
    }
  }
  pthread_mutex_lock(&G.mu);
  G.inflight -= r->n;
  G.done = (next && next->valid) ? next->pos : G.head; /* runs complete in order */
  G.poses_sent += sent;
  G.tracked += ok ? r->n : 0;
  pthread_cond_broadcast(&G.nonfull);
  if (waiting() == 0 && G.inflight == 0) pthread_cond_broadcast(&G.idle);
  pthread_mutex_unlock(&G.mu);
}

static void* worker_main(void* arg) {
  (void)arg;
  float* poses[2];
  uint32_t* status[2];
  int* inl[2];
  for (int k = 0; k < 2; ++k) {
    poses[k] = (float*)youth_cuda_host_alloc(sizeof(float) * 12 * (size_t)G.batch);
    status[k] = (uint32_t*)youth_cuda_host_alloc(sizeof(uint32_t) * (size_t)G.batch);
    inl[k] = (int*)youth_cuda_host_alloc(sizeof(int));
  }
  PendingRun pend;
  memset(&pend, 0, sizeof(pend));
  int turn = 0;
  for (;;) {
    pthread_mutex_lock(&G.mu);
    while (waiting() == 0 && !pend.valid && !atomic_load(&G.stop_req)) pthread_cond_wait(&G.nonempty, &G.mu);
    if (waiting() == 0 && !pend.valid) { /* stop requested and everything collected */
      pthread_mutex_unlock(&G.mu);
      break;
    }
    /* claim the longest run of consecutive slots: contiguous in the pinned ring, at most
     * one batch.  Claimed frames leave the queue at once; [done, head) fences their slots. */
    int n = waiting();
    if (n > G.batch) n = G.batch;
    const int first = (int)(G.head % (uint64_t)G.qcap);
    if (n > G.qcap - first) n = G.qcap - first;
    const uint64_t pos = G.head;
    G.head += (uint64_t)n;
    G.inflight += n;
    pthread_mutex_unlock(&G.mu);

    PendingRun cur;
    memset(&cur, 0, sizeof(cur));
    if (n > 0) {
      const uint16_t* src[1] = {G.ring + frame_px() * (size_t)first};
      cur.valid = 1;
      cur.n = n;
      cur.first = first;
      cur.pos = pos;
      cur.buf = turn;
      turn ^= 1;
      cur.ok = youth_cuda_track_batch(G.h, src, n, YOUTH_MEM_HOST_PINNED, G.ts + first, NULL);
      cur.base = youth_cuda_frame_count(G.h, 0) - n;
      if (cur.ok) {
        cur.ok = youth_cuda_read_last_inliers_async(G.h, 0, inl[cur.buf]) &&
                 youth_cuda_read_trajectory_async(G.h, 0, cur.base, n, poses[cur.buf], status[cur.buf], &cur.ticket) == n;
      }
      if (!cur.ok) fprintf(stderr, "AlgorithmModule: tracking failed: %s\n", youth_cuda_last_error());
    }
    if (pend.valid) collect_run(&pend, &cur, poses, status, inl);
    pend = cur;
  }
  for (int k = 0; k < 2; ++k) {
    youth_cuda_host_free(poses[k]);
    youth_cuda_host_free(status[k]);
    youth_cuda_host_free(inl[k]);
  }
  return NULL;
}

void initSlamModule(const char* config_file, const char* vocabulary_file) {
  (void)vocabulary_file; /* an ORB vocabulary has no meaning for a dense depth tracker */
  if (atomic_load(&G.running)) {
    fprintf(stderr, "AlgorithmModule: already running\n");
    return;
  }
  if (!youth_config_from_yaml(config_file, &G.cfg)) {
    fprintf(stderr, "AlgorithmModule: cannot read config '%s'\n", config_file ? config_file : "");
    return;
  }
  if (G.lossless < 0) {
    const char* e = getenv("YOUTH_SLAM_LOSSLESS");
    G.lossless = (e && *e == '1') ? 1 : 0;
  }
  if (G.batch < 0) {
    const char* e = getenv("YOUTH_SLAM_BATCH");
    G.batch = e ? atoi(e) : 8;
    if (G.batch < 1) G.batch = 1;
    if (G.batch > MAX_GROUP) G.batch = MAX_GROUP;
  }
  const char* dev = getenv("YOUTH_SLAM_DEVICE");
  G.cfg.device = dev ? atoi(dev) : 0;
  G.cfg.n_streams = 1;
  G.cfg.batch = G.batch;
  const char* cap = getenv("YOUTH_SLAM_TRAJ_CAPACITY");
  G.cfg.traj_capacity = cap ? atoi(cap) : (1 << 20); /* 9.7 h of 30 fps frames; 52 B per pose on the device */
  const char* ppt = getenv("YOUTH_SLAM_ICP_PPT"); /* reduction geometry (youth_cuda_config.icp_ppt), validated by init */
  if (ppt && atoi(ppt) > 0) G.cfg.icp_ppt = atoi(ppt);
  if (!youth_cuda_init(&G.cfg, &G.h)) {
    fprintf(stderr, "AlgorithmModule: failed to initialise the CUDA tracker: %s\n", youth_cuda_last_error());
    G.h = NULL;
    return;
  }
  {
    /* YOUTH_SLAM_MODE=model: track against a fused TSDF model instead of the previous frame
     * (include/youth_model.h) -- what TrackRGBD does with its map, SLAM.cpp:54 */
    const char* mode = getenv("YOUTH_SLAM_MODE");
    atomic_store(&G.model_on, 0);
    if (mode && !strcmp(mode, "model")) {
      youth_tsdf_config tc;
      youth_tsdf_default_config(&tc);
      if (!youth_cuda_enable_model(G.h, &tc)) {
        fprintf(stderr, "AlgorithmModule: cannot enable frame-to-model tracking: %s\n", youth_cuda_last_error());
        youth_cuda_destroy(G.h);
        G.h = NULL;
        return;
      }
      atomic_store(&G.model_on, 1);
    }
  }
  /* the waiting frames, the slot being filled, two runs in flight and one more being assembled always fit */
  G.qcap = QUEUE_HIGH_WATER + 4 + 3 * G.batch;
  G.ring = (uint16_t*)youth_cuda_host_alloc(frame_px() * sizeof(uint16_t) * (size_t)G.qcap);
  G.ts = (uint32_t*)calloc((size_t)G.qcap, sizeof(uint32_t));
  if (!G.ring || !G.ts) {
    fprintf(stderr, "AlgorithmModule: cannot allocate the host frame ring\n");
    youth_cuda_host_free(G.ring);
    free(G.ts);
    youth_cuda_destroy(G.h);
    G.h = NULL;
    G.ring = NULL;
    G.ts = NULL;
    return;
  }
  G.done = G.head = G.tail = 0;
  G.inflight = G.filling = 0;
  G.accepted = G.dropped = G.tracked = G.poses_sent = 0;
  G.pose_mq = (mqd_t)-1;
  {
    /* YOUTH_SLAM_POSE_MQ = queue name (e.g. /logger_viewer_queue): publish poses there */
    const char* q = getenv("YOUTH_SLAM_POSE_MQ");
    if (q && *q) {
      struct mq_attr attr;
      memset(&attr, 0, sizeof(attr));
      attr.mq_maxmsg = 10; /* the reference's queue geometry, loggingModule.c:137-141 */
      attr.mq_msgsize = MAX_MSG_SIZE;
      G.pose_mq = mq_open(q, O_WRONLY | O_NONBLOCK | O_CREAT, 0644, &attr);
      if (G.pose_mq == (mqd_t)-1) fprintf(stderr, "AlgorithmModule: cannot open pose queue %s (poses not published)\n", q);
    }
  }
  atomic_store(&G.last_inliers, 0);
  atomic_store(&G.stop_req, 0);
  atomic_store(&G.failed, 0);
  copy_pool_start();
  if (pthread_create(&G.worker, NULL, worker_main, NULL) != 0) {
    copy_pool_stop();
    fprintf(stderr, "AlgorithmModule: cannot start the worker thread\n");
    youth_cuda_host_free(G.ring);
    free(G.ts);
    youth_cuda_destroy(G.h);
    G.h = NULL;
    G.ring = NULL;
    G.ts = NULL;
    if (G.pose_mq != (mqd_t)-1) {
      mq_close(G.pose_mq);
      G.pose_mq = (mqd_t)-1;
    }
    return;
  }
  atomic_store(&G.running, 1);
}

void stopSlamModule(void) {
  if (!atomic_load(&G.running)) return;
  pthread_mutex_lock(&G.mu);
  atomic_store(&G.stop_req, 1);
  pthread_cond_broadcast(&G.nonempty);
  pthread_cond_broadcast(&G.nonfull);
  pthread_cond_broadcast(&G.prod);
  /* a producer that holds a ring slot finishes (its commit sees stop_req and discards the frame) */
  while (G.filling) pthread_cond_wait(&G.prod, &G.mu);
  pthread_mutex_unlock(&G.mu);
  pthread_join(G.worker, NULL); /* the worker drains what is queued before leaving */
  copy_pool_stop();                 /* no producer holds a slot any more (waited for above) */
  /* a producer that passed the `running` test before stop_req was set either holds the mutex now (wait for it
   * to leave) or will see stop_req once it has it */
  pthread_mutex_lock(&G.mu);
  atomic_store(&G.running, 0);
  youth_cuda_handle* h = G.h; /* threads that waited for the idle worker under the mutex find no handle from here on */
  uint16_t* ring = G.ring;
  uint32_t* ts = G.ts;
  G.h = NULL;
  G.ring = NULL;
  G.ts = NULL;
  pthread_mutex_unlock(&G.mu);
  youth_cuda_destroy(h);
  youth_cuda_host_free(ring);
  free(ts);
  if (G.pose_mq != (mqd_t)-1) {
    mq_close(G.pose_mq);
    G.pose_mq = (mqd_t)-1;
  }
}

/* ---- producer side: reserve a ring slot, fill it, publish it ---- */

/* Reserve the next ring slot for a frame of the configured size and return its address in page-locked host
 * memory, NULL when no frame can be taken (module not running / stopping / failed, wrong size; lossy mode with
 * the ring physically full: the frame counts as accepted-and-dropped).  The caller fills width*height uint16
 * depth values (mm) and calls youthSlamCommitSlot(); one slot is outstanding at a time -- a second producer
 * waits here until the first has committed.  This is the zero-copy form of processSlamFrame(): a producer that
 * assembles frames anyway (the logger's chunk reassembly, loggingModule.c:303-327; a file reader) writes them
 * where the copy engine reads them, instead of into a buffer of its own that processSlamFrame() then copies. */
uint16_t* youthSlamAcquireSlot(int width, int height) {
  if (!atomic_load(&G.running) || atomic_load(&G.failed)) return NULL;
  if (width != G.cfg.width || height != G.cfg.height) return NULL;
  pthread_mutex_lock(&G.mu);
  while (G.filling && !atomic_load(&G.stop_req)) pthread_cond_wait(&G.prod, &G.mu);
  if (G.lossless)
    while (G.tail - G.done >= (uint64_t)G.qcap - 1 && !atomic_load(&G.stop_req) && !atomic_load(&G.failed))
      pthread_cond_wait(&G.nonfull, &G.mu);
  if (atomic_load(&G.stop_req) || atomic_load(&G.failed) || !G.ring) { /* stopSlamModule() is under way or done */
    pthread_mutex_unlock(&G.mu);
    return NULL;
  }
  if (!G.lossless && waiting() > QUEUE_HIGH_WATER) {
    /* SLAM.cpp:163-167: more than 10 waiting -> drop the oldest down to 5.  The survivors (the newest 5, six or
     * more slots further on, so source and destination never overlap) move down to the head of the queue: the
     * ring keeps the shape [in flight][waiting][free], runs stay contiguous, and the write position can never
     * run round into frames the tracker still reads however much faster than the tracker the producer is.
     * (3 MB of copies under the mutex, only on the overloaded lossy path; no producer is filling a slot now.) */
    const int drop = waiting() - QUEUE_LOW_WATER;
    for (int k = 0; k < QUEUE_LOW_WATER; ++k) {
      const size_t from = (size_t)((G.head + (uint64_t)(drop + k)) % (uint64_t)G.qcap), to = (size_t)((G.head + (uint64_t)k) % (uint64_t)G.qcap);
      memcpy(G.ring + frame_px() * to, G.ring + frame_px() * from, frame_px() * sizeof(uint16_t));
      G.ts[to] = G.ts[from];
    }
    G.tail = G.head + QUEUE_LOW_WATER;
    G.dropped += drop;
  }
  if (G.tail - G.done >= (uint64_t)G.qcap - 1) { /* cannot happen in the lossy mode (at most 11 waiting + two runs in flight) */
    G.accepted++;
    G.dropped++;
    pthread_mutex_unlock(&G.mu);
    return NULL;
  }
  G.filling = 1;
  uint16_t* slot = G.ring + frame_px() * (size_t)(G.tail % (uint64_t)G.qcap);
  pthread_mutex_unlock(&G.mu);
  return slot;
}

/* Publish the slot returned by the last youthSlamAcquireSlot().  1 = queued for tracking, 0 = discarded (no slot
 * outstanding, or the module is stopping). */
int youthSlamCommitSlot(uint32_t timestamp) {
  pthread_mutex_lock(&G.mu);
  if (!G.filling) {
    pthread_mutex_unlock(&G.mu);
    return 0;
  }
  int ok = 0;
  if (!atomic_load(&G.stop_req) && G.ring) {
    G.ts[G.tail % (uint64_t)G.qcap] = timestamp;
    G.tail++;
    G.accepted++;
    if (waiting() == 1) pthread_cond_signal(&G.nonempty); /* the worker only sleeps on an empty queue */
    ok = 1;
  }
  G.filling = 0;
  pthread_cond_broadcast(&G.prod);
  pthread_mutex_unlock(&G.mu);
  return ok;
}

/* Give the slot of the last youthSlamAcquireSlot() back unpublished (the producer could not complete the frame). */
void youthSlamAbortSlot(void) {
  pthread_mutex_lock(&G.mu);
  if (G.filling) {
    G.filling = 0;
    pthread_cond_broadcast(&G.prod);
  }
  pthread_mutex_unlock(&G.mu);
}

int processSlamFrame(const int16_t* depth_data, const uint8_t* color_data, int width, int height, uint32_t timestamp) {
  (void)color_data; /* depth-only tracker; colour passes through the pipeline untouched */
  if (!depth_data) return 0;
  if (!atomic_load(&G.running) || width != G.cfg.width || height != G.cfg.height) return 0;
  long dropped_before = 0;
  if (!G.lossless) {
    pthread_mutex_lock(&G.mu);
    dropped_before = G.dropped;
    pthread_mutex_unlock(&G.mu);
  }
  uint16_t* slot = youthSlamAcquireSlot(width, height);
  if (!slot) {
    /* lossy mode with a full ring: the frame was accepted and dropped, like any other dropped frame */
    int was_dropped = 0;
    if (!G.lossless && atomic_load(&G.running) && !atomic_load(&G.stop_req) && !atomic_load(&G.failed)) {
      pthread_mutex_lock(&G.mu);
      was_dropped = G.dropped > dropped_before;
      pthread_mutex_unlock(&G.mu);
    }
    return was_dropped;
  }
  /* the synchronous copy of SLAM.cpp:133-134, outside the mutex: the caller may reuse its buffer on return.
   * reference depth is int16_t; values >= 32768 are reinterpreted as uint16 like the CV_16UC1 view at
   * SLAM.cpp:133 and then rejected by the depth_max gate */
  copy_in_parallel(slot, depth_data, frame_px() * sizeof(uint16_t));
  return youthSlamCommitSlot(timestamp);
}

/* ---- additions for headless drivers (not in the reference facade) ---- */

/* lossless: 1 = producer blocks instead of dropping; batch: frames per launch group.
 * Call before initSlamModule; -1 keeps the environment/default choice. */
void youthSlamSetOptions(int lossless, int batch) {
  if (lossless >= 0) G.lossless = lossless ? 1 : 0;
  if (batch >= 1) G.batch = batch > MAX_GROUP ? MAX_GROUP : batch;
}

/* The tracker handle is used by the worker thread; other threads touch it only while the worker has nothing
 * queued or in flight, and hold the mutex meanwhile so that it cannot claim new frames (a producer waits in
 * processSlamFrame for that long). */
static void lock_idle(void) {
  pthread_mutex_lock(&G.mu);
  while (waiting() > 0 || G.inflight > 0) pthread_cond_wait(&G.idle, &G.mu);
}

/* wait until every accepted frame has been tracked */
void youthSlamDrain(void) {
  if (!atomic_load(&G.running)) return;
  pthread_mutex_lock(&G.mu);
  while (waiting() > 0 || G.inflight > 0) pthread_cond_wait(&G.idle, &G.mu);
  pthread_mutex_unlock(&G.mu);
}

/* Replay of FRAME_TYPE_DEPTH_PACKED records: n YD16 streams back to back (offsets[n+1]) go to the device
 * packed and are unpacked there (youth_cuda_track_batch_packed).  Lossless by construction; frames queued
 * through processSlamFrame() before this call are tracked first, so the order of the recording is kept. */
int youthSlamProcessPackedFrames(const uint8_t* streams, const uint64_t* offsets, int n, int width, int height,
                                 const uint32_t* timestamps) {
  if (!atomic_load(&G.running) || !streams || !offsets || n < 1) return 0;
  if (width != G.cfg.width || height != G.cfg.height) return 0;
  float* poses = (float*)malloc(sizeof(float) * 12 * (size_t)G.batch);
  uint32_t* status = (uint32_t*)calloc((size_t)G.batch, sizeof(uint32_t));
  if (!poses || !status) {
    free(poses);
    free(status);
    return 0;
  }
  lock_idle(); /* the worker is idle: the handle is ours */
  int ok = G.h != NULL;
  for (int f0 = 0; f0 < n && ok; f0 += G.batch) {
    const int cn = n - f0 < G.batch ? n - f0 : G.batch;
    const uint8_t* sp[1] = {streams};
    const uint64_t* op[1] = {offsets + f0};
    ok = youth_cuda_track_batch_packed(G.h, sp, op, cn, YOUTH_MEM_HOST, timestamps ? timestamps + f0 : NULL, poses);
    if (!ok) {
      fprintf(stderr, "AlgorithmModule: packed tracking failed: %s\n", youth_cuda_last_error());
      break;
    }
    atomic_store(&G.last_inliers, youth_cuda_last_inliers(G.h, 0));
    G.accepted += cn;
    G.tracked += cn;
    if (G.pose_mq != (mqd_t)-1) {
      const int base = youth_cuda_frame_count(G.h, 0) - cn;
      const uint32_t inl = (uint32_t)atomic_load(&G.last_inliers);
      if (youth_cuda_get_trajectory(G.h, 0, base, cn, NULL, NULL, status) != cn) memset(status, 0, sizeof(uint32_t) * (size_t)cn);
      char msg[sizeof(MessageHeader) + sizeof(YouthPoseMsg)];
      for (int i = 0; i < cn; ++i) {
        const size_t len = youth_pose_msg_build(msg, base + i, timestamps ? timestamps[f0 + i] : 0u,
                                                poses + 12 * (size_t)i, status[i], inl);
        if (mq_send(G.pose_mq, msg, len, 0) == 0) G.poses_sent++;
      }
    }
  }
  pthread_mutex_unlock(&G.mu);
  free(poses);
  free(status);
  return ok;
}

/* Frames that already live in page-locked host memory (youth_cuda_host_alloc: a recording loaded there, a producer
 * whose own buffers are page-locked): no CPU copy at all -- the copy engine reads them where they are.  n frames
 * back to back, tracked in launch groups of the configured size with two groups in flight, in order after
 * everything queued through processSlamFrame().  Synchronous: returns when all n frames are tracked, so the caller
 * may reuse the memory (the ownership rule of SLAM.cpp:133-134, at run granularity). */
int youthSlamProcessPinnedFrames(const uint16_t* frames, int n, int width, int height, const uint32_t* timestamps) {
  if (!atomic_load(&G.running) || !frames || n < 1 || atomic_load(&G.failed)) return 0;
  if (width != G.cfg.width || height != G.cfg.height) return 0;
  float* poses[2];
  uint32_t* status[2];
  int* inl[2];
  for (int k = 0; k < 2; ++k) {
    poses[k] = (float*)youth_cuda_host_alloc(sizeof(float) * 12 * (size_t)G.batch);
    status[k] = (uint32_t*)youth_cuda_host_alloc(sizeof(uint32_t) * (size_t)G.batch);
    inl[k] = (int*)youth_cuda_host_alloc(sizeof(int));
  }
  lock_idle(); /* the worker is idle and cannot claim frames while we hold the mutex: the handle is ours */
  int ok = G.h != NULL && poses[0] && poses[1] && status[0] && status[1] && inl[0] && inl[1];
  struct {
    int valid, n, base, f0;
    uint64_t ticket;
  } pend = {0, 0, 0, 0, 0}, cur;
  int turn = 0;
  long sent = 0;
  for (int f0 = 0; ok && (f0 < n || pend.valid); f0 += G.batch) {
    memset(&cur, 0, sizeof(cur));
    if (f0 < n) {
      const int cn = n - f0 < G.batch ? n - f0 : G.batch;
      const uint16_t* src[1] = {frames + frame_px() * (size_t)f0};
      ok = youth_cuda_track_batch(G.h, src, cn, YOUTH_MEM_HOST_PINNED, timestamps ? timestamps + f0 : NULL, NULL);
      cur.base = youth_cuda_frame_count(G.h, 0) - cn;
      ok = ok && youth_cuda_read_last_inliers_async(G.h, 0, inl[turn]) &&
           youth_cuda_read_trajectory_async(G.h, 0, cur.base, cn, poses[turn], status[turn], &cur.ticket) == cn;
      cur.valid = ok;
      cur.n = cn;
      cur.f0 = f0;
      turn ^= 1;
    }
    if (pend.valid) {
      const int b = turn ^ (cur.valid ? 0 : 1); /* the buffer of the run submitted before `cur` */
      ok = youth_cuda_wait_ticket(G.h, pend.ticket) && ok;
      if (ok) {
        atomic_store(&G.last_inliers, *inl[b]);
        G.accepted += pend.n;
        G.tracked += pend.n;
        if (G.pose_mq != (mqd_t)-1) {
          char msg[sizeof(MessageHeader) + sizeof(YouthPoseMsg)];
          for (int i = 0; i < pend.n; ++i) {
            const size_t len = youth_pose_msg_build(msg, pend.base + i, timestamps ? timestamps[pend.f0 + i] : 0u,
                                                    poses[b] + 12 * (size_t)i, status[b][i], (uint32_t)*inl[b]);
            if (mq_send(G.pose_mq, msg, len, 0) == 0) ++sent;
          }
        }
      }
    }
    pend = cur;
  }
  if (!ok) {
    fprintf(stderr, "AlgorithmModule: tracking failed: %s\n", youth_cuda_last_error());
    if (G.h) youth_cuda_sync(G.h); /* nothing of the caller's memory is in flight when we return */
    atomic_store(&G.failed, 1);
  }
  G.poses_sent += sent;
  pthread_mutex_unlock(&G.mu);
  for (int k = 0; k < 2; ++k) {
    youth_cuda_host_free(poses[k]);
    youth_cuda_host_free(status[k]);
    youth_cuda_host_free(inl[k]);
  }
  return ok ? 1 : 0;
}

int youthSlamGetTrajectory(float* poses_out, uint32_t* timestamps_out, uint32_t* status_out, int max_frames) {
  if (!atomic_load(&G.running)) return 0;
  lock_idle();
  int n = G.h ? youth_cuda_get_trajectory(G.h, 0, 0, max_frames, poses_out, timestamps_out, status_out) : 0;
  pthread_mutex_unlock(&G.mu);
  return n < 0 ? 0 : n;
}

void youthSlamStats(long* accepted, long* dropped, long* tracked) {
  pthread_mutex_lock(&G.mu);
  if (accepted) *accepted = G.accepted;
  if (dropped) *dropped = G.dropped;
  if (tracked) *tracked = G.tracked;
  pthread_mutex_unlock(&G.mu);
}

int saveSlamMap(const char* map_file) {
  if (!atomic_load(&G.running) || !map_file) {
    fprintf(stderr, "AlgorithmModule: not running\n");
    return 0;
  }
  lock_idle(); /* every accepted frame is in the trajectory, and the handle is ours while we copy it out */
  const int n = G.h ? youth_cuda_frame_count(G.h, 0) : -1;
  float* poses = (float*)malloc(sizeof(float) * 12 * (size_t)(n > 0 ? n : 1));
  uint32_t* ts = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
  if (n < 0 || !poses || !ts) {
    pthread_mutex_unlock(&G.mu);
    free(poses);
    free(ts);
    return 0;
  }
  int got = youth_cuda_get_trajectory(G.h, 0, 0, n, poses, ts, NULL);
  pthread_mutex_unlock(&G.mu);
  if (got < 0) got = 0;
  char path[1024];
  snprintf(path, sizeof(path), "%s_trajectory.txt", map_file);
  int ok = youth_tum_write(path, poses, ts, got);
  /* key frames: first frame, then whenever the camera moved > 0.10 m or > 10 degrees */
  int nk = 0;
  for (int i = 0; i < got; ++i) {
    int take = (i == 0);
    if (!take) {
      const float* a = poses + 12 * (size_t)nk - 12; /* last key frame (compacted in place) */
      const float* b = poses + 12 * (size_t)i;
      const float dx = b[3] - a[3], dy = b[7] - a[7], dz = b[11] - a[11];
      float tr = 0.f; /* trace of Ra^T Rb */
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) tr += a[4 * r + c] * b[4 * r + c];
      take = (dx * dx + dy * dy + dz * dz > 0.01f) || (tr < 1.0f + 2.0f * 0.98480775f);
    }
    if (take) {
      memmove(poses + 12 * (size_t)nk, poses + 12 * (size_t)i, sizeof(float) * 12);
      ts[nk] = ts[i];
      ++nk;
    }
  }
  snprintf(path, sizeof(path), "%s_keyframes.txt", map_file);
  ok = youth_tum_write(path, poses, ts, nk) && ok;
  free(poses);
  free(ts);
  return ok ? 1 : 0;
}

int isSlamModuleRunning(void) { return atomic_load(&G.running) ? 1 : 0; }

int getSlamMapPoints(void) {
  if (!atomic_load(&G.running)) return 0;
  if (atomic_load(&G.model_on)) {
    /* frame-to-model: the map has a size of its own -- the voxels the fused surface passes through
     * (GetAllMapPoints().size(), SLAM.cpp:212-217) */
    lock_idle(); /* the worker is idle: the handle is ours */
    const long long n = G.h ? youth_cuda_model_surface_voxels(G.h, 0) : 0;
    pthread_mutex_unlock(&G.mu);
    return n < 0 ? 0 : (n > 2147483647LL ? 2147483647 : (int)n);
  }
  return atomic_load(&G.last_inliers);
}

void resetSlam(void) {
  if (!atomic_load(&G.running)) return;
  lock_idle();
  if (G.h) youth_cuda_reset(G.h, -1);
  atomic_store(&G.last_inliers, 0);
  atomic_store(&G.failed, 0); /* the trajectory is empty again: the tracker can take frames */
  pthread_mutex_unlock(&G.mu);
}
