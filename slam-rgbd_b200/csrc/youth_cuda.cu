/*
 * youth_cuda.cu -- libyouth_cuda.so: the C ABI declared in include/youth_cuda.h over the
 * sm_100a kernels in youth_ingest.cu, youth_icp.cu, youth_model.cuh and youth_codec.cuh.  No CPU fallback: every compute entry point
 * fails (returns 0, youth_cuda_last_error() says why) when no CUDA device is usable.
 *
 * Device-resident state per handle (all in HBM, sized at init, nothing allocated per frame):
 *   ring      per stream R = batch+1 slots; per slot and pyramid level: float depth,
 *             three float2 planes (vx,vy) (vz,nx) (ny,nz) (24 B/pixel), uint8 pyramid count
 *   raw       2 x [S][batch] uint16 frames (double-buffered H2D landing zone)
 *   pairs     per (stream, frame-in-group): double+float relative pose, per-run partial
 *             sums [max_runs][32] float, reduced sums [32] double, status
 *   sequence  per stream: frame count, world pose (double), trajectory [cap][12] float,
 *             status [cap], last inlier count
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "youth_cuda.h"
#include "youth_codec.h"
#include "youth_common.cuh"
#include "youth_model.cuh"

#define YK_CHUNK_FRAMES 16 /* host-fed groups are copied + preprocessed in chunks of at least this many frames */
#define YK_MAX_CHUNKS 64
#define YK_GRAPH_MAX_N 8   /* groups of at most this many frames per stream run as one captured CUDA graph */
#define YK_GRAPH_SLOTS (2 * YK_GRAPH_MAX_N)
#define YK_ICP_QUEUES 8
#define YK_TICKETS (4 * YK_MAX_STREAMS) /* two steps in flight with one read per stream each, twice over */
#define YK_ICP_LAST_CTA_MAX_CTAS 444 /* one wave of the 152-register variant: 3 CTAs on each of 148 SMs */

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 0;
}

#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

extern "C" const char* youth_cuda_last_error(void) { return g_err; }
extern "C" int youth_cuda_abi_version(void) { return YOUTH_CUDA_ABI_VERSION; }

/* ------------------------------------------------------------------ handle */

struct youth_cuda_handle {
  youth_cuda_config cfg;
  int S, B, R, P; /* streams, batch, ring slots, max pairs */
  LevelGeom lv[YOUTH_MAX_LEVELS];
  int npix[YOUTH_MAX_LEVELS];
  int ppr[YOUTH_MAX_LEVELS];   /* pixels per lane per run */
  int nruns[YOUTH_MAX_LEVELS]; /* runs per frame pair */
  int max_runs;
  cudaStream_t stream, copy_stream;
  bool own_stream;
  /* pair-group schedule of stages 3-5 (youth_cuda_set_icp_schedule): all iterations of `icp_group` pairs are
   * enqueued back to back on one of `icp_nq` side streams, so the maps of a group stay in L2 across iterations */
  int icp_group, icp_nq;
  cudaStream_t icp_q[YK_ICP_QUEUES];
  cudaEvent_t icp_fork, icp_join[YK_ICP_QUEUES];
  /* ring */
  float* depth[YOUTH_MAX_LEVELS];
  float2* maps[YOUTH_MAX_LEVELS]; /* [S][R][3][npix]: (vx,vy) (vz,nx) (ny,nz) planes */
  uint8_t* pyrcnt[YOUTH_MAX_LEVELS];
  /* raw landing zone */
  uint16_t* raw[2];
  uint16_t* pinned[2];
  cudaEvent_t raw_free[2], chunk_ready[2][YK_MAX_CHUNKS];
  bool raw_used[2];
  int raw_turn;
  /* bilateral tables */
  float ws[49];
  float* wr;
  float* wt; /* product table of the bilateral (IngestParams.wt) */
  /* reciprocal form of the vertex divisions (div_cfg), used when the exhaustive device check at init passed */
  int fast_div;
  float r_df, r_fx[YOUTH_MAX_LEVELS], r_fy[YOUTH_MAX_LEVELS];
  /* the float depth pyramid and the pyramid sample counts are only stored (and their buffers only exist) for the
   * parity read-back and for the model ray cast's depth hint: youth_cuda_debug_enable_maps / youth_cuda_enable_model */
  bool debug_maps;
  int range_cut;
  cudaEvent_t ticket_ev[YK_TICKETS]; /* youth_cuda_read_trajectory_async / youth_cuda_wait_ticket */
  uint64_t ticket_next;
  int host_range; /* YOUTH_HOST_RANGE: frames per tracking range of host-fed groups (0 = whole group, see host_range_frames) */
  bool ingest_generic; /* YOUTH_INGEST_GENERIC=1: force the per-tap-product bilateral (A/B runs, tests) */
  /* k_icp_fused: the whole iteration schedule of a few pairs in one launch (YOUTH_ICP_FUSED=1 turns it on,
   * YOUTH_ICP_FUSED_COOP=0 launches it without the cooperative attribute) */
  bool fused, fused_coop;
  int fused_max_ctas;  /* CTAs of k_icp_fused the device holds at once */
  unsigned int* gen;   /* [P] per-pair generation counters, zero between launches */
  /* pair state */
  double* pose_d;
  float* pose_f;
  float* partials;
  double* sums;
  uint32_t* pair_status;
  unsigned int* tickets;
  /* sequence state */
  int* seq_count;
  double* world;
  float* traj;
  uint32_t* traj_status;
  int* last_inliers;
  int* d_head;       /* device ring head (slot of the next frame) */
  int* h_count;      /* host mirror of seq_count */
  uint32_t* h_ts;    /* [S][cap] timestamps (host only) */
  long long total;   /* frames per stream since init / full reset (ring position) */
  int32_t* corr_dbg; /* debug correspondence map (level-0 sized) */
  cudaEvent_t t0, t1;
  uint64_t launches;
  /* CUDA graphs of the whole kernel schedule for small host-fed groups (launch-bound live path) */
  struct {
    cudaGraphExec_t exec;
    int n, k;
    uint64_t launches; /* kernels inside the graph */
  } graphs[YK_GRAPH_SLOTS];
  bool graphs_enabled;
  /* per-kernel-class event timing (youth_cuda_profile_*): off on the normal path */
  bool prof_on;
  cudaEvent_t* prof_ev; /* pairs: [2*i] before, [2*i+1] after launch i */
  int* prof_cls;
  int prof_n, prof_cap;
  double prof_ms[YOUTH_PROF_CLASSES];
  uint64_t prof_launches[YOUTH_PROF_CLASSES];
  /* frame-to-model tracking (include/youth_model.h), off until youth_cuda_enable_model */
  struct {
    bool on;
    youth_tsdf_config cfg;
    TsdfGeom geom;
    short2* vol;                     /* [S][dz][dy][dx] */
    float2* maps[YOUTH_MAX_LEVELS];  /* [S][3][npix] ray-cast model maps */
    float* world_f;                  /* [S][12] */
    uint32_t* last_status;           /* [S] */
    cudaGraphExec_t graph;           /* one frame of every sequence: stages 1-5, fusion, ray cast */
    uint64_t graph_launches;
  } m;
  /* packed (YD16) input path, created on first use (youth_cuda_track_batch_packed) */
  struct youth_codec* codec;
  unsigned long long* pk_h_off[2]; /* pinned: device offsets of the group's streams */
  unsigned long long* pk_d_off[2];
  cudaEvent_t pk_free;             /* the group's last decode has read the packed bytes */
  bool pk_used;
};

/* bracket one launch with events when profiling is on */
struct ProfScope {
  youth_cuda_handle* h;
  int idx;
  ProfScope(youth_cuda_handle* hh, int cls) : h(hh), idx(-1) {
    h->launches++;
    if (!h->prof_on || h->prof_n >= h->prof_cap) return;
    idx = h->prof_n++;
    h->prof_cls[idx] = cls;
    cudaEventRecord(h->prof_ev[2 * idx], h->stream);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(h->prof_ev[2 * idx + 1], h->stream);
  }
};

static int prof_flush(youth_cuda_handle* h) {
  if (h->prof_n == 0) return 1;
  CU(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < h->prof_n; ++i) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
    h->prof_ms[h->prof_cls[i]] += (double)ms;
    h->prof_launches[h->prof_cls[i]]++;
  }
  h->prof_n = 0;
  return 1;
}

extern "C" int youth_cuda_default_config(youth_cuda_config* c) {
  if (!c) return fail("null config");
  memset(c, 0, sizeof(*c));
  c->width = 640;
  c->height = 480;
  c->fx = 570.3f;
  c->fy = 570.3f;
  c->cx = 320.0f;
  c->cy = 240.0f;
  c->depth_factor = 1000.0f;
  c->levels = 3;
  c->iters[0] = 10;
  c->iters[1] = 5;
  c->iters[2] = 4;
  c->iters[3] = 4;
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
  c->device = 0;
  c->stream = NULL;
  return 1;
}

static int validate(const youth_cuda_config* c) {
  if (c->levels < 1 || c->levels > YOUTH_MAX_LEVELS) return fail("levels must be 1..%d", YOUTH_MAX_LEVELS);
  const int div = 1 << (c->levels - 1);
  if (c->width <= 0 || c->height <= 0 || c->width % 8 || c->width % div || c->height % div)
    return fail("width/height must be positive, width %% 8 == 0, both divisible by 2^(levels-1)");
  if (c->width > 16384 || c->height > 16384) return fail("image too large");
  if (!(c->fx > 0.f) || !(c->fy > 0.f) || !(c->depth_factor > 0.f)) return fail("bad intrinsics / depth factor");
  if (c->depth_min_mm < 1 || c->depth_max_mm > 65535 || c->depth_min_mm > c->depth_max_mm)
    return fail("depth range must satisfy 1 <= min <= max <= 65535");
  if (c->icp_ppt < 1 || c->icp_ppt > 256 || (c->icp_ppt & (c->icp_ppt - 1))) return fail("icp_ppt must be a power of two in 1..256");
  if (c->n_streams < 1 || c->n_streams > YK_MAX_STREAMS) return fail("n_streams must be 1..%d", YK_MAX_STREAMS);
  if (c->batch < 1 || c->batch > 1024) return fail("batch must be 1..1024");
  if (c->traj_capacity < 1) return fail("traj_capacity must be >= 1");
  if (c->bilateral && !(c->sigma_space_px > 0.f && c->sigma_range_mm > 0.f)) return fail("bilateral sigmas must be > 0");
  for (int l = 0; l < c->levels; ++l)
    if (c->iters[l] < 0 || c->iters[l] > 1000) return fail("iters[%d] out of range", l);
  return 1;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t count) {
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
  return e;
}

#include "youth_codec.cuh"

extern "C" void youth_cuda_destroy(youth_cuda_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  for (int l = 0; l < YOUTH_MAX_LEVELS; ++l) {
    cudaFree(h->depth[l]);
    cudaFree(h->maps[l]);
    cudaFree(h->pyrcnt[l]);
  }
  for (int k = 0; k < 2; ++k) {
    cudaFree(h->raw[k]);
    if (h->pinned[k]) cudaFreeHost(h->pinned[k]);
    if (h->raw_free[k]) cudaEventDestroy(h->raw_free[k]);
    for (int c2 = 0; c2 < YK_MAX_CHUNKS; ++c2)
      if (h->chunk_ready[k][c2]) cudaEventDestroy(h->chunk_ready[k][c2]);
  }
  cudaFree(h->wr);
  cudaFree(h->wt);
  cudaFree(h->pose_d);
  cudaFree(h->pose_f);
  cudaFree(h->partials);
  cudaFree(h->sums);
  cudaFree(h->pair_status);
  cudaFree(h->tickets);
  cudaFree(h->gen);
  cudaFree(h->seq_count);
  cudaFree(h->world);
  cudaFree(h->traj);
  cudaFree(h->traj_status);
  cudaFree(h->last_inliers);
  cudaFree(h->d_head);
  for (int g2 = 0; g2 < YK_GRAPH_SLOTS; ++g2)
    if (h->graphs[g2].exec) cudaGraphExecDestroy(h->graphs[g2].exec);
  cudaFree(h->corr_dbg);
  cudaFree(h->m.vol);
  for (int l = 0; l < YOUTH_MAX_LEVELS; ++l) cudaFree(h->m.maps[l]);
  cudaFree(h->m.world_f);
  cudaFree(h->m.last_status);
  if (h->m.graph) cudaGraphExecDestroy(h->m.graph);
  if (h->codec) youth_codec_destroy(h->codec);
  for (int k = 0; k < 2; ++k) {
    if (h->pk_h_off[k]) cudaFreeHost(h->pk_h_off[k]);
    cudaFree(h->pk_d_off[k]);
  }
  if (h->pk_free) cudaEventDestroy(h->pk_free);
  if (h->prof_ev) {
    for (int i = 0; i < 2 * h->prof_cap; ++i) cudaEventDestroy(h->prof_ev[i]);
    free(h->prof_ev);
    free(h->prof_cls);
  }
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->t1) cudaEventDestroy(h->t1);
  for (int k = 0; k < YK_ICP_QUEUES; ++k) {
    if (h->icp_q[k]) {
      cudaStreamSynchronize(h->icp_q[k]);
      cudaStreamDestroy(h->icp_q[k]);
    }
    if (h->icp_join[k]) cudaEventDestroy(h->icp_join[k]);
  }
  if (h->icp_fork) cudaEventDestroy(h->icp_fork);
  for (int k = 0; k < YK_TICKETS; ++k)
    if (h->ticket_ev[k]) cudaEventDestroy(h->ticket_ev[k]);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  free(h->h_count);
  free(h->h_ts);
  delete h;
}

/* Page-locked staging for pageable input (YOUTH_MEM_HOST): 2 x P frames.  Callers that only ever hand over
 * page-locked or device memory (the facade's ring, bench.py, the multi-GPU driver) never pin it. */
static int ensure_staging(youth_cuda_handle* h) {
  const size_t bytes = (size_t)h->P * h->cfg.width * h->cfg.height * sizeof(uint16_t);
  for (int k = 0; k < 2; ++k)
    if (!h->pinned[k]) CU(cudaHostAlloc((void**)&h->pinned[k], bytes, cudaHostAllocDefault));
  return 1;
}

/* buffers + kernel instantiation for the float depth pyramid and the pyramid sample counts (parity read-back, model
 * ray-cast hint).  Cached graphs were captured with the instantiation that does not store them: drop them. */
static int ensure_debug_maps(youth_cuda_handle* h) {
  if (h->debug_maps) return 1;
  const size_t slots = (size_t)h->S * h->R;
  for (int l = 0; l < h->cfg.levels; ++l) {
    CU(dalloc(&h->depth[l], slots * h->npix[l]));
    CU(dalloc(&h->pyrcnt[l], slots * h->npix[l]));
  }
  for (int g2 = 0; g2 < YK_GRAPH_SLOTS; ++g2)
    if (h->graphs[g2].exec) {
      cudaGraphExecDestroy(h->graphs[g2].exec);
      h->graphs[g2].exec = NULL;
    }
  if (h->m.graph) {
    cudaGraphExecDestroy(h->m.graph);
    h->m.graph = NULL;
  }
  h->debug_maps = true;
  return 1;
}

extern "C" int youth_cuda_debug_enable_maps(youth_cuda_handle* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  return ensure_debug_maps(h);
}

static int init_impl(const youth_cuda_config* cfg, youth_cuda_handle* h) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("no CUDA device available (%s); libyouth_cuda has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (have %d)", cfg->device, ndev);
  CU(cudaSetDevice(cfg->device));
  h->cfg = *cfg;
  h->S = cfg->n_streams;
  h->B = cfg->batch;
  h->R = cfg->batch + 1;
  h->P = h->S * h->B;
  /* per-level geometry: identical float operation sequence to the checker */
  {
    int w = cfg->width, hh = cfg->height;
    float fx = cfg->fx, fy = cfg->fy, cx = cfg->cx, cy = cfg->cy;
    for (int l = 0; l < cfg->levels; ++l) {
      if (l > 0) {
        w /= 2;
        hh /= 2;
        fx = fx * 0.5f;
        fy = fy * 0.5f;
        cx = (cx - 0.5f) * 0.5f;
        cy = (cy - 0.5f) * 0.5f;
      }
      h->lv[l] = LevelGeom{w, hh, fx, fy, cx, cy, cx + 0.5f, cy + 0.5f};
      h->npix[l] = w * hh;
    }
  }
  h->max_runs = 0;
  for (int l = 0; l < cfg->levels; ++l) {
    /* reduction geometry (part of the spec): a run is 32 * ppr consecutive pixels, ppr shrinks
     * 4x per level so that every level has about the same number of runs */
    int ppr = cfg->icp_ppt >> (2 * l);
    if (ppr < 1) ppr = 1;
    h->ppr[l] = ppr;
    h->nruns[l] = (h->npix[l] + 32 * ppr - 1) / (32 * ppr);
    if (h->nruns[l] > h->max_runs) h->max_runs = h->nruns[l];
  }
  if (cfg->stream) {
    h->stream = (cudaStream_t)cfg->stream;
    h->own_stream = false;
  } else {
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < YK_ICP_QUEUES; ++k) {
    CU(cudaStreamCreateWithFlags(&h->icp_q[k], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->icp_join[k], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&h->icp_fork, cudaEventDisableTiming));
  {
    const char* g = getenv("YOUTH_INGEST_GENERIC");
    h->ingest_generic = g && *g == '1';
    const char* r = getenv("YOUTH_HOST_RANGE");
    h->host_range = r ? atoi(r) : 0;
  }
  h->icp_group = 0;
  h->icp_nq = 1;
  const size_t slots = (size_t)h->S * h->R;
  for (int l = 0; l < cfg->levels; ++l) {
    CU(dalloc(&h->maps[l], slots * 3 * h->npix[l]));
  }
  const size_t frame_px = (size_t)cfg->width * cfg->height;
  for (int k = 0; k < 2; ++k) {
    CU(dalloc(&h->raw[k], (size_t)h->P * frame_px));
    h->pinned[k] = NULL; /* staging for pageable input: page-locked on first YOUTH_MEM_HOST call (ensure_staging) */
    CU(cudaEventCreateWithFlags(&h->raw_free[k], cudaEventDisableTiming));
    for (int c2 = 0; c2 < YK_MAX_CHUNKS; ++c2) CU(cudaEventCreateWithFlags(&h->chunk_ready[k][c2], cudaEventDisableTiming));
  }
  /* bilateral tables (the only libm use; depends on the config alone) */
  {
    int cut = (int)(3.0f * cfg->sigma_range_mm);
    if (cut > YK_RANGE_LUT_MAX - 2) cut = YK_RANGE_LUT_MAX - 2;
    if (cut < 0) cut = 0;
    h->range_cut = cut;
    const double ss = (double)cfg->sigma_space_px, sr = (double)cfg->sigma_range_mm;
    for (int dy = -3; dy <= 3; ++dy)
      for (int dx = -3; dx <= 3; ++dx)
        h->ws[(dy + 3) * 7 + dx + 3] = cfg->bilateral ? (float)exp(-(double)(dx * dx + dy * dy) / (2.0 * ss * ss)) : 0.f;
    float wr[YK_RANGE_LUT_MAX];
    for (int i = 0; i <= cut; ++i) wr[i] = cfg->bilateral ? (float)exp(-((double)i * (double)i) / (2.0 * sr * sr)) : 0.f;
    wr[cut + 1] = 0.0f;
    CU(dalloc(&h->wr, (size_t)YK_RANGE_LUT_MAX));
    CU(cudaMemcpy(h->wr, wr, sizeof(float) * (cut + 2), cudaMemcpyHostToDevice));
    /* product table of k_ingest<YK_INGEST_BILATERAL_WT>: one IEEE single-precision multiply per entry, the
     * product the generic path (and the CPU statement) forms per tap */
    CU(dalloc(&h->wt, (size_t)YK_WT_ROWS * YK_WT_STRIDE));
    if (cut + 2 <= YK_WT_STRIDE) {
      float wt[YK_WT_ROWS * YK_WT_STRIDE];
      for (int row = 0; row < YK_WT_ROWS; ++row) {
        const float w = h->ws[(row / 4 + 3) * 7 + (row % 4) + 3]; /* class (|dy|, |dx|) = (row / 4, row % 4) */
        for (int i = 0; i < YK_WT_STRIDE; ++i) wt[row * YK_WT_STRIDE + i] = i < cut + 2 ? w * wr[i] : 0.0f;
      }
      CU(cudaMemcpy(h->wt, wt, sizeof(wt), cudaMemcpyHostToDevice));
    }
  }
  h->fast_div = 0;
  {
    /* The divisions of the back-projection by depth_factor, fx, fy (viewerModule.c:343-345) in the reciprocal form
     * div_cfg, when the device confirms -- exhaustively, here -- that it gives the IEEE quotient for every dividend
     * that can occur: d in [1, 65535], and (u - c) * z with z = d / depth_factor; with the bounds below they are 0
     * or lie in [2^-53, 2^51], inside the range k_div_check covers.  YOUTH_NO_FAST_DIV=1 keeps the divisions. */
    const float df = cfg->depth_factor;
    bool ok = df >= 0x1p-20f && df <= 0x1p20f && !getenv("YOUTH_NO_FAST_DIV");
    unsigned long long* d_bad = NULL;
    unsigned long long bad = 0;
    CU(cudaMalloc((void**)&d_bad, sizeof(bad)));
    CU(cudaMemset(d_bad, 0, sizeof(bad)));
    h->r_df = 1.0f / df;
    float checked[1 + 2 * YOUTH_MAX_LEVELS]; /* every distinct divisor is checked once (fx == fy for the Astra) */
    int n_checked = 0;
    auto check_once = [&](float b, float r) {
      for (int k = 0; k < n_checked; ++k)
        if (checked[k] == b) return;
      checked[n_checked++] = b;
      yk_launch_div_check(b, r, d_bad);
    };
    if (ok) check_once(df, h->r_df);
    for (int l = 0; l < cfg->levels; ++l) {
      const LevelGeom& g = h->lv[l];
      ok = ok && g.fx >= 0x1p-10f && g.fx <= 0x1p20f && g.fy >= 0x1p-10f && g.fy <= 0x1p20f &&
           (g.cx == 0.0f || (fabsf(g.cx) >= 0x1p-10f && fabsf(g.cx) <= 0x1p20f)) &&
           (g.cy == 0.0f || (fabsf(g.cy) >= 0x1p-10f && fabsf(g.cy) <= 0x1p20f));
      h->r_fx[l] = 1.0f / g.fx;
      h->r_fy[l] = 1.0f / g.fy;
      if (ok) {
        check_once(g.fx, h->r_fx[l]);
        check_once(g.fy, h->r_fy[l]);
      }
    }
    CU(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    CU(cudaGetLastError());
    h->fast_div = ok && bad == 0 ? 1 : 0;
  }
  h->debug_maps = false;
  {
    const char* e = getenv("YOUTH_DEBUG_MAPS");
    if (e && *e == '1' && !ensure_debug_maps(h)) return 0;
  }
  CU(dalloc(&h->pose_d, (size_t)h->P * 12));
  CU(dalloc(&h->pose_f, (size_t)h->P * 12));
  CU(dalloc(&h->partials, (size_t)h->P * h->max_runs * 32));
  CU(dalloc(&h->sums, (size_t)h->P * 32));
  CU(dalloc(&h->pair_status, (size_t)h->P));
  CU(dalloc(&h->tickets, (size_t)h->P));
  CU(dalloc(&h->gen, (size_t)h->P));
  {
    const char* f = getenv("YOUTH_ICP_FUSED");
    h->fused = f && *f == '1'; /* measured slower than one launch per iteration (profiles/README.md): opt-in */
    const char* fc = getenv("YOUTH_ICP_FUSED_COOP");
    h->fused_coop = !(fc && *fc == '0');
    int per_sm = 0, sms = 0, coop = 0;
    CU(yk_icp_fused_ctas_per_sm(&per_sm));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device));
    h->fused_max_ctas = per_sm * sms;
    if (!coop) h->fused_coop = false;
  }
  CU(dalloc(&h->seq_count, (size_t)h->S));
  CU(dalloc(&h->world, (size_t)h->S * 12));
  CU(dalloc(&h->traj, (size_t)h->S * cfg->traj_capacity * 12));
  CU(dalloc(&h->traj_status, (size_t)h->S * cfg->traj_capacity));
  CU(dalloc(&h->last_inliers, (size_t)h->S));
  CU(dalloc(&h->d_head, (size_t)1));
  CU(dalloc(&h->corr_dbg, frame_px));
  h->h_count = (int*)calloc(h->S, sizeof(int));
  h->h_ts = (uint32_t*)calloc((size_t)h->S * cfg->traj_capacity, sizeof(uint32_t));
  if (!h->h_count || !h->h_ts) return fail("host allocation failed");
  {
    const char* g = getenv("YOUTH_CUDA_GRAPHS");
    h->graphs_enabled = !(g && *g == '0');
  }
  CU(cudaEventCreate(&h->t0));
  CU(cudaEventCreate(&h->t1));
  CU(cudaDeviceSynchronize());
  return 1;
}

extern "C" int youth_cuda_init(const youth_cuda_config* cfg, youth_cuda_handle** out) {
  if (!cfg || !out) return fail("null argument");
  *out = NULL;
  if (!validate(cfg)) return 0;
  youth_cuda_handle* h = new youth_cuda_handle();
  memset((void*)h, 0, sizeof(*h));
  if (!init_impl(cfg, h)) {
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    youth_cuda_destroy(h);
    memcpy(g_err, keep, sizeof(keep));
    return 0;
  }
  *out = h;
  return 1;
}

extern "C" int youth_cuda_set_icp_schedule(youth_cuda_handle* h, int pairs_per_group, int queues) {
  if (!h) return fail("null handle");
  if (pairs_per_group < 0) return fail("pairs_per_group must be >= 0");
  if (queues < 1 || queues > YK_ICP_QUEUES) return fail("queues must be 1..%d", YK_ICP_QUEUES);
  h->icp_group = pairs_per_group;
  h->icp_nq = queues;
  return 1;
}

/* Pageable host frames are staged through page-locked memory before the call returns (the caller may reuse its
 * buffer, SLAM.cpp:133-134).  A large group (184 MB for 300 VGA frames) is far beyond the CPU caches and is read
 * next by the copy engine, not by this core: non-temporal stores skip the read-for-ownership of the destination
 * lines (measured on a Xeon host with the facade's identical copy-in: 13.0 instead of 6.5 GB/s).  Small copies
 * (live frames) stay with memcpy and the cache. */
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
static void stage_copy(void* dst, const void* src, size_t bytes) {
#if defined(__SSE2__)
  if (bytes >= ((size_t)4 << 20) && ((uintptr_t)dst & 15u) == 0) {
    char* d = (char*)dst;
    const char* s = (const char*)src;
    size_t i = 0;
    for (; i + 64 <= bytes; i += 64) {
      const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 16));
      const __m128i c = _mm_loadu_si128((const __m128i*)(s + i + 32)), e = _mm_loadu_si128((const __m128i*)(s + i + 48));
      _mm_stream_si128((__m128i*)(d + i), a);
      _mm_stream_si128((__m128i*)(d + i + 16), b);
      _mm_stream_si128((__m128i*)(d + i + 32), c);
      _mm_stream_si128((__m128i*)(d + i + 48), e);
    }
    _mm_sfence(); /* before the cudaMemcpyAsync that hands the buffer to the copy engine */
    if (i < bytes) memcpy(d + i, s + i, bytes - i);
    return;
  }
#endif
  memcpy(dst, src, bytes);
}

/* ------------------------------------------------------------------ launches */

static RingGeom ring_of(const youth_cuda_handle* h, int n) {
  RingGeom r;
  r.n = n;
  r.head = h->d_head; /* == total % R, kept on the device */
  r.R = h->R;
  r.S = h->S;
  return r;
}

/* one iteration of the pairs ip.f0 .. ip.f0 + ip.fn - 1 of every sequence (`pairs` = S * fn) on stream q.
 * Few pairs per launch (live / frame-to-model): the latency of the tail matters, share it between the warps
 * of the last CTA; many pairs: no block barrier, the tails of different pairs overlap anyway. */
static void launch_icp_on(youth_cuda_handle* h, const IcpParams& ip, int fn, int level, cudaStream_t q) {
  /* grid = (CTAs of a pair, frames of the range, sequences): the kernel reads its pair off blockIdx without a
   * division (the division's live range cost k_icp eight spill instructions per two pixels) */
  const dim3 grid((h->nruns[level] + YK_ICP_WARPS - 1) / YK_ICP_WARPS, fn, h->S);
  yk_launch_icp((long long)grid.x * fn * h->S <= YK_ICP_LAST_CTA_MAX_CTAS, grid, q, ip);
}

/* grid of one k_icp_fused launch over fn frames of every sequence, 0 when the fused kernel does not apply:
 * turned off, a profiled step (per-level launch times are wanted), no iterations, or more CTAs than the
 * device holds at once (the CTAs of a pair wait for each other inside the kernel) */
static int fused_grid_x(const youth_cuda_handle* h, int fn) {
  if (!h->fused || h->prof_on) return 0;
  int gx = 0;
  for (int l = 0; l < h->cfg.levels; ++l)
    if (h->cfg.iters[l] > 0) {
      const int c = (h->nruns[l] + YK_ICP_WARPS - 1) / YK_ICP_WARPS;
      if (c > gx) gx = c;
    }
  /* Launches captured into a CUDA graph (live frames, frame-to-model) and launches without the cooperative
   * attribute have no residency guarantee from the driver: they stay below half of the device, so that the
   * fused launches of two handles running at once still fit together. */
  const bool guaranteed = h->fused_coop && !(h->graphs_enabled && (h->m.on || fn <= YK_GRAPH_MAX_N));
  const long long cap = guaranteed ? h->fused_max_ctas : h->fused_max_ctas / 2;
  if (gx == 0 || (long long)gx * fn * h->S > cap) return 0;
  return gx;
}

/* all iterations of all levels for frames [f0, f0 + fn) of every sequence's group, one launch on stream q */
static int launch_icp_fused(youth_cuda_handle* h, const RingGeom& ring, int gx, int f0, int fn, cudaStream_t q) {
  const youth_cuda_config& c = h->cfg;
  IcpFusedParams fp;
  memset(&fp, 0, sizeof(fp));
  for (int level = c.levels - 1; level >= 0; --level) {
    if (c.iters[level] <= 0) continue;
    IcpFusedLevel& L = fp.lv[fp.nlv++];
    L.maps = h->maps[level];
    L.model = h->m.on ? h->m.maps[level] : NULL;
    L.g = h->lv[level];
    L.npix = h->npix[level];
    L.ppr = h->ppr[level];
    L.nruns = h->nruns[level];
    L.iters = c.iters[level];
  }
  fp.ring = ring;
  fp.max_runs = h->max_runs;
  fp.dist2_thr = c.dist_thresh_m * c.dist_thresh_m;
  fp.cos_thr = c.cos_thresh;
  fp.pose_f = h->pose_f;
  fp.pose_d = h->pose_d;
  fp.seq_count = h->seq_count;
  fp.partials = h->partials;
  fp.tickets = h->tickets;
  fp.gen = h->gen;
  fp.sums = h->sums;
  fp.pair_status = h->pair_status;
  fp.min_inliers = c.min_inliers;
  fp.f0 = f0;
  bool coop = h->fused_coop;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CU(cudaStreamIsCapturing(q, &cap));
  if (cap != cudaStreamCaptureStatusNone) coop = false; /* graph nodes: plain launch, see fused_grid_x */
  h->launches++;
  CU(yk_launch_icp_fused(dim3(gx, fn, h->S), q, coop, fp));
  return 1;
}

static IcpParams icp_params(const youth_cuda_handle* h, int level, const RingGeom& ring) {
  IcpParams ip;
  memset(&ip, 0, sizeof(ip));
  ip.maps = h->maps[level];
  ip.g = h->lv[level];
  ip.ring = ring;
  ip.npix = h->npix[level];
  ip.ppr = h->ppr[level];
  ip.nruns = h->nruns[level];
  ip.max_runs = h->max_runs;
  ip.dist2_thr = h->cfg.dist_thresh_m * h->cfg.dist_thresh_m;
  ip.cos_thr = h->cfg.cos_thresh;
  ip.pose_f = h->pose_f;
  ip.seq_count = h->seq_count;
  ip.partials = h->partials;
  ip.corr = NULL;
  ip.dbg_cur_slot = -1;
  ip.dbg_prev_slot = -1;
  ip.dbg_stream = 0;
  ip.tickets = h->tickets;
  ip.pose_d = h->pose_d;
  ip.pose_f_out = h->pose_f;
  ip.sums = h->sums;
  ip.pair_status = h->pair_status;
  ip.min_inliers = h->cfg.min_inliers;
  ip.do_solve = 1;
  ip.model = h->m.on ? h->m.maps[level] : NULL;
  ip.f0 = 0;
  ip.fn = ring.n;
  return ip;
}

/* stages 1 + 2 for frames [frame0, frame0 + cn) of every stream's group of n frames; raw_dev[s] points at
 * the first frame of the GROUP.  Chunks let host-fed groups overlap their H2D copies with ingest. */
static int enqueue_preprocess(youth_cuda_handle* h, const uint16_t* const* raw_dev, int n, int frame0, int cn) {
  const youth_cuda_config& c = h->cfg;
  const RingGeom ring = ring_of(h, n);
  const int frames = h->S * cn;
  /* stage 1 + 2a */
  {
    IngestParams ip;
    memset(&ip, 0, sizeof(ip));
    for (int s = 0; s < h->S; ++s) ip.raw[s] = raw_dev[s];
    for (int l = 0; l < c.levels; ++l) {
      ip.depth[l] = h->depth[l];
      ip.maps[l] = h->maps[l];
      ip.pyrcnt[l] = h->pyrcnt[l];
      ip.lv[l] = h->lv[l];
    }
    ip.pose_d = h->pose_d;
    ip.pose_f = h->pose_f;
    ip.pair_status = h->pair_status;
    ip.ring = ring;
    ip.frame0 = frame0;
    ip.chunk_n = cn;
    ip.levels = c.levels;
    ip.dmin = c.depth_min_mm;
    ip.dmax = c.depth_max_mm;
    ip.range_cut = h->range_cut;
    for (int dy = 0; dy < 7; ++dy)
      for (int k = 0; k < 8; ++k)
        ip.wsp[dy * 8 + k] = make_float2(k <= 6 ? h->ws[dy * 7 + k] : 0.0f, k >= 1 ? h->ws[dy * 7 + k - 1] : 0.0f);
    ip.wr = h->wr;
    ip.depth_factor = c.depth_factor;
    ip.pyr_thr = 3.0f * c.sigma_range_mm;
    dim3 grid((c.width + 63) / 64, (c.height + 15) / 16, frames); /* 64 x 16 pixel tiles (youth_ingest.cu) */
    ProfScope ps(h, YOUTH_PROF_INGEST);
    ip.wt = reinterpret_cast<const float4*>(h->wt);
    ip.r_df = h->r_df;
    for (int l = 0; l < c.levels; ++l) {
      ip.r_fx[l] = h->r_fx[l];
      ip.r_fy[l] = h->r_fy[l];
    }
    const int mode = !c.bilateral ? YK_INGEST_RAW
                                  : (h->range_cut + 2 <= YK_WT_STRIDE && !h->ingest_generic ? YK_INGEST_BILATERAL_WT : YK_INGEST_BILATERAL);
    yk_launch_ingest(mode, h->fast_div != 0, h->debug_maps, grid, h->stream, ip);
  }
  /* stage 2b */
  {
    NormalParams np;
    memset(&np, 0, sizeof(np));
    int total = 0;
    for (int l = 0; l < c.levels; ++l) {
      np.maps[l] = h->maps[l];
      np.lv[l] = h->lv[l];
      if (l >= 1) total += h->npix[l];
    }
    np.first_level = 1;
    np.ring = ring;
    np.frame0 = frame0;
    np.chunk_n = cn;
    np.levels = c.levels;
    if (total > 0) {
      dim3 grid((total + 255) / 256, frames);
      ProfScope ps(h, YOUTH_PROF_NORMALS);
      yk_launch_normals(h->fast_div != 0, grid, h->stream, np);
    }
  }
  CU(cudaGetLastError());
  return 1;
}

/* Host-fed groups: frames per tracking range (a multiple of the copy chunk), so that stages 3-5 of the
 * first range could iterate while later frames are still being copied.  Measured on B200 (300 frames from
 * pinned host memory, ranges of 160 / 100 frames): no gain -- what the overlap saves (H2D of the later
 * ranges, 55 GB/s also under load) the shorter launches lose in partial waves and per-launch tails -- so
 * the default is one range; YOUTH_HOST_RANGE=<frames> keeps the experiment reproducible. */
static int host_range_frames(const youth_cuda_handle* h, int n_frames, int chunk) {
  int r = h->host_range;
  if (r <= 0) return n_frames;
  r = ((r + chunk - 1) / chunk) * chunk;
  return r >= n_frames ? n_frames : r;
}

/* Frames per copy + preprocess chunk of a host-fed group.  A caller that waits for every group wants small
 * chunks (ingest of chunk c runs under the H2D of chunk c + 1, the tail after the last copy is one small
 * chunk).  A caller that keeps two groups in flight (ticket API) has this group's copies hidden under the
 * previous group's kernels anyway, and 19 ingest launches of 16 frames (4 800 CTAs = 4.05 waves each) waste
 * a fifth of their last wave: when the previous group is still running, a quarter of the group per chunk. */
static int host_chunk_frames(youth_cuda_handle* h, int n_frames, int k) {
  int ch = (n_frames + YK_MAX_CHUNKS - 1) / YK_MAX_CHUNKS;
  if (ch < YK_CHUNK_FRAMES) ch = YK_CHUNK_FRAMES;
  if (h->host_range <= 0 && h->raw_used[k ^ 1] && cudaEventQuery(h->raw_free[k ^ 1]) == cudaErrorNotReady) {
    const int big = (n_frames + 3) / 4;
    if (big > ch) ch = big;
  }
  return ch;
}

/* stages 3-5 for frames [f0, f0 + fn) of every stream's group of n frames: coarse to fine, fixed iteration
 * schedule, one launch per iteration, no host sync.  Pairs are independent, so a group may be tracked in
 * several frame ranges (host-fed groups: the first range iterates while later frames are still in flight);
 * the arithmetic of a pair does not depend on which launch carries it. */
static int enqueue_icp_range(youth_cuda_handle* h, int n, int f0, int fn) {
  const youth_cuda_config& c = h->cfg;
  const RingGeom ring = ring_of(h, n);
  const int G = h->icp_group;
  if (G > 0 && fn > G) {
    /* frame-range schedule (youth_cuda_set_icp_schedule): the whole coarse-to-fine schedule of G frames per
     * sequence before the next G are touched, ranges round-robin over side streams so that the tail of one
     * range's launch (cross-run reduction + solve) overlaps the sweep of another's.  Measured on B200: no
     * gain over one launch per iteration (profiles/README.md); kept as a tested knob.  Profiled steps use
     * one queue so that a launch's CUDA-event duration is that launch alone. */
    const int ngroups = (fn + G - 1) / G;
    int K = h->prof_on ? 1 : h->icp_nq;
    if (K > ngroups) K = ngroups;
    if (K > 1) {
      CU(cudaEventRecord(h->icp_fork, h->stream));
      for (int k = 0; k < K; ++k) CU(cudaStreamWaitEvent(h->icp_q[k], h->icp_fork, 0));
    }
    for (int g = 0; g < ngroups; ++g) {
      const int gn = (g + 1) * G <= fn ? G : fn - g * G;
      if (const int gx = fused_grid_x(h, gn)) {
        if (!launch_icp_fused(h, ring, gx, f0 + g * G, gn, K > 1 ? h->icp_q[g % K] : h->stream)) return 0;
        continue;
      }
      for (int level = c.levels - 1; level >= 0; --level) {
        IcpParams ip = icp_params(h, level, ring);
        ip.f0 = f0 + g * G;
        ip.fn = gn;
        for (int it = 0; it < c.iters[level]; ++it) {
          if (K > 1) {
            h->launches++;
            launch_icp_on(h, ip, gn, level, h->icp_q[g % K]);
          } else {
            ProfScope ps(h, YOUTH_PROF_ICP0 + level);
            launch_icp_on(h, ip, gn, level, h->stream);
          }
        }
      }
      if (h->prof_on && h->prof_n > h->prof_cap - 256 && !prof_flush(h)) return 0;
    }
    if (K > 1)
      for (int k = 0; k < K; ++k) {
        CU(cudaEventRecord(h->icp_join[k], h->icp_q[k]));
        CU(cudaStreamWaitEvent(h->stream, h->icp_join[k], 0));
      }
  } else if (const int gx = fused_grid_x(h, fn)) {
    if (!launch_icp_fused(h, ring, gx, f0, fn, h->stream)) return 0;
  } else {
    for (int level = c.levels - 1; level >= 0; --level) {
      IcpParams ip = icp_params(h, level, ring);
      ip.f0 = f0;
      ip.fn = fn;
      for (int it = 0; it < c.iters[level]; ++it) {
        ProfScope ps(h, YOUTH_PROF_ICP0 + level);
        launch_icp_on(h, ip, fn, level, h->stream);
      }
    }
  }
  CU(cudaGetLastError());
  if (h->prof_on && h->prof_n > h->prof_cap - 256) return prof_flush(h);
  return 1;
}

/* pose chain + trajectory append for the whole group (after every range of stages 3-5) */
static int enqueue_compose(youth_cuda_handle* h, int n) {
  const RingGeom ring = ring_of(h, n);
  ComposeParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.ring = ring;
  cp.seq_count = h->seq_count;
  cp.world = h->world;
  cp.pose_d = h->pose_d;
  cp.sums = h->sums;
  cp.pair_status = h->pair_status;
  cp.traj = h->traj;
  cp.traj_status = h->traj_status;
  cp.last_inliers = h->last_inliers;
  cp.head = h->d_head;
  cp.cap = h->cfg.traj_capacity;
  cp.world_f = h->m.on ? h->m.world_f : NULL;
  cp.last_status = h->m.on ? h->m.last_status : NULL;
  {
    ProfScope ps(h, YOUTH_PROF_MISC);
    yk_launch_compose(h->S, h->stream, cp);
  }
  CU(cudaGetLastError());
  return 1;
}

/* stages 3-5 + pose chain for a whole group of n frames per stream */
static int enqueue_icp(youth_cuda_handle* h, int n) { return enqueue_icp_range(h, n, 0, n) && enqueue_compose(h, n); }

/* ------------------------------------------------------------------ frame-to-model (include/youth_model.h) */

extern "C" int youth_tsdf_default_config(youth_tsdf_config* t) {
  if (!t) return fail("null config");
  memset(t, 0, sizeof(*t));
  t->dim[0] = 256;
  t->dim[1] = 128;
  t->dim[2] = 256;
  t->voxel_m = 0.025f;
  t->origin[0] = -3.2f;
  t->origin[1] = -1.6f;
  t->origin[2] = -1.2f;
  t->trunc_m = 0.1f;
  t->max_weight = 64;
  t->near_m = 0.4f;
  t->far_m = 8.0f;
  return 1;
}

static size_t model_voxels(const youth_cuda_handle* h) {
  return (size_t)h->m.geom.dx * h->m.geom.dy * h->m.geom.dz;
}

/* fresh volume (tsdf 1, weight 0) and empty model maps for sequence `stream` (-1 = all) */
static int model_clear(youth_cuda_handle* h, int stream) {
  const size_t nv = model_voxels(h);
  const int s0 = stream < 0 ? 0 : stream, s1 = stream < 0 ? h->S : stream + 1;
  k_fill_u32<<<1184, 256, 0, h->stream>>>(reinterpret_cast<uint32_t*>(h->m.vol + (size_t)s0 * nv), (size_t)(s1 - s0) * nv,
                                          0x00007FFFu); /* short2 (32767, 0) */
  h->launches++;
  CU(cudaGetLastError());
  return 1;
}

extern "C" int youth_cuda_model_enabled(const youth_cuda_handle* h) { return h && h->m.on ? 1 : 0; }

extern "C" int youth_cuda_enable_model(youth_cuda_handle* h, const youth_tsdf_config* t) {
  if (!h || !t) return fail("null argument");
  if (h->m.on) return fail("frame-to-model tracking is already enabled");
  if (h->total != 0) return fail("enable the model before the first frame (or after a full reset)");
  for (int a = 0; a < 3; ++a)
    if (t->dim[a] < 8 || t->dim[a] > 1024) return fail("tsdf dim[%d] must be 8..1024", a);
  if (t->dim[0] % 8) return fail("tsdf dim[0] must be a multiple of 8");
  if (!(t->voxel_m > 0.f) || !(t->trunc_m > 0.f)) return fail("voxel_m and trunc_m must be > 0");
  if (t->max_weight < 1 || t->max_weight > 32767) return fail("max_weight must be 1..32767");
  if (!(t->near_m > 0.f) || !(t->far_m > t->near_m)) return fail("need 0 < near_m < far_m");
  CU(cudaSetDevice(h->cfg.device));
  if (!ensure_debug_maps(h)) return 0; /* the ray cast starts its march from the fused frame's depth pyramid */
  h->m.cfg = *t;
  TsdfGeom& g = h->m.geom;
  g.dx = t->dim[0];
  g.dy = t->dim[1];
  g.dz = t->dim[2];
  g.vs = t->voxel_m;
  g.inv_vs = 1.0f / t->voxel_m;
  g.ox = t->origin[0];
  g.oy = t->origin[1];
  g.oz = t->origin[2];
  g.mu = t->trunc_m;
  g.step = t->trunc_m * 0.8f;
  g.maxw = t->max_weight;
  g.near_m = t->near_m;
  g.far_m = t->far_m;
  CU(dalloc(&h->m.vol, (size_t)h->S * model_voxels(h)));
  for (int l = 0; l < h->cfg.levels; ++l) CU(dalloc(&h->m.maps[l], (size_t)h->S * 3 * h->npix[l]));
  CU(dalloc(&h->m.world_f, (size_t)h->S * 12));
  CU(dalloc(&h->m.last_status, (size_t)h->S));
  h->m.on = true;
  if (!model_clear(h, -1)) return 0;
  CU(cudaStreamSynchronize(h->stream));
  return 1;
}

/* fusion of the newest frame (slot < 0) or of ring slot `slot`, sequences [s0, s0 + ns) */
static int enqueue_integrate(youth_cuda_handle* h, int s0, int ns, int slot, bool honour_status) {
  IntegrateParams p;
  memset(&p, 0, sizeof(p));
  p.vol = h->m.vol;
  p.t = h->m.geom;
  p.maps0 = h->maps[0];
  p.g = h->lv[0];
  p.ring = ring_of(h, 1);
  p.world_f = h->m.world_f;
  p.last_status = honour_status ? h->m.last_status : NULL;
  p.depth_factor = h->cfg.depth_factor;
  p.slot = slot;
  p.stream0 = s0;
  const int zchunks = (p.t.dz + YM_ZCHUNK - 1) / YM_ZCHUNK;
  const dim3 grid((p.t.dx + 31) / 32, (p.t.dy + 7) / 8, (unsigned)(ns * zchunks));
  ProfScope ps(h, YOUTH_PROF_INTEGRATE);
  k_tsdf_integrate<<<grid, 256, 0, h->stream>>>(p);
  CU(cudaGetLastError());
  return 1;
}

/* hint_slot: -2 = no march-start hint, -1 = the newest frame's depth pyramid, >= 0 = that ring slot */
static int enqueue_raycast(youth_cuda_handle* h, int s0, int ns, int hint_slot) {
  RaycastParams p;
  memset(&p, 0, sizeof(p));
  p.vol = h->m.vol;
  p.t = h->m.geom;
  int total = 0;
  for (int l = 0; l < h->cfg.levels; ++l) {
    p.model[l] = h->m.maps[l];
    p.lv[l] = h->lv[l];
    total += (h->npix[l] + 31) & ~31; /* every level padded to whole warps */
  }
  p.levels = h->cfg.levels;
  p.world_f = h->m.world_f;
  p.stream0 = s0;
  for (int l = 0; l < h->cfg.levels; ++l) p.depth[l] = h->depth[l];
  p.ring = ring_of(h, 1);
  p.hint = hint_slot >= -1;
  p.hint_slot = hint_slot;
  p.depth_factor = h->cfg.depth_factor;
  const dim3 grid((total + 255) / 256, (unsigned)ns);
  {
    ProfScope ps(h, YOUTH_PROF_RAYCAST);
    k_tsdf_raycast<<<grid, 256, 0, h->stream>>>(p);
  }
  /* model normals = stage 2b on the ray-cast vertex maps (every level): the model maps are a one-slot ring */
  {
    NormalParams np;
    memset(&np, 0, sizeof(np));
    int px = 0;
    for (int l = 0; l < h->cfg.levels; ++l) {
      np.maps[l] = h->m.maps[l] + (size_t)s0 * 3 * h->npix[l];
      np.lv[l] = h->lv[l];
      px += h->npix[l];
    }
    np.first_level = 0;
    np.ring = ring_of(h, 1);
    np.ring.R = 1;
    np.ring.S = ns;
    np.frame0 = 0;
    np.chunk_n = 1;
    np.levels = h->cfg.levels;
    const dim3 ngrid((px + 255) / 256, (unsigned)ns);
    ProfScope ps(h, YOUTH_PROF_RAYCAST);
    yk_launch_normals(h->fast_div != 0, ngrid, h->stream, np);
  }
  CU(cudaGetLastError());
  return 1;
}

/* one frame of every sequence: stages 1-5 against the model, fusion, ray cast */
static int enqueue_model_frame(youth_cuda_handle* h, const uint16_t* const* raw_dev) {
  return enqueue_preprocess(h, raw_dev, 1, 0, 1) && enqueue_icp(h, 1) && enqueue_integrate(h, 0, h->S, -1, true) &&
         enqueue_raycast(h, 0, h->S, -1);
}

static int finish_group(youth_cuda_handle* h, int n_frames, const uint32_t* timestamps_ms, float* poses_out);

/* Frame-to-model groups are chains: frame after frame, each one H2D/D2D copy into a fixed landing
 * slot followed by the captured graph of the whole per-frame schedule. */
static int track_batch_model(youth_cuda_handle* h, const uint16_t* const* depth, int n_frames, int mem_kind,
                             const uint32_t* timestamps_ms, float* poses_out) {
  if (mem_kind != YOUTH_MEM_HOST && mem_kind != YOUTH_MEM_DEVICE && mem_kind != YOUTH_MEM_HOST_PINNED)
    return fail("unknown mem_kind %d", mem_kind);
  const size_t frame_px = (size_t)h->cfg.width * h->cfg.height, frame_bytes = frame_px * sizeof(uint16_t);
  const uint16_t* dev_ptrs[YK_MAX_STREAMS];
  for (int s = 0; s < h->S; ++s) dev_ptrs[s] = h->raw[0] + (size_t)s * frame_px;
  const cudaMemcpyKind kind = mem_kind == YOUTH_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int i = 0; i < n_frames; ++i) {
    for (int s = 0; s < h->S; ++s) /* pageable sources are staged by the runtime before the call returns */
      CU(cudaMemcpyAsync(h->raw[0] + (size_t)s * frame_px, depth[s] + (size_t)i * frame_px, frame_bytes, kind, h->stream));
    if (h->graphs_enabled && !h->prof_on) {
      if (!h->m.graph) {
        cudaGraph_t graph = NULL;
        const uint64_t before = h->launches;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int ok = enqueue_model_frame(h, dev_ptrs);
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (!ok || ce != cudaSuccess || !graph) {
          if (graph) cudaGraphDestroy(graph);
          return fail("CUDA graph capture failed: %s", ce != cudaSuccess ? cudaGetErrorString(ce) : youth_cuda_last_error());
        }
        h->m.graph_launches = h->launches - before;
        h->launches = before;
        const cudaError_t ie = cudaGraphInstantiate(&h->m.graph, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
      }
      CU(cudaGraphLaunch(h->m.graph, h->stream));
      h->launches += h->m.graph_launches;
    } else if (!enqueue_model_frame(h, dev_ptrs)) {
      return 0;
    }
  }
  return finish_group(h, n_frames, timestamps_ms, poses_out);
}

/* host bookkeeping of a group that has been enqueued, and the optional blocking pose read-back */
static int finish_group(youth_cuda_handle* h, int n_frames, const uint32_t* timestamps_ms, float* poses_out) {
  for (int s = 0; s < h->S; ++s) {
    for (int i = 0; i < n_frames; ++i)
      h->h_ts[(size_t)s * h->cfg.traj_capacity + h->h_count[s] + i] =
          timestamps_ms ? timestamps_ms[i] : (uint32_t)(((long long)(h->h_count[s] + i) * 100) / 3);
    h->h_count[s] += n_frames;
  }
  h->total += n_frames;
  if (poses_out) {
    for (int s = 0; s < h->S; ++s) {
      const float* src = h->traj + ((size_t)s * h->cfg.traj_capacity + (h->h_count[s] - n_frames)) * 12;
      CU(cudaMemcpyAsync(poses_out + (size_t)s * n_frames * 12, src, sizeof(float) * 12 * n_frames,
                         cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
  }
  return 1;
}

extern "C" int youth_cuda_track_batch(youth_cuda_handle* h, const uint16_t* const* depth, int n_frames,
                                      int mem_kind, const uint32_t* timestamps_ms, float* poses_out) {
  if (!h || !depth) return fail("null argument");
  if (n_frames < 1 || n_frames > h->B) return fail("n_frames must be 1..batch (%d)", h->B);
  for (int s = 0; s < h->S; ++s) {
    if (!depth[s]) return fail("depth[%d] is NULL", s);
    if (h->h_count[s] + n_frames > h->cfg.traj_capacity) return fail("trajectory capacity (%d) exceeded", h->cfg.traj_capacity);
  }
  CU(cudaSetDevice(h->cfg.device));
  if (h->m.on) return track_batch_model(h, depth, n_frames, mem_kind, timestamps_ms, poses_out);
  const size_t frame_px = (size_t)h->cfg.width * h->cfg.height;
  const size_t seq_bytes = frame_px * sizeof(uint16_t) * (size_t)n_frames;
  const uint16_t* dev_ptrs[YK_MAX_STREAMS];
  if (mem_kind == YOUTH_MEM_DEVICE) {
    for (int s = 0; s < h->S; ++s) dev_ptrs[s] = depth[s];
    if (!enqueue_preprocess(h, dev_ptrs, n_frames, 0, n_frames)) return 0;
    if (!enqueue_icp(h, n_frames)) return 0;
  } else if (mem_kind == YOUTH_MEM_HOST || mem_kind == YOUTH_MEM_HOST_PINNED) {
    const int k = h->raw_turn;
    h->raw_turn ^= 1;
    /* the landing zone may still be read by the group launched two calls ago */
    if (mem_kind == YOUTH_MEM_HOST && !ensure_staging(h)) return 0;
    if (h->raw_used[k]) {
      if (mem_kind == YOUTH_MEM_HOST)
        CU(cudaEventSynchronize(h->raw_free[k])); /* the pinned staging copy is about to be overwritten */
      CU(cudaStreamWaitEvent(h->copy_stream, h->raw_free[k], 0));
    }
    if (h->graphs_enabled && !h->prof_on && n_frames <= YK_GRAPH_MAX_N) {
      /* launch-bound small group: one H2D per stream on the compute stream, then the whole kernel
       * schedule as ONE captured graph (valid from call to call: ring head and counters live on the device) */
      for (int s = 0; s < h->S; ++s) {
        uint16_t* dst = h->raw[k] + (size_t)s * n_frames * frame_px;
        const uint16_t* src = depth[s];
        if (mem_kind == YOUTH_MEM_HOST) {
          uint16_t* stage = h->pinned[k] + (size_t)s * n_frames * frame_px;
          stage_copy(stage, src, seq_bytes);
          src = stage;
        }
        CU(cudaMemcpyAsync(dst, src, seq_bytes, cudaMemcpyHostToDevice, h->stream));
        dev_ptrs[s] = dst;
      }
      const int gi = (n_frames - 1) * 2 + k;
      if (!h->graphs[gi].exec) {
        cudaGraph_t graph = NULL;
        const uint64_t before = h->launches;
        CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int ok = enqueue_preprocess(h, dev_ptrs, n_frames, 0, n_frames) && enqueue_icp(h, n_frames);
        cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (!ok || ce != cudaSuccess || !graph) {
          if (graph) cudaGraphDestroy(graph);
          return fail("CUDA graph capture failed: %s", ce != cudaSuccess ? cudaGetErrorString(ce) : youth_cuda_last_error());
        }
        h->graphs[gi].launches = h->launches - before;
        h->launches = before;
        ce = cudaGraphInstantiate(&h->graphs[gi].exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        h->graphs[gi].n = n_frames;
        h->graphs[gi].k = k;
      }
      CU(cudaGraphLaunch(h->graphs[gi].exec, h->stream));
      h->launches += h->graphs[gi].launches;
    } else {
      /* copy and preprocess in chunks: the H2D of chunk c+1 overlaps ingest of chunk c; and track in
       * frame ranges: stages 3-5 of range r iterate while the copy stream brings in range r+1 */
      const int ch = host_chunk_frames(h, n_frames, k);
      const int range = host_range_frames(h, n_frames, ch);
      for (int s = 0; s < h->S; ++s) dev_ptrs[s] = h->raw[k] + (size_t)s * n_frames * frame_px;
      int ci = 0, tracked = 0;
      for (int f0 = 0; f0 < n_frames; f0 += ch, ++ci) {
        const int cn = n_frames - f0 < ch ? n_frames - f0 : ch;
        const size_t off = (size_t)f0 * frame_px, bytes = (size_t)cn * frame_px * sizeof(uint16_t);
        for (int s = 0; s < h->S; ++s) {
          uint16_t* dst = h->raw[k] + (size_t)s * n_frames * frame_px + off;
          const uint16_t* src = depth[s] + off;
          if (mem_kind == YOUTH_MEM_HOST) {
            uint16_t* stage = h->pinned[k] + (size_t)s * n_frames * frame_px + off;
            stage_copy(stage, src, bytes); /* synchronous copy: the caller may reuse its buffer (SLAM.cpp:133-134) */
            src = stage;
          }
          CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
        }
        CU(cudaEventRecord(h->chunk_ready[k][ci], h->copy_stream));
        CU(cudaStreamWaitEvent(h->stream, h->chunk_ready[k][ci], 0));
        if (!enqueue_preprocess(h, dev_ptrs, n_frames, f0, cn)) return 0;
        const int have = f0 + cn;
        if (have - tracked >= range && n_frames - have >= range / 2) { /* no short last range */
          if (!enqueue_icp_range(h, n_frames, tracked, have - tracked)) return 0;
          tracked = have;
        }
      }
      if (tracked < n_frames && !enqueue_icp_range(h, n_frames, tracked, n_frames - tracked)) return 0;
      if (!enqueue_compose(h, n_frames)) return 0;
    }
    CU(cudaEventRecord(h->raw_free[k], h->stream));
    h->raw_used[k] = true;
  } else {
    return fail("unknown mem_kind %d", mem_kind);
  }
  return finish_group(h, n_frames, timestamps_ms, poses_out);
}

/* A group whose frames arrive as YD16 streams: the packed bytes cross PCIe (about a third of the raw
 * frames), k_yd16_decode unpacks them straight into the raw landing zone, then the normal schedule.
 * Chunked like the raw host path: the H2D of chunk c+1 overlaps decode + ingest of chunk c. */
extern "C" int youth_cuda_track_batch_packed(youth_cuda_handle* h, const uint8_t* const* streams,
                                             const uint64_t* const* offsets, int n_frames, int mem_kind,
                                             const uint32_t* timestamps_ms, float* poses_out) {
  if (!h || !streams || !offsets) return fail("null argument");
  if (n_frames < 1 || n_frames > h->B) return fail("n_frames must be 1..batch (%d)", h->B);
  if (mem_kind != YOUTH_MEM_HOST && mem_kind != YOUTH_MEM_HOST_PINNED) return fail("packed input must be host memory");
  for (int s = 0; s < h->S; ++s) {
    if (!streams[s] || !offsets[s]) return fail("streams[%d] / offsets[%d] is NULL", s, s);
    if (h->h_count[s] + n_frames > h->cfg.traj_capacity) return fail("trajectory capacity (%d) exceeded", h->cfg.traj_capacity);
  }
  CU(cudaSetDevice(h->cfg.device));
  if (!h->codec) {
    if (!youth_codec_create(h->cfg.width, h->cfg.height, h->P, h->cfg.device, &h->codec)) return 0;
    for (int k = 0; k < 2; ++k) {
      CU(cudaHostAlloc((void**)&h->pk_h_off[k], sizeof(unsigned long long) * ((size_t)h->P + 1), cudaHostAllocDefault));
      CU(dalloc(&h->pk_d_off[k], (size_t)h->P + 1));
    }
    CU(cudaEventCreateWithFlags(&h->pk_free, cudaEventDisableTiming));
  }
  youth_codec* c = h->codec;
  for (int s = 0; s < h->S; ++s)
    for (int i = 0; i < n_frames; ++i) {
      if (offsets[s][i + 1] < offsets[s][i]) return fail("offsets must be non-decreasing");
      if (!codec_check_header(c, streams[s] + offsets[s][i], offsets[s][i + 1] - offsets[s][i], s * n_frames + i)) return 0;
    }
  if (h->m.on) {
    /* frame-to-model: a chain of single frames.  Unpack the whole group into the codec's raw buffer on the
     * tracking stream, then run the chain from device memory. */
    CU(cudaStreamSynchronize(h->stream)); /* the previous group no longer reads the staging buffers */
    unsigned long long* O = h->pk_h_off[0];
    unsigned long long base = 0;
    for (int s = 0; s < h->S; ++s) {
      for (int i = 0; i < n_frames; ++i) O[(size_t)s * n_frames + i] = base + (offsets[s][i] - offsets[s][0]);
      base += offsets[s][n_frames] - offsets[s][0];
    }
    O[(size_t)h->S * n_frames] = base;
    if (base > (unsigned long long)h->P * c->stride) return fail("packed group larger than the context");
    const size_t frame_px = (size_t)h->cfg.width * h->cfg.height;
    CU(cudaMemcpyAsync(h->pk_d_off[0], O, sizeof(unsigned long long) * ((size_t)h->S * n_frames + 1), cudaMemcpyHostToDevice,
                       h->stream));
    for (int s = 0; s < h->S; ++s)
      CU(cudaMemcpyAsync(c->d_packed + O[(size_t)s * n_frames], streams[s] + offsets[s][0],
                         (size_t)(offsets[s][n_frames] - offsets[s][0]), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), h->stream));
    h->launches += 2;
    if (!codec_enqueue_decode(c, h->stream, c->d_packed, h->pk_d_off[0], c->d_tile_sum, h->S * n_frames, c->d_raw)) return 0;
    CU(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    const uint16_t* dev[YK_MAX_STREAMS];
    for (int s = 0; s < h->S; ++s) dev[s] = c->d_raw + (size_t)s * n_frames * frame_px;
    if (!track_batch_model(h, dev, n_frames, YOUTH_MEM_DEVICE, timestamps_ms, poses_out)) return 0;
    if (poses_out && *c->h_err) return fail("malformed YD16 stream (device check 0x%x); reset the tracker", *c->h_err);
    return 1;
  }
  const int k = h->raw_turn;
  h->raw_turn ^= 1;
  if (h->raw_used[k]) CU(cudaEventSynchronize(h->raw_free[k])); /* pk_h_off[k] and raw[k] of two calls ago are free */
  if (h->pk_used) CU(cudaStreamWaitEvent(h->copy_stream, h->pk_free, 0)); /* d_packed of the previous group was read */
  /* device placement of the streams: sequence after sequence, byte granular */
  unsigned long long* O = h->pk_h_off[k];
  unsigned long long base = 0;
  for (int s = 0; s < h->S; ++s) {
    for (int i = 0; i < n_frames; ++i) O[(size_t)s * n_frames + i] = base + (offsets[s][i] - offsets[s][0]);
    base += offsets[s][n_frames] - offsets[s][0];
  }
  O[(size_t)h->S * n_frames] = base;
  if (base > (unsigned long long)h->P * c->stride) return fail("packed group larger than the context");
  CU(cudaMemcpyAsync(h->pk_d_off[k], O, sizeof(unsigned long long) * ((size_t)h->S * n_frames + 1), cudaMemcpyHostToDevice,
                     h->copy_stream));
  CU(cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), h->copy_stream));
  const size_t frame_px = (size_t)h->cfg.width * h->cfg.height;
  const uint16_t* dev_ptrs[YK_MAX_STREAMS];
  for (int s = 0; s < h->S; ++s) dev_ptrs[s] = h->raw[k] + (size_t)s * n_frames * frame_px;
  const int ch = host_chunk_frames(h, n_frames, k);
  const int range = host_range_frames(h, n_frames, ch);
  int ci = 0, tracked = 0;
  for (int f0 = 0; f0 < n_frames; f0 += ch, ++ci) {
    const int cn = n_frames - f0 < ch ? n_frames - f0 : ch;
    for (int s = 0; s < h->S; ++s) {
      const size_t j0 = (size_t)s * n_frames + f0;
      CU(cudaMemcpyAsync(c->d_packed + O[j0], streams[s] + offsets[s][f0], (size_t)(O[j0 + cn] - O[j0]),
                         cudaMemcpyHostToDevice, h->copy_stream));
    }
    CU(cudaEventRecord(h->chunk_ready[k][ci], h->copy_stream));
    CU(cudaStreamWaitEvent(h->stream, h->chunk_ready[k][ci], 0));
    for (int s = 0; s < h->S; ++s) {
      const size_t j0 = (size_t)s * n_frames + f0;
      ProfScope ps(h, YOUTH_PROF_MISC);
      h->launches++; /* two kernels per decode */
      if (!codec_enqueue_decode(c, h->stream, c->d_packed, h->pk_d_off[k] + j0, c->d_tile_sum + j0 * c->tiles, cn,
                                h->raw[k] + j0 * frame_px))
        return 0;
    }
    if (!enqueue_preprocess(h, dev_ptrs, n_frames, f0, cn)) return 0;
    const int have = f0 + cn;
    if (have - tracked >= range && n_frames - have >= range / 2) {
      if (!enqueue_icp_range(h, n_frames, tracked, have - tracked)) return 0;
      tracked = have;
    }
  }
  CU(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaEventRecord(h->pk_free, h->stream));
  h->pk_used = true;
  if (tracked < n_frames && !enqueue_icp_range(h, n_frames, tracked, n_frames - tracked)) return 0;
  if (!enqueue_compose(h, n_frames)) return 0;
  CU(cudaEventRecord(h->raw_free[k], h->stream));
  h->raw_used[k] = true;
  if (!finish_group(h, n_frames, timestamps_ms, poses_out)) return 0;
  if (poses_out && *c->h_err) return fail("malformed YD16 stream (device check 0x%x); reset the tracker", *c->h_err);
  return 1;
}

extern "C" int youth_cuda_track(youth_cuda_handle* h, const uint16_t* depth_mm, uint32_t timestamp_ms,
                                float pose_out[12]) {
  if (!h) return fail("null handle");
  if (h->S != 1) return fail("youth_cuda_track needs a single-sequence handle (n_streams == 1)");
  const uint16_t* d[1] = {depth_mm};
  return youth_cuda_track_batch(h, d, 1, YOUTH_MEM_HOST, &timestamp_ms, pose_out);
}

extern "C" int youth_cuda_sync(youth_cuda_handle* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->copy_stream));
  CU(cudaStreamSynchronize(h->stream));
  if (h->codec && *h->codec->h_err) {
    const unsigned int e = *h->codec->h_err;
    *h->codec->h_err = 0;
    return fail("malformed YD16 stream in an earlier packed group (device check 0x%x); reset the tracker", e);
  }
  return 1;
}

extern "C" int youth_cuda_reset(youth_cuda_handle* h, int stream) {
  if (!h) return fail("null handle");
  if (stream < -1 || stream >= h->S) return fail("stream out of range");
  CU(cudaSetDevice(h->cfg.device));
  if (stream < 0) {
    CU(cudaMemsetAsync(h->seq_count, 0, sizeof(int) * h->S, h->stream));
    CU(cudaMemsetAsync(h->last_inliers, 0, sizeof(int) * h->S, h->stream));
    CU(cudaMemsetAsync(h->d_head, 0, sizeof(int), h->stream));
    memset(h->h_count, 0, sizeof(int) * h->S);
    h->total = 0; /* stream-ordered: later groups see the cleared counters */
  } else {
    CU(cudaMemsetAsync(h->seq_count + stream, 0, sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->last_inliers + stream, 0, sizeof(int), h->stream));
    h->h_count[stream] = 0;
  }
  if (h->m.on && !model_clear(h, stream)) return 0;
  return 1;
}

extern "C" int youth_cuda_frame_count(youth_cuda_handle* h, int stream) {
  if (!h || stream < 0 || stream >= h->S) return -1;
  return h->h_count[stream];
}

extern "C" int youth_cuda_get_trajectory(youth_cuda_handle* h, int stream, int first, int max_frames,
                                         float* poses_out, uint32_t* timestamps_out, uint32_t* status_out) {
  if (!h || stream < 0 || stream >= h->S || first < 0 || max_frames < 0) {
    fail("bad argument");
    return -1;
  }
  int n = h->h_count[stream] - first;
  if (n > max_frames) n = max_frames;
  if (n <= 0) return 0;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return -1;
  const size_t base = (size_t)stream * h->cfg.traj_capacity + first;
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && poses_out)
    e = cudaMemcpy(poses_out, h->traj + base * 12, sizeof(float) * 12 * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && status_out)
    e = cudaMemcpy(status_out, h->traj_status + base, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    fail("trajectory read-back failed: %s", cudaGetErrorString(e));
    return -1;
  }
  if (timestamps_out) memcpy(timestamps_out, h->h_ts + base, sizeof(uint32_t) * n);
  return n;
}

/* Stream-ordered read-back: the copies are enqueued behind everything submitted so far and the call
 * returns; the ticket names an event recorded behind them.  With it a caller keeps two groups in flight
 * -- submit group g+1 (its H2D runs on the copy stream under group g's kernels), then collect group g. */
extern "C" int youth_cuda_read_trajectory_async(youth_cuda_handle* h, int stream, int first, int max_frames,
                                                float* poses_out, uint32_t* status_out, uint64_t* ticket_out) {
  if (!h || stream < 0 || stream >= h->S || first < 0 || max_frames < 0 || !ticket_out) {
    fail("bad argument");
    return -1;
  }
  int n = h->h_count[stream] - first;
  if (n > max_frames) n = max_frames;
  if (n < 0) n = 0;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return -1;
  const size_t base = (size_t)stream * h->cfg.traj_capacity + first;
  cudaError_t e = cudaSuccess;
  if (n > 0 && poses_out)
    e = cudaMemcpyAsync(poses_out, h->traj + base * 12, sizeof(float) * 12 * n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && n > 0 && status_out)
    e = cudaMemcpyAsync(status_out, h->traj_status + base, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, h->stream);
  const uint64_t t = h->ticket_next;
  cudaEvent_t& ev = h->ticket_ev[t % YK_TICKETS];
  if (e == cudaSuccess && !ev) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventRecord(ev, h->stream);
  if (e != cudaSuccess) {
    fail("asynchronous trajectory read-back failed: %s", cudaGetErrorString(e));
    return -1;
  }
  h->ticket_next = t + 1;
  *ticket_out = t;
  return n;
}

extern "C" int youth_cuda_read_last_inliers_async(youth_cuda_handle* h, int stream, int* inliers_out) {
  if (!h || stream < 0 || stream >= h->S || !inliers_out) return fail("bad argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(inliers_out, h->last_inliers + stream, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return 1;
}

extern "C" int youth_cuda_wait_ticket(youth_cuda_handle* h, uint64_t ticket) {
  if (!h) return fail("null handle");
  if (ticket >= h->ticket_next) return fail("unknown ticket");
  /* When more than YK_TICKETS reads were issued since, the slot holds the event of a LATER read.  Every read is
   * recorded on the same stream, so that event completes after this ticket's copies: waiting for it covers the
   * ticket (never return without a wait -- the caller is about to read the destination buffers). */
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaEventSynchronize(h->ticket_ev[ticket % YK_TICKETS]));
  return 1;
}

extern "C" int youth_cuda_last_inliers(youth_cuda_handle* h, int stream) {
  if (!h || stream < 0 || stream >= h->S) return -1;
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return -1;
  int v = 0;
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
  if (cudaMemcpy(&v, h->last_inliers + stream, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return v;
}

extern "C" void* youth_cuda_trajectory_device_ptr(youth_cuda_handle* h, int stream) {
  if (!h || stream < 0 || stream >= h->S) return NULL;
  return h->traj + (size_t)stream * h->cfg.traj_capacity * 12;
}

extern "C" void* youth_cuda_host_alloc(size_t bytes) {
  void* p = NULL;
  /* portable: page-locked for every device of the process (a multi-GPU host pins once, any handle may read it) */
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    fail("cudaHostAlloc(%zu) failed", bytes);
    return NULL;
  }
  return p;
}

extern "C" void youth_cuda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

/* Device helpers for C hosts that do not link the CUDA runtime themselves (the native multi-GPU harness hands
 * these buffers to NCCL). */
extern "C" int youth_cuda_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" int youth_cuda_set_device(int device) {
  CU(cudaSetDevice(device));
  return 1;
}

/* the stream the handle launches on: work enqueued there by the caller (a collective on the trajectories) is
 * ordered after everything submitted so far, with no host synchronisation */
extern "C" void* youth_cuda_stream(youth_cuda_handle* h) { return h ? (void*)h->stream : NULL; }

extern "C" void* youth_cuda_device_alloc(size_t bytes) {
  void* p = NULL;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail("cudaMalloc of %zu bytes failed", bytes);
    return NULL;
  }
  return p;
}

extern "C" void youth_cuda_device_free(void* p) {
  if (p) cudaFree(p);
}

extern "C" int youth_cuda_copy_to_host(void* dst, const void* src_device, size_t bytes) {
  CU(cudaMemcpy(dst, src_device, bytes, cudaMemcpyDeviceToHost));
  return 1;
}

extern "C" int youth_cuda_copy_to_device(void* dst_device, const void* src, size_t bytes) {
  CU(cudaMemcpy(dst_device, src, bytes, cudaMemcpyHostToDevice));
  return 1;
}

extern "C" int youth_cuda_device_sync(void) {
  CU(cudaDeviceSynchronize());
  return 1;
}

/* ------------------------------------------------------------------ parity hooks */

static int slot_of_frame(const youth_cuda_handle* h, int frame) {
  if (frame < 0 || frame >= h->total || frame < h->total - h->R) return -1;
  return (int)(frame % h->R);
}

/* the device keeps three float2 planes per map set; the read-back presents them as the float4
 * (x, y, z, valid) maps / validity mask of the specification */
static int read_planes(const float2* planes, size_t np, int what, void* dst, size_t dst_bytes) {
  if (dst_bytes < (what == YOUTH_DBG_MASK ? np : np * 16)) return fail("dst too small");
  float* tmp = (float*)malloc(np * 6 * sizeof(float));
  if (!tmp) return fail("host allocation failed");
  cudaError_t e = cudaMemcpy(tmp, planes, np * 6 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    free(tmp);
    return fail("map read-back failed: %s", cudaGetErrorString(e));
  }
  for (size_t i = 0; i < np; ++i) {
    const float vx = tmp[2 * i], vy = tmp[2 * i + 1], vz = tmp[2 * np + 2 * i];
    const float nx = tmp[2 * np + 2 * i + 1], ny = tmp[4 * np + 2 * i], nz = tmp[4 * np + 2 * i + 1];
    const int vok = vz > 0.0f, nok = YK_N_VALID(nx);
    if (what == YOUTH_DBG_MASK) {
      ((uint8_t*)dst)[i] = (uint8_t)(vok | (nok << 1));
    } else {
      float* o = (float*)dst + 4 * i;
      if (what == YOUTH_DBG_VERTEX) {
        o[0] = vx; o[1] = vy; o[2] = vz; o[3] = vok ? 1.0f : 0.0f;
      } else {
        o[0] = nok ? nx : 0.0f; o[1] = ny; o[2] = nz; o[3] = nok ? 1.0f : 0.0f;
      }
    }
  }
  free(tmp);
  return 1;
}

extern "C" long long youth_cuda_model_surface_voxels(youth_cuda_handle* h, int stream) {
  if (!h || !h->m.on || stream < 0 || stream >= h->S) {
    fail("frame-to-model tracking is not enabled / stream out of range");
    return -1;
  }
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return -1;
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(h->sums); /* pair scratch, free between groups */
  unsigned long long cnt = 0;
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
  if (cudaMemsetAsync(d_cnt, 0, sizeof(cnt), h->stream) != cudaSuccess) return -1;
  const TsdfGeom& g = h->m.geom;
  k_tsdf_surface_count<<<1184, 256, 0, h->stream>>>(h->m.vol + (size_t)stream * model_voxels(h), g.dx, g.dy, g.dz, d_cnt);
  h->launches++;
  if (cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
  return (long long)cnt;
}

extern "C" int youth_cuda_debug_read_volume(youth_cuda_handle* h, int stream, int16_t* dst, size_t dst_bytes) {
  if (!h || !dst) return fail("null argument");
  if (!h->m.on) return fail("frame-to-model tracking is not enabled");
  if (stream < 0 || stream >= h->S) return fail("stream out of range");
  const size_t nv = model_voxels(h);
  if (dst_bytes < nv * 4) return fail("dst too small");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaMemcpy(dst, h->m.vol + (size_t)stream * nv, nv * 4, cudaMemcpyDeviceToHost));
  return 1;
}

extern "C" int youth_cuda_debug_read_model(youth_cuda_handle* h, int what, int stream, int level, float* dst, size_t dst_bytes) {
  if (!h || !dst) return fail("null argument");
  if (!h->m.on) return fail("frame-to-model tracking is not enabled");
  if (stream < 0 || stream >= h->S || level < 0 || level >= h->cfg.levels) return fail("stream/level out of range");
  if (what != YOUTH_DBG_VERTEX && what != YOUTH_DBG_NORMAL) return fail("what must be YOUTH_DBG_VERTEX or YOUTH_DBG_NORMAL");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  const size_t np = (size_t)h->npix[level];
  return read_planes(h->m.maps[level] + (size_t)stream * 3 * np, np, what, dst, dst_bytes);
}

static int slot_of_frame(const youth_cuda_handle* h, int frame);

extern "C" int youth_cuda_debug_integrate(youth_cuda_handle* h, int stream, int frame, const float pose[12]) {
  if (!h || !pose) return fail("null argument");
  if (!h->m.on) return fail("frame-to-model tracking is not enabled");
  if (stream < 0 || stream >= h->S) return fail("stream out of range");
  const int slot = slot_of_frame(h, frame);
  if (slot < 0) return fail("frame %d is not resident in the ring", frame);
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->m.world_f + stream * 12, pose, sizeof(float) * 12, cudaMemcpyHostToDevice, h->stream));
  if (!enqueue_integrate(h, stream, 1, slot, false)) return 0;
  CU(cudaStreamSynchronize(h->stream));
  return 1;
}

extern "C" int youth_cuda_debug_raycast(youth_cuda_handle* h, int stream, const float pose[12], int hint_frame) {
  if (!h || !pose) return fail("null argument");
  if (!h->m.on) return fail("frame-to-model tracking is not enabled");
  if (stream < 0 || stream >= h->S) return fail("stream out of range");
  int hint_slot = -2;
  if (hint_frame >= 0) {
    hint_slot = slot_of_frame(h, hint_frame);
    if (hint_slot < 0) return fail("frame %d is not resident in the ring", hint_frame);
  }
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaMemcpyAsync(h->m.world_f + stream * 12, pose, sizeof(float) * 12, cudaMemcpyHostToDevice, h->stream));
  if (!enqueue_raycast(h, stream, 1, hint_slot)) return 0;
  CU(cudaStreamSynchronize(h->stream));
  return 1;
}

extern "C" int youth_cuda_debug_read(youth_cuda_handle* h, int what, int stream, int frame, int level, void* dst,
                                     size_t dst_bytes) {
  if (!h || !dst) return fail("null argument");
  if (stream < 0 || stream >= h->S || level < 0 || level >= h->cfg.levels) return fail("stream/level out of range");
  const int slot = slot_of_frame(h, frame);
  if (slot < 0) return fail("frame %d is not resident in the ring", frame);
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  const size_t np = (size_t)h->npix[level];
  const size_t off = ((size_t)stream * h->R + slot) * np;
  switch (what) {
    case YOUTH_DBG_DEPTH:
      if (!h->debug_maps) return fail("the depth pyramid is only stored after youth_cuda_debug_enable_maps (call it before tracking)");
      if (dst_bytes < np * 4) return fail("dst too small");
      CU(cudaMemcpy(dst, h->depth[level] + off, np * 4, cudaMemcpyDeviceToHost));
      return 1;
    case YOUTH_DBG_VERTEX:
    case YOUTH_DBG_NORMAL:
    case YOUTH_DBG_MASK: {
      return read_planes(h->maps[level] + ((size_t)stream * h->R + slot) * 3 * np, np, what, dst, dst_bytes);
    }
    case YOUTH_DBG_PYRCNT:
      if (!h->debug_maps) return fail("the pyramid sample counts are only stored after youth_cuda_debug_enable_maps (call it before tracking)");
      if (dst_bytes < np) return fail("dst too small");
      CU(cudaMemcpy(dst, h->pyrcnt[level] + off, np, cudaMemcpyDeviceToHost));
      return 1;
    default:
      return fail("unknown debug selector %d", what);
  }
}

extern "C" long long youth_cuda_debug_rcp_check(youth_cuda_handle* h, uint32_t lo_bits, uint32_t hi_bits) {
  if (!h) {
    fail("null handle");
    return -1;
  }
  if (lo_bits > hi_bits) {
    fail("lo_bits > hi_bits");
    return -1;
  }
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return -1;
  unsigned long long* d = NULL;
  unsigned long long bad = 0;
  if (cudaMalloc((void**)&d, sizeof(bad)) != cudaSuccess) return -1;
  cudaMemsetAsync(d, 0, sizeof(bad), h->stream);
  yk_launch_rcp_check(h->stream, lo_bits, hi_bits, d);
  h->launches++;
  cudaMemcpyAsync(&bad, d, sizeof(bad), cudaMemcpyDeviceToHost, h->stream);
  const cudaError_t e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) {
    fail("k_rcp_check failed: %s", cudaGetErrorString(e));
    return -1;
  }
  return (long long)bad;
}

extern "C" int youth_cuda_debug_icp(youth_cuda_handle* h, int stream, int frame, int level, const float pose[12],
                                    double* sums_out, int32_t* corr_out) {
  if (!h || !pose || !sums_out) return fail("null argument");
  if (stream < 0 || stream >= h->S || level < 0 || level >= h->cfg.levels) return fail("stream/level out of range");
  const int cur = slot_of_frame(h, frame), prev = slot_of_frame(h, frame - 1);
  if (cur < 0 || prev < 0) return fail("frames %d and %d must both be resident in the ring", frame - 1, frame);
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaStreamSynchronize(h->stream));
  /* pair slot 0 is free between groups: borrow it */
  CU(cudaMemcpyAsync(h->pose_f, pose, sizeof(float) * 12, cudaMemcpyHostToDevice, h->stream));
  RingGeom ring = ring_of(h, 1);
  IcpParams ip = icp_params(h, level, ring);
  ip.corr = corr_out ? h->corr_dbg : NULL;
  ip.dbg_cur_slot = cur;
  ip.dbg_prev_slot = prev;
  ip.dbg_stream = stream;
  ip.do_solve = 0;
  ip.model = NULL; /* always frame `frame` against frame - 1, also on a frame-to-model handle */
  {
    ProfScope ps(h, YOUTH_PROF_ICP0 + level);
    const dim3 grid((h->nruns[level] + YK_ICP_WARPS - 1) / YK_ICP_WARPS, 1);
    yk_launch_icp_debug(grid, h->stream, ip);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(sums_out, h->sums, sizeof(double) * 32, cudaMemcpyDeviceToHost, h->stream));
  if (corr_out)
    CU(cudaMemcpyAsync(corr_out, h->corr_dbg, sizeof(int32_t) * h->npix[level], cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return 1;
}

extern "C" int youth_cuda_timer_start(youth_cuda_handle* h) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaEventRecord(h->t0, h->stream));
  return 1;
}

extern "C" int youth_cuda_timer_stop(youth_cuda_handle* h, float* ms_out) {
  if (!h || !ms_out) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  CU(cudaEventRecord(h->t1, h->stream));
  CU(cudaEventSynchronize(h->t1));
  CU(cudaEventElapsedTime(ms_out, h->t0, h->t1));
  return 1;
}

extern "C" uint64_t youth_cuda_launch_count(youth_cuda_handle* h) { return h ? h->launches : 0; }

extern "C" int youth_cuda_profile_enable(youth_cuda_handle* h, int on) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->cfg.device));
  if (on && !h->prof_ev) {
    h->prof_cap = 4096;
    h->prof_ev = (cudaEvent_t*)calloc((size_t)2 * h->prof_cap, sizeof(cudaEvent_t));
    h->prof_cls = (int*)calloc((size_t)h->prof_cap, sizeof(int));
    if (!h->prof_ev || !h->prof_cls) return fail("host allocation failed");
    for (int i = 0; i < 2 * h->prof_cap; ++i) CU(cudaEventCreate(&h->prof_ev[i]));
  }
  if (!on && h->prof_on && !prof_flush(h)) return 0;
  if (on && !h->prof_on) {
    memset(h->prof_ms, 0, sizeof(h->prof_ms));
    memset(h->prof_launches, 0, sizeof(h->prof_launches));
    h->prof_n = 0;
  }
  h->prof_on = on != 0;
  return 1;
}

extern "C" int youth_cuda_profile_read(youth_cuda_handle* h, double* ms_out, uint64_t* launches_out) {
  if (!h || !ms_out || !launches_out) return fail("null argument");
  CU(cudaSetDevice(h->cfg.device));
  if (!prof_flush(h)) return 0;
  for (int i = 0; i < YOUTH_PROF_CLASSES; ++i) {
    ms_out[i] = h->prof_ms[i];
    launches_out[i] = h->prof_launches[i];
  }
  return 1;
}
