// membench2.cu -- where does the level-0 k_icp access pattern stop scaling on B200?
// Same geometry as the bench: 300 frame pairs of 640x480, three float2 planes per frame, grid (19, 300) x 128,
// one warp per run of 32 x 128 pixels, lane l of run k walks pixels j * 32 * nruns + 32 * k + l.
// Each mode isolates one half of the pattern (trivial arithmetic, so the time is what the memory system allows):
//   S   stream only (3 x LDG.64 per pixel, SD deep)
//   G   gather only (address does not depend on a load)
//   SG  stream -> dependent gather (what k_icp does), SD / GD deep
//   S4  stream only with 128-bit loads (a lane owns two adjacent pixels)
//   B   stream through a per-warp shared-memory ring filled by 1-D bulk async copies (3 x 256 B per step,
//       complete_tx on one mbarrier per slot), DS deep, + dependent gather GD deep in registers
//   BG  as B, and the gather through per-lane cp.async (8 B) into a second ring, DG deep
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/membench2 tools/membench2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int W = 640, NPIX = 640 * 480;
constexpr int PPR = 128, NRUNS = (NPIX + 32 * PPR - 1) / (32 * PPR); // 75
constexpr int PSTEP = 32 * NRUNS;

struct Rec { float2 a, b, c; };
__device__ __forceinline__ float sum_rec(const Rec& r) { return r.a.x + r.a.y + r.b.x + r.b.y + r.c.x + r.c.y; }
__device__ __forceinline__ Rec ld_rec(const float2* f, int p) {
  Rec r;
  r.a = __ldg(f + p);
  r.b = __ldg(f + NPIX + p);
  r.c = __ldg(f + 2 * NPIX + p);
  return r;
}
__device__ __forceinline__ int gather_index(int p) {
  int q = p + W + 2 + ((p >> 7) & 1);
  return q < NPIX ? q : p;
}

enum { M_S = 0, M_G = 1, M_SG = 2, M_S4 = 3 };

template <int MODE, int SD, int GD, int MINB>
__global__ void __launch_bounds__(128, MINB) k(const float2* __restrict__ base, float* __restrict__ out, int persistent, int pairs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // persistent: grid = SMs x wanted CTAs per SM, every CTA walks the (19, pairs) work items -- occupancy is capped
  // without touching the L1 / shared-memory split
  for (int work = persistent ? blockIdx.x : 0; work < (persistent ? 19 * pairs : 1); work += gridDim.x) {
  const int bx = persistent ? work % 19 : blockIdx.x, by = persistent ? work / 19 : blockIdx.y;
  const int run = bx * 4 + warp, pair = by;
  if (run >= NRUNS) continue;
  const float2* cur = base + (size_t)(pair + 1) * 3 * NPIX;
  const float2* prv = base + (size_t)pair * 3 * NPIX;
  float acc = 0.f;
  if (MODE == M_S4) {
    // a lane owns pixels 2l, 2l+1 of a 64-pixel span; run k step j covers pixels j * 64 * nruns2 + 64 * k ...
    constexpr int NR2 = NRUNS, PST2 = 64 * NR2; // 64 steps instead of 128
    const int p0 = run * 64 + 2 * lane;
    float4 s[SD][3];
#pragma unroll
    for (int d = 0; d < SD; ++d)
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) s[d][pl] = __ldg((const float4*)(cur + pl * NPIX + p0 + d * PST2));
    for (int j = 0; j < PPR / 2; ++j) {
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) acc += s[0][pl].x + s[0][pl].y + s[0][pl].z + s[0][pl].w;
#pragma unroll
      for (int d = 0; d + 1 < SD; ++d)
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) s[d][pl] = s[d + 1][pl];
      const int pn = p0 + (j + SD) * PST2;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl)
        s[SD - 1][pl] = (j + SD < PPR / 2 && pn < NPIX) ? __ldg((const float4*)(cur + pl * NPIX + pn)) : make_float4(0, 0, 0, 0);
    }
  } else {
    const int p0 = run * 32 + lane;
    Rec s[SD], g[GD];
    const Rec zero = {make_float2(0, 0), make_float2(0, 0), make_float2(0, 0)};
#pragma unroll
    for (int d = 0; d < SD; ++d) s[d] = (MODE != M_G && p0 + d * PSTEP < NPIX) ? ld_rec(cur, p0 + d * PSTEP) : zero;
#pragma unroll
    for (int d = 0; d < GD; ++d) g[d] = zero;
    for (int j = 0; j < PPR; ++j) {
      const int p = p0 + j * PSTEP;
      const float sv = sum_rec(s[0]);
      acc += sum_rec(g[0]);
#pragma unroll
      for (int d = 0; d + 1 < GD; ++d) g[d] = g[d + 1];
      if (MODE != M_S) g[GD - 1] = (p < NPIX) ? ld_rec(prv, gather_index(p) + (sv > 1e30f ? 1 : 0)) : zero;
#pragma unroll
      for (int d = 0; d + 1 < SD; ++d) s[d] = s[d + 1];
      const int pn = p0 + (j + SD) * PSTEP;
      if (MODE != M_G) s[SD - 1] = (j + SD < PPR && pn < NPIX) ? ld_rec(cur, pn) : zero;
      acc += sv;
    }
#pragma unroll
    for (int d = 0; d < GD; ++d) acc += sum_rec(g[d]);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[pair * NRUNS + run] = acc;
  }
}

// ---- bulk-copy ring ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// MODE B: DS-deep bulk ring for the stream, GD-deep register gather.  MODE BG (DG > 0): gather through cp.async ring
template <int DS, int GD, int DG, int MINB>
__global__ void __launch_bounds__(128, MINB) kb(const float2* __restrict__ base, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int run = blockIdx.x * 4 + warp, pair = blockIdx.y;
  // per warp: DS slots x 3 planes x 256 B, then DS mbarriers; then DG slots x 3 x 256 B gather ring
  constexpr int RING = DS * 768, GRING = DG * 768;
  unsigned char* wbase = smem + warp * (RING + GRING + 8 * DS);
  float2* ring = (float2*)wbase;
  float2* gring = (float2*)(wbase + RING);
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t gring_s = (uint32_t)__cvta_generic_to_shared(gring);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(wbase + RING + GRING);
  if (run >= NRUNS) return;
  const float2* cur = base + (size_t)(pair + 1) * 3 * NPIX;
  const float2* prv = base + (size_t)pair * 3 * NPIX;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < DS; ++d) mbar_init(bar_s + 8 * d, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int w0 = run * 32; // first pixel of the warp's span at step 0
  auto issue = [&](int j) { // lane 0 only; all spans whole (npix = 75 * 32 * 128)
    const int slot = j % DS;
    const uint32_t bar = bar_s + 8 * slot;
    mbar_expect_tx(bar, 768);
    const float2* src = cur + w0 + (size_t)j * PSTEP;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) bulk_g2s(ring_s + slot * 768 + pl * 256, src + (size_t)pl * NPIX, 256, bar);
  };
  if (lane == 0)
    for (int d = 0; d < DS; ++d) issue(d);
  const int p0 = run * 32 + lane;
  Rec g[GD > 0 ? GD : 1];
  const Rec zero = {make_float2(0, 0), make_float2(0, 0), make_float2(0, 0)};
#pragma unroll
  for (int d = 0; d < (GD > 0 ? GD : 1); ++d) g[d] = zero;
  float acc = 0.f;
  for (int j = 0; j < PPR; ++j) {
    const int slot = j % DS;
    mbar_wait(bar_s + 8 * slot, (j / DS) & 1);
    const float2* cell = ring + slot * 96 + lane;
    const Rec sc = {cell[0], cell[32], cell[64]};
    const float sv = sum_rec(sc);
    __syncwarp(); // every lane has read the slot
    if (lane == 0 && j + DS < PPR) issue(j + DS);
    const int p = p0 + j * PSTEP;
    const int q = gather_index(p) + (sv > 1e30f ? 1 : 0);
    if constexpr (DG > 0) {
      // consume the gather issued DG steps ago, then issue this step's into the same slot
      const int gs = j % DG;
      asm volatile("cp.async.wait_group %0;" ::"n"(DG - 1) : "memory");
      const float2* gc = gring + gs * 96 + lane;
      if (j >= DG) acc += gc[0].x + gc[0].y + gc[32].x + gc[32].y + gc[64].x + gc[64].y;
      const uint32_t dst = gring_s + gs * 768 + lane * 8;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n\t"
                   "cp.async.ca.shared.global [%0 + 256], [%2], 8;\n\t"
                   "cp.async.ca.shared.global [%0 + 512], [%3], 8;\n\t"
                   "cp.async.commit_group;" ::"r"(dst), "l"(prv + q), "l"(prv + NPIX + q), "l"(prv + 2 * NPIX + q)
                   : "memory");
    } else if constexpr (GD > 0) {
      acc += sum_rec(g[0]);
#pragma unroll
      for (int d = 0; d + 1 < GD; ++d) g[d] = g[d + 1];
      g[GD - 1] = ld_rec(prv, q);
    }
    acc += sv;
  }
  if (DG > 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (DG == 0 && GD > 0)
#pragma unroll
    for (int d = 0; d < GD; ++d) acc += sum_rec(g[d]);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[pair * NRUNS + run] = acc;
}

// ---- prefetch-to-L1 pipeline: depth without registers ----
// stream: prefetch.global.L1 of pixel j + DS, register load of pixel j + 1 (an L1 hit if the prefetch landed);
// gather: prefetch of the matched record when its index is known, register load DG steps later.
__device__ __forceinline__ void pf_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int DS, int DG, int MINB>
__global__ void __launch_bounds__(128, MINB) kp(const float2* __restrict__ base, float* __restrict__ out, int pairs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int work = blockIdx.x; work < 19 * pairs; work += gridDim.x) {
    const int bx = work % 19, pair = work / 19;
    const int run = bx * 4 + warp;
    if (run >= NRUNS) continue;
    const float2* cur = base + (size_t)(pair + 1) * 3 * NPIX;
    const float2* prv = base + (size_t)pair * 3 * NPIX;
    const int p0 = run * 32 + lane;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < DS; ++d) {
      const float2* a = cur + p0 + d * PSTEP;
      pf_l1(a); pf_l1(a + NPIX); pf_l1(a + 2 * NPIX);
    }
    int q[DG];
#pragma unroll
    for (int d = 0; d < DG; ++d) q[d] = p0;
    Rec s = ld_rec(cur, p0);
#pragma unroll 2
    for (int j = 0; j < PPR; ++j) {
      const int p = p0 + j * PSTEP;
      if (j + DS < PPR) {
        const float2* a = cur + p + DS * PSTEP;
        pf_l1(a); pf_l1(a + NPIX); pf_l1(a + 2 * NPIX);
      }
      const float sv = sum_rec(s);
      if (j + 1 < PPR) s = ld_rec(cur, p + PSTEP);
      const int qn = gather_index(p) + (sv > 1e30f ? 1 : 0);
      pf_l1(prv + qn); pf_l1(prv + NPIX + qn); pf_l1(prv + 2 * NPIX + qn);
      if (j >= DG) acc += sum_rec(ld_rec(prv, q[0]));
#pragma unroll
      for (int d = 0; d + 1 < DG; ++d) q[d] = q[d + 1];
      q[DG - 1] = qn;
      acc += sv;
    }
#pragma unroll
    for (int d = 0; d < DG; ++d) acc += sum_rec(ld_rec(prv, q[d]));
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[pair * NRUNS + run] = acc;
  }
}

static const float2* g_base;
static float* g_out;
static int g_pairs = 300;

template <typename F>
static void time_it(const char* name, int minb, double bytes_per_px, F launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaGetLastError());
  CK(cudaEventRecord(a));
  const int reps = 10;
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double us = ms * 1e3 / reps, gb = bytes_per_px * NPIX * g_pairs / 1e9;
  printf("%-28s CTAs/SM<=%d  %8.1f us/launch  %7.0f GB/s requested  (%.3f us/pair)\n", name, minb, us, gb / (us * 1e-6), us / g_pairs);
  fflush(stdout);
}

// force = resident CTAs per SM wanted (0: whatever the registers allow), through a persistent grid
template <int MODE, int SD, int GD, int MINB>
static void run(const char* name, double bpp, int force = 0) {
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k<MODE, SD, GD, MINB>, 128, 0));
  if (force > maxb) return;
  char nm[64];
  snprintf(nm, sizeof nm, "%s SD=%d GD=%d occ=%d", name, SD, GD, force ? force : maxb);
  if (force)
    time_it(nm, MINB, bpp, [force] { k<MODE, SD, GD, MINB><<<dim3(148 * force), 128>>>(g_base, g_out, 1, g_pairs); });
  else
    time_it(nm, MINB, bpp, [] { k<MODE, SD, GD, MINB><<<dim3(19, g_pairs), 128>>>(g_base, g_out, 0, g_pairs); });
}

template <int DS, int GD, int DG, int MINB>
static void runb(const char* name) {
  const int smem = 4 * (DS * 768 + DG * 768 + 8 * DS);
  CK(cudaFuncSetAttribute(kb<DS, GD, DG, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kb<DS, GD, DG, MINB>, 128, smem));
  char nm[64];
  snprintf(nm, sizeof nm, "%s DS=%d GD=%d DG=%d occ=%d", name, DS, GD, DG, maxb);
  time_it(nm, MINB, (GD > 0 || DG > 0) ? 48 : 24, [smem] { kb<DS, GD, DG, MINB><<<dim3(19, g_pairs), 128, smem>>>(g_base, g_out); });
}

template <int DS, int DG, int MINB>
static void runp(int force) {
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, kp<DS, DG, MINB>, 128, 0));
  if (force > maxb) return;
  char nm[64];
  snprintf(nm, sizeof nm, "P   prefetch.L1 DS=%d DG=%d occ=%d", DS, DG, force);
  time_it(nm, MINB, 48, [force] { kp<DS, DG, MINB><<<dim3(148 * force), 128>>>(g_base, g_out, g_pairs); });
}

int main(int argc, char** argv) {
  if (argc > 1) g_pairs = atoi(argv[1]);
  float2* base;
  float* out;
  CK(cudaMalloc(&base, (size_t)(g_pairs + 1) * NPIX * 24));
  CK(cudaMemset(base, 0, (size_t)(g_pairs + 1) * NPIX * 24));
  CK(cudaMalloc(&out, 1 << 22));
  g_base = base;
  g_out = out;
  printf("pairs=%d  unique bytes per launch %.3f GB (one DRAM pass at 6547.5 GB/s = %.1f us)\n", g_pairs, (g_pairs + 1.0) * NPIX * 24 / 1e9,
         (g_pairs + 1.0) * NPIX * 24 / 6547.5e9 * 1e6);
  // stream only
  run<M_S, 2, 1, 5>("S   stream LDG.64", 24);
  run<M_S, 4, 1, 5>("S   stream LDG.64", 24);
  run<M_S, 8, 1, 5>("S   stream LDG.64", 24);
  run<M_S, 4, 1, 8>("S   stream LDG.64", 24);
  run<M_S, 4, 1, 12>("S   stream LDG.64", 24);
  run<M_S4, 2, 1, 5>("S4  stream LDG.128", 24);
  run<M_S4, 4, 1, 5>("S4  stream LDG.128", 24);
  run<M_S4, 4, 1, 8>("S4  stream LDG.128", 24);
  // gather only
  run<M_G, 1, 2, 5>("G   gather LDG.64", 24);
  run<M_G, 1, 4, 5>("G   gather LDG.64", 24);
  run<M_G, 1, 4, 8>("G   gather LDG.64", 24);
  // both, register pipelines
  run<M_SG, 2, 2, 5>("SG  stream+gather", 48);
  run<M_SG, 3, 3, 4>("SG  stream+gather", 48);
  run<M_SG, 4, 4, 4>("SG  stream+gather", 48);
  run<M_SG, 2, 2, 8>("SG  stream+gather", 48);
  run<M_SG, 2, 2, 12>("SG  stream+gather", 48);
  run<M_SG, 4, 4, 8>("SG  stream+gather", 48);
  // the same at the occupancies k_icp can have (96 registers: 5 CTAs, 128: 4, 80: 6, 64: 8)
  for (int f : {3, 4, 5, 6, 8}) {
    run<M_SG, 2, 2, 4>("SG  forced occupancy", 48, f);
    run<M_SG, 3, 3, 4>("SG  forced occupancy", 48, f);
    run<M_SG, 4, 4, 4>("SG  forced occupancy", 48, f);
    run<M_SG, 8, 4, 3>("SG  forced occupancy", 48, f);
    run<M_SG, 8, 8, 3>("SG  forced occupancy", 48, f);
  }
  for (int f : {4, 5, 6, 8}) {
    runp<4, 2, 4>(f);
    runp<4, 4, 4>(f);
    runp<8, 4, 4>(f);
    runp<8, 8, 4>(f);
    runp<12, 6, 4>(f);
  }
  // bulk ring stream
  runb<4, 0, 0, 5>("B   bulk stream only");
  runb<8, 0, 0, 5>("B   bulk stream only");
  runb<8, 0, 0, 8>("B   bulk stream only");
  runb<8, 2, 0, 5>("B   bulk + reg gather");
  runb<8, 4, 0, 5>("B   bulk + reg gather");
  runb<8, 2, 0, 8>("B   bulk + reg gather");
  runb<8, 4, 0, 8>("B   bulk + reg gather");
  runb<8, 0, 4, 5>("BG  bulk + cp.async gather");
  runb<8, 0, 8, 5>("BG  bulk + cp.async gather");
  runb<8, 0, 8, 8>("BG  bulk + cp.async gather");
  runb<8, 0, 8, 12>("BG  bulk + cp.async gather");
  runb<4, 0, 4, 16>("BG  bulk + cp.async gather");
  return 0;
}
