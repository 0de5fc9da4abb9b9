// membench.cu -- practical memory roofline of the k_icp access pattern on B200:
// per pixel, stream the current frame's maps and gather the previous frame's maps at a
// nearby pixel, for 32 frame pairs of 640x480 (same footprint as one level-0 ICP launch).
// Loads are staged in registers and consumed SD (stream) / GD (gather) iterations later, so
// each warp really has that many requests in flight.  Trivial arithmetic: the time is what
// the memory system allows at the given occupancy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/membench tools/membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int NPIX = 640 * 480, PAIRS = 32, FRAMES = PAIRS + 1;

__device__ __forceinline__ int gather_index(int p) {
  int q = p + 640 + 2 + ((p >> 7) & 1);
  return q < NPIX ? q : p;
}

struct Rec { float4 a; float4 b; };

// MODE 0: two float4 arrays (32 B/px/frame); MODE 2: three float2 planes (24 B/px/frame);
// MODE 3: the same layout, planes 1 and 2 only (16 B/px/frame) -- what k_icp requests when it recomputes
// vx, vy from vz (-DYK_ICP_XY=3)
template <int MODE>
__device__ __forceinline__ Rec load_rec(const float* f, int p) {
  Rec r;
  if (MODE == 0) {
    r.a = __ldg((const float4*)f + p);
    r.b = __ldg((const float4*)(f + (size_t)NPIX * 4) + p);
  } else if (MODE == 3) {
    const float2 y = __ldg((const float2*)(f + (size_t)NPIX * 2) + p), z = __ldg((const float2*)(f + (size_t)NPIX * 4) + p);
    r.a = make_float4(y.x, y.y, z.x, z.y);
    r.b = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    const float2 x = __ldg((const float2*)f + p), y = __ldg((const float2*)(f + (size_t)NPIX * 2) + p),
                 z = __ldg((const float2*)(f + (size_t)NPIX * 4) + p);
    r.a = make_float4(x.x, x.y, y.x, y.y);
    r.b = make_float4(z.x, z.y, 0.f, 0.f);
  }
  return r;
}
__device__ __forceinline__ float sum_rec(const Rec& r) { return r.a.x + r.a.y + r.a.z + r.a.w + r.b.x + r.b.y + r.b.z + r.b.w; }

template <int MODE, int SD, int GD>
__global__ void __launch_bounds__(128) k(const float* __restrict__ base, float* __restrict__ out, int ppr, int nruns) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int blocks_per_pair = (nruns + 3) / 4;
  // persistent CTAs (grid = resident CTAs) so that occupancy is limited without giving up L1
  for (int work = blockIdx.x; work < blocks_per_pair * PAIRS; work += gridDim.x) {
  const int pair = work / blocks_per_pair;
  const int run = (work - pair * blocks_per_pair) * 4 + warp;
  if (run >= nruns) continue;
  const size_t fstride = (size_t)NPIX * 8;
  const float* cur = base + (size_t)(pair + 1) * fstride;
  const float* prv = base + (size_t)pair * fstride;
  const int p0 = run * 32 + lane, pstep = 32 * nruns;
  Rec s[SD], g[GD];
  const Rec zero = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
#pragma unroll
  for (int d = 0; d < SD; ++d) s[d] = (d < ppr && p0 + d * pstep < NPIX) ? load_rec<MODE>(cur, p0 + d * pstep) : zero;
#pragma unroll
  for (int d = 0; d < GD; ++d) g[d] = zero;
  float acc = 0.f;
  for (int j = 0; j < ppr; ++j) {
    const int p = p0 + j * pstep;
    // consume the oldest streamed record -> (fake) projection -> issue its gather
    const float sv = sum_rec(s[0]);
    acc += sum_rec(g[0]);  // consume the oldest gather
#pragma unroll
    for (int d = 0; d + 1 < GD; ++d) g[d] = g[d + 1];
    g[GD - 1] = (p < NPIX) ? load_rec<MODE>(prv, gather_index(p) + (sv > 1e30f ? 1 : 0)) : zero;
#pragma unroll
    for (int d = 0; d + 1 < SD; ++d) s[d] = s[d + 1];
    const int pn = p0 + (j + SD) * pstep;
    s[SD - 1] = (j + SD < ppr && pn < NPIX) ? load_rec<MODE>(cur, pn) : zero;
    acc += sv;
  }
#pragma unroll
  for (int d = 0; d < GD; ++d) acc += sum_rec(g[d]);
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[pair * nruns + run] = acc;
  }
}

template <int MODE, int SD, int GD>
void run(const float* base, float* out, int ppr, int blocks_per_sm) {
  const double bytes_per_px = MODE == 0 ? 64 : (MODE == 3 ? 32 : 48);
  int maxb = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k<MODE, SD, GD>, 128, 0));
  if (blocks_per_sm > maxb) return;
  const int smem = 0;
  const int nruns = (NPIX + 32 * ppr - 1) / (32 * ppr);
  dim3 grid(148 * blocks_per_sm);
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) k<MODE, SD, GD><<<grid, 128, smem>>>(base, out, ppr, nruns);
  CK(cudaEventRecord(a));
  const int reps = 20;
  for (int i = 0; i < reps; ++i) k<MODE, SD, GD><<<grid, 128, smem>>>(base, out, ppr, nruns);
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  const double us = ms * 1e3 / reps, gb = bytes_per_px * NPIX * PAIRS / 1e9;
  printf("%s SD=%d GD=%d ppr=%3d warps/SM=%2d (max %2d) %7.1f us/launch %6.0f GB/s requested\n",
         MODE == 0 ? "float4x2 " : (MODE == 3 ? "2xfloat2 " : "3xfloat2 "), SD, GD, ppr, 4 * (blocks_per_sm < maxb ? blocks_per_sm : maxb), 4 * maxb, us,
         gb / (us * 1e-6));
}

int main() {
  float *base, *out;
  CK(cudaMalloc(&base, (size_t)FRAMES * NPIX * 32));
  CK(cudaMemset(base, 0, (size_t)FRAMES * NPIX * 32));
  CK(cudaMalloc(&out, 1 << 22));
  /* the depth x occupancy points of the k_icp variants (make next-variants): 20 warps x 2 (today), 16 x 3, 12 x 4,
   * 12 x 5, 12 x 6 pixels in flight per lane, three-plane and two-plane records, ppr 128 as bench.py runs */
  run<2, 2, 2>(base, out, 128, 5);
  run<2, 3, 3>(base, out, 128, 4);
  run<2, 4, 4>(base, out, 128, 3);
  run<3, 2, 2>(base, out, 128, 5);
  run<3, 3, 3>(base, out, 128, 4);
  run<3, 5, 5>(base, out, 128, 3);
  run<3, 6, 6>(base, out, 128, 3);
  for (int ppr : {32, 64}) {
    for (int bps : {5, 6, 8, 12}) {
      run<0, 1, 1>(base, out, ppr, bps);
      run<0, 2, 1>(base, out, ppr, bps);
      run<0, 2, 2>(base, out, ppr, bps);
      run<0, 3, 2>(base, out, ppr, bps);
      run<2, 1, 1>(base, out, ppr, bps);
      run<2, 2, 1>(base, out, ppr, bps);
      run<2, 2, 2>(base, out, ppr, bps);
      run<2, 3, 2>(base, out, ppr, bps);
      run<3, 2, 2>(base, out, ppr, bps);
      run<3, 3, 2>(base, out, ppr, bps);
    }
  }
  return 0;
}
