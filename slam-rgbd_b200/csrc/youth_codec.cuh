/*
 * youth_codec.cuh -- sm_100a kernels and context of the YD16 lossless depth codec
 * (format and C ABI: include/youth_codec.h; CPU statement: oracle/youth_codec_oracle.c).
 * Included by youth_cuda.cu (one translation unit, one libyouth_cuda.so).
 *
 * Byte/integer work, HBM-bound by design: one pass over the raw frame (2 B/pixel in) and one pass
 * over the packed stream (about 0.66 B/pixel out on Astra-shaped depth).
 *
 *   k_yd16_encode   a CTA owns a tile of 256 blocks (8192 pixels): 128-bit loads into a swizzled shared
 *                   tile; ONE THREAD PER BLOCK (its 64 bytes in 16 registers: mask, first reading and bit
 *                   width in one pass, codes streamed through a 64-bit accumulator in a second) -- about
 *                   a fifth of the instructions of a warp-per-block formulation with ballots and
 *                   shuffles, which ran at 9 % of the HBM roofline; the CTA scans the block sizes, the
 *                   tile's place in the stream comes from a decoupled look-back over the tile descriptors
 *                   of the same frame (single pass: the raw frame is read once), and the staged payload
 *                   is written with aligned 32-bit stores (funnel shift).  Tiles take their index from
 *                   an atomic ticket, so a tile only ever waits for tiles that already started.
 *   k_yd16_tilesum  per-tile sums of the block-size table (decoder pre-pass, 1 B/block)
 *   k_yd16_decode   tile payload staged in shared memory with aligned 32-bit loads; one thread per block:
 *                   unaligned bit-field reads by funnel shift, running sum of the deltas, 128-bit
 *                   stores of the unpacked tile through the swizzled shared tile
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "youth_codec.h"

#define YC_TILE_BLOCKS 256
#define YC_THREADS 256
#define YC_WARPS 8
#define YC_SPIN_LIMIT (1u << 24)
#define YC_ERR_SPIN 1u
#define YC_ERR_HEADER 2u
#define YC_ERR_SIZES 4u
#define YC_ERR_BLOCK 8u

struct YcEncParams {
  const uint16_t* in;  /* n frames, tightly packed */
  uint8_t* out;        /* frame i at out + i * out_stride */
  size_t out_stride;
  int width, height, npix;
  uint32_t nb;         /* blocks per frame */
  int tiles;           /* tiles per frame */
  int n_frames;
  int vec_ok;          /* frames are 16-byte aligned: 128-bit loads */
  unsigned long long* desc; /* [n][tiles] look-back descriptors, zero before the launch */
  unsigned int* ticket;     /* zero before the launch */
  uint32_t* frame_bytes;    /* [n] stream length of every frame */
  unsigned int* err;
};

struct YcDecParams {
  const uint8_t* in;          /* all streams, device copy */
  const unsigned long long* offsets; /* [n+1], relative to `in` */
  uint16_t* out;              /* n frames tightly packed */
  int width, height, npix;
  uint32_t nb;
  int tiles;
  int n_frames;
  int vec_ok;
  uint32_t* tile_sum;         /* [n][tiles] */
  unsigned int* err;
};

__device__ __forceinline__ unsigned long long yc_ld_volatile(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void yc_st_volatile(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* 32 bits starting at byte position `pos` of a word array (needs one word of slack) */
__device__ __forceinline__ uint32_t yc_rd32(const uint32_t* w, uint32_t pos) {
  return __funnelshift_r(w[pos >> 2], w[(pos >> 2) + 1], (pos & 3u) * 8u);
}

/* physical 16-byte chunk of logical chunk k (0..3) of block b inside the shared tile: the XOR keeps the
 * per-thread 16-byte accesses (thread = block, 64-byte stride) free of bank conflicts */
__device__ __forceinline__ int yc_chunk(int b, int k) { return (b << 2) | (k ^ ((b >> 1) & 3)); }

/* load one tile (256 blocks = 1024 chunks of 8 pixels) of a frame into the swizzled shared tile */
__device__ __forceinline__ void yc_load_tile(uint4* s_in, const uint16_t* src, int p0, int npix, int vec_ok, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tid + YC_THREADS * i;
    const int px = p0 + 8 * c;
    uint4 q;
    if (vec_ok && px + 8 <= npix) {
      q = ldg_nc_u4(reinterpret_cast<const uint4*>(src + px));
    } else {
      uint32_t h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = (px + j < npix) ? (uint32_t)src[px + j] : 0u;
      q = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    }
    s_in[yc_chunk(c >> 2, c & 3)] = q;
  }
}

/* exclusive scan of one value per thread over the CTA (256 threads); *total = sum */
__device__ __forceinline__ uint32_t yc_block_scan(uint32_t v, uint32_t* s_wsum, int tid, uint32_t* total) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  uint32_t base = 0, all = 0;
#pragma unroll
  for (int w = 0; w < YC_WARPS; ++w) {
    const uint32_t x = s_wsum[w];
    if (w < warp) base += x;
    all += x;
  }
  *total = all;
  return base + inc - v;
}

/* One THREAD per block of 32 pixels (256 blocks per CTA): the block's 64 bytes sit in 16 registers, the
 * mask / first value / bit width come from one pass over them, the CTA scans the 256 block sizes, a second
 * pass streams the zig-zag codes through a 64-bit accumulator into the block's place in the shared
 * staging buffer; the tile's place in the stream comes from the decoupled look-back; the whole CTA
 * copies the staged payload out with aligned 32-bit stores. */
__global__ void __launch_bounds__(YC_THREADS) k_yd16_encode(const __grid_constant__ YcEncParams P) {
  __shared__ __align__(16) uint4 s_in[YC_TILE_BLOCKS * 4];
  __shared__ __align__(16) uint8_t s_out[YC_TILE_BLOCKS * YOUTH_CODEC_MAX_BLOCK_BYTES + 16];
  __shared__ uint32_t s_wsum[YC_WARPS];
  __shared__ uint32_t s_tile, s_excl;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(P.ticket, 1u);
  __syncthreads();
  const int frame = (int)(s_tile / (uint32_t)P.tiles), t = (int)(s_tile % (uint32_t)P.tiles);
  const uint32_t b0 = (uint32_t)t * YC_TILE_BLOCKS;
  const uint16_t* src = P.in + (size_t)frame * P.npix;
  uint8_t* dst_frame = P.out + (size_t)frame * P.out_stride;
  yc_load_tile(s_in, src, (int)(b0 * 32u), P.npix, P.vec_ok, tid);
  __syncthreads();

  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 q = s_in[yc_chunk(tid, k)];
    w[4 * k] = q.x;
    w[4 * k + 1] = q.y;
    w[4 * k + 2] = q.z;
    w[4 * k + 3] = q.w;
  }
  /* pass 1: mask, first reading, width of the widest code */
  uint32_t mask = 0, first = 0, prev = 0, zor = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const uint32_t v = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu;
    const bool nzv = v != 0u;
    const int d = (int)(int16_t)(uint16_t)(v - prev);
    const uint32_t z = (uint32_t)((d << 1) ^ (d >> 15)) & 0xffffu;
    zor |= (nzv && mask != 0u) ? z : 0u;
    first = (nzv && mask == 0u) ? v : first;
    prev = nzv ? v : prev;
    mask |= (nzv ? 1u : 0u) << i;
  }
  const uint32_t bits = zor ? 32u - (uint32_t)__clz(zor) : 0u;
  const uint32_t nbytes = mask ? ((uint32_t)(__popc(mask) - 1) * bits + 7u) >> 3 : 0u;
  const bool in_frame = b0 + (uint32_t)tid < P.nb;
  const uint32_t sz = in_frame ? (mask ? 7u + nbytes : 4u) : 0u;
  if (in_frame) dst_frame[YOUTH_CODEC_HEADER_BYTES + b0 + tid] = (uint8_t)sz; /* block-size table */
  uint32_t total;
  const uint32_t off = yc_block_scan(sz, s_wsum, tid, &total);

  /* pass 2: the block's bytes into the staging buffer */
  if (in_frame) {
    uint8_t* o = s_out + off;
    o[0] = (uint8_t)mask;
    o[1] = (uint8_t)(mask >> 8);
    o[2] = (uint8_t)(mask >> 16);
    o[3] = (uint8_t)(mask >> 24);
    if (mask) {
      o[4] = (uint8_t)first;
      o[5] = (uint8_t)(first >> 8);
      o[6] = (uint8_t)bits;
      o += 7;
      if (bits) {
        unsigned long long acc = 0ull;
        uint32_t nb = 0;
        bool seen = false;
        prev = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const uint32_t v = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu;
          if (v != 0u) {
            if (seen) {
              const int d = (int)(int16_t)(uint16_t)(v - prev);
              const uint32_t z = (uint32_t)((d << 1) ^ (d >> 15)) & 0xffffu;
              acc |= (unsigned long long)z << nb;
              nb += bits;
              if (nb >= 8u) { /* nb < 24 here: at most two whole bytes are ready */
                *o++ = (uint8_t)acc;
                acc >>= 8;
                nb -= 8u;
              }
              if (nb >= 8u) {
                *o++ = (uint8_t)acc;
                acc >>= 8;
                nb -= 8u;
              }
            }
            seen = true;
            prev = v;
          }
        }
        if (nb) *o = (uint8_t)acc;
      }
    }
  }

  /* where does this tile's payload start?  Decoupled look-back over the tiles of this frame:
   * descriptor = status << 32 | bytes, status 1 = tile aggregate, 2 = inclusive prefix. */
  if (warp == 0) {
    unsigned long long* desc = P.desc + (size_t)frame * P.tiles;
    uint32_t excl = 0;
    if (t > 0) {
      if (lane == 0) yc_st_volatile(desc + t, (1ull << 32) | total);
      int j = t - 1;
      bool done = false;
      while (!done) {
        const int idx = j - lane;
        unsigned long long dsc = 2ull << 32;
        uint32_t spins = 0;
        if (idx >= 0) {
          dsc = yc_ld_volatile(desc + idx);
          while ((dsc >> 32) == 0ull && spins < YC_SPIN_LIMIT) {
            __nanosleep(20);
            dsc = yc_ld_volatile(desc + idx);
            ++spins;
          }
        }
        if (__any_sync(0xffffffffu, spins >= YC_SPIN_LIMIT)) { /* never expected; do not hang the device */
          if (lane == 0) atomicOr(P.err, YC_ERR_SPIN);
          break;
        }
        const uint32_t pm = __ballot_sync(0xffffffffu, (dsc >> 32) == 2ull);
        const int stop = pm ? __ffs(pm) - 1 : 31;
        uint32_t val = (lane <= stop) ? (uint32_t)dsc : 0u;
        val = __reduce_add_sync(0xffffffffu, val);
        excl += val;
        if (pm) done = true;
        else j -= 32;
      }
    }
    if (lane == 0) {
      yc_st_volatile(desc + t, (2ull << 32) | (unsigned long long)(excl + total));
      s_excl = excl;
    }
  }
  __syncthreads(); /* staging complete, s_excl known */

  /* copy out: head bytes up to 4-byte alignment, aligned words through a funnel shift of the
   * (differently aligned) staging buffer, tail bytes */
  {
    uint8_t* dst = dst_frame + YOUTH_CODEC_HEADER_BYTES + P.nb + s_excl;
    uint32_t head = (uint32_t)((4u - ((uint32_t)(uintptr_t)dst & 3u)) & 3u);
    if (head > total) head = total;
    if ((uint32_t)tid < head) dst[tid] = s_out[tid];
    const uint32_t nwords = (total - head) >> 2;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_out);
    uint32_t* dstw = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t k = tid; k < nwords; k += YC_THREADS) dstw[k] = yc_rd32(sw, head + 4u * k);
    const uint32_t done_bytes = head + 4u * nwords;
    if ((uint32_t)tid < total - done_bytes) dst[done_bytes + tid] = s_out[done_bytes + tid];
  }
  if (t == P.tiles - 1 && tid == 0) { /* the last tile knows the payload size: frame header */
    const uint32_t pay = s_excl + total;
    uint32_t* hd = reinterpret_cast<uint32_t*>(dst_frame); /* frames start 16-byte aligned */
    hd[0] = YOUTH_CODEC_MAGIC;
    hd[1] = (uint32_t)P.width | ((uint32_t)P.height << 16);
    hd[2] = P.nb;
    hd[3] = pay;
    P.frame_bytes[frame] = YOUTH_CODEC_HEADER_BYTES + P.nb + pay;
  }
}

__global__ void __launch_bounds__(YC_THREADS) k_yd16_tilesum(const __grid_constant__ YcDecParams P) {
  __shared__ uint32_t s_part[YC_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.x, frame = blockIdx.y;
  const uint8_t* base = P.in + P.offsets[frame];
  const uint32_t gb = (uint32_t)t * YC_TILE_BLOCKS + tid;
  uint32_t v = gb < P.nb ? base[YOUTH_CODEC_HEADER_BYTES + gb] : 0u;
  v = __reduce_add_sync(0xffffffffu, v);
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (tid == 0) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < YC_WARPS; ++w) s += s_part[w];
    P.tile_sum[(size_t)frame * P.tiles + t] = s;
  }
}

#define YC_STAGE_WORDS ((YC_TILE_BLOCKS * YOUTH_CODEC_MAX_BLOCK_BYTES + 3) / 4 + 4)

/* One THREAD per block: the tile's payload is staged in shared memory with aligned 32-bit loads, every
 * thread walks its block's codes with unaligned 32-bit reads (funnel shift of two staged words) and a
 * running sum, the unpacked tile leaves through the swizzled shared tile with 128-bit stores. */
__global__ void __launch_bounds__(YC_THREADS) k_yd16_decode(const __grid_constant__ YcDecParams P) {
  __shared__ __align__(16) uint4 s_px[YC_TILE_BLOCKS * 4];
  __shared__ uint32_t s_stage[YC_STAGE_WORDS];
  __shared__ uint32_t s_wsum[YC_WARPS];
  __shared__ uint32_t s_base;
  __shared__ int s_bad;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = blockIdx.x, frame = blockIdx.y;
  const uint8_t* base = P.in + P.offsets[frame];
  const unsigned long long len = P.offsets[frame + 1] - P.offsets[frame];
  const uint32_t b0 = (uint32_t)t * YC_TILE_BLOCKS;
  const uint32_t* tsum = P.tile_sum + (size_t)frame * P.tiles;
  if (tid == 0) s_bad = 0;

  /* header check (every tile reads it: 16 bytes, L2-resident) */
  uint32_t pay_bytes = 0;
  {
    uint32_t hd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) /* streams are only byte aligned */
      hd[k] = (uint32_t)base[4 * k] | ((uint32_t)base[4 * k + 1] << 8) | ((uint32_t)base[4 * k + 2] << 16) |
              ((uint32_t)base[4 * k + 3] << 24);
    pay_bytes = hd[3];
    const bool ok = hd[0] == YOUTH_CODEC_MAGIC && hd[1] == ((uint32_t)P.width | ((uint32_t)P.height << 16)) &&
                    hd[2] == P.nb && len == (unsigned long long)YOUTH_CODEC_HEADER_BYTES + P.nb + pay_bytes;
    if (!ok) { /* uniform over the CTA; the host checks the same fields first */
      if (tid == 0) atomicOr(P.err, YC_ERR_HEADER);
      return;
    }
  }
  /* payload start of this tile = sum of the earlier tiles' sums */
  if (warp == 0) {
    uint32_t s = 0, all = 0;
    for (int k = lane; k < P.tiles; k += 32) {
      const uint32_t v = tsum[k];
      all += v;
      if (k < t) s += v;
    }
    s = __reduce_add_sync(0xffffffffu, s);
    all = __reduce_add_sync(0xffffffffu, all);
    if (lane == 0) {
      s_base = s;
      if (all != pay_bytes) {
        atomicOr(P.err, YC_ERR_SIZES);
        s_bad = 1;
      }
    }
  }
  /* block offsets inside the tile: exclusive scan of the 256 sizes */
  const bool in_frame = b0 + (uint32_t)tid < P.nb;
  const uint32_t sz = in_frame ? base[YOUTH_CODEC_HEADER_BYTES + b0 + tid] : 0u;
  uint32_t total;
  const uint32_t off = yc_block_scan(sz, s_wsum, tid, &total); /* contains a __syncthreads: s_base / s_bad are visible */
  if (s_bad) return;
  if (total > (uint32_t)(YC_TILE_BLOCKS * YOUTH_CODEC_MAX_BLOCK_BYTES)) { /* a legal tile never is */
    if (tid == 0) atomicOr(P.err, YC_ERR_SIZES);
    return;
  }
  /* stage the tile's payload with aligned 32-bit loads (the device copy has slack behind the last stream) */
  const uint8_t* g0 = base + YOUTH_CODEC_HEADER_BYTES + P.nb + s_base;
  const uint32_t shift = (uint32_t)(uintptr_t)g0 & 3u;
  const uint32_t* a0 = reinterpret_cast<const uint32_t*>(g0 - shift);
  const uint32_t nwords = (shift + total + 3u) >> 2;
  for (uint32_t k = tid; k < nwords + 2u; k += YC_THREADS) s_stage[k] = k < nwords ? __ldg(a0 + k) : 0u;
  __syncthreads();

  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = 0u;
  if (in_frame) {
    const uint32_t o = shift + off;
    bool bad = sz < 4u;
    if (!bad) {
      const uint32_t mask = yc_rd32(s_stage, o);
      if (mask == 0u) {
        bad = sz != 4u;
      } else if (sz < 7u) {
        bad = true;
      } else {
        const uint32_t fb = yc_rd32(s_stage, o + 4u);
        const uint32_t first = fb & 0xffffu, bits = (fb >> 16) & 0xffu;
        bad = bits > 16u || sz != 7u + (((uint32_t)(__popc(mask) - 1) * bits + 7u) >> 3);
        if (!bad) {
          const uint32_t vmask = (1u << bits) - 1u;
          uint32_t cur = first, bitpos = (o + 7u) * 8u;
          bool seen = false;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((mask >> i) & 1u) {
              if (seen && bits) {
                const uint32_t x = yc_rd32(s_stage, bitpos >> 3);
                const uint32_t z = (x >> (bitpos & 7u)) & vmask;
                cur = (cur + ((z >> 1) ^ (0u - (z & 1u)))) & 0xffffu; /* un-zig-zag, modulo 2^16 */
                bitpos += bits;
              }
              seen = true;
              w[i >> 1] |= cur << (16 * (i & 1));
            }
          }
        }
      }
    }
    if (bad) atomicOr(P.err, YC_ERR_BLOCK);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) s_px[yc_chunk(tid, k)] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
  __syncthreads();
  uint16_t* dst = P.out + (size_t)frame * P.npix;
  const int p0 = (int)(b0 * 32u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tid + YC_THREADS * i;
    const int px = p0 + 8 * c;
    const uint4 q = s_px[yc_chunk(c >> 2, c & 3)];
    if (P.vec_ok && px + 8 <= P.npix) {
      *reinterpret_cast<uint4*>(dst + px) = q;
    } else {
      const uint32_t h[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (px + j < P.npix) dst[px + j] = (uint16_t)(h[j >> 1] >> (16 * (j & 1)));
    }
  }
}

/* ------------------------------------------------------------------ host side: context + C ABI
 * (uses fail() / CU() of youth_cuda.cu, which includes this file) */

struct youth_codec {
  int width, height, npix, device, max_frames;
  uint32_t nb;
  int tiles;
  size_t stride; /* device stride between packed frames: youth_codec_max_bytes rounded up to 16 */
  cudaStream_t stream;
  uint16_t* d_raw;       /* [max_frames][npix] */
  uint8_t* d_packed;     /* [max_frames][stride] + slack */
  unsigned long long* d_desc;
  unsigned int* d_ticket;
  uint32_t* d_frame_bytes;
  unsigned int* d_err;
  unsigned long long* d_offsets; /* [max_frames + 1] */
  uint32_t* d_tile_sum;
  uint32_t* h_frame_bytes; /* pinned */
  unsigned int* h_err;     /* pinned */
  unsigned long long* h_offsets; /* pinned */
  cudaEvent_t e0, e1;
  float last_ms;
  uint64_t launches;
};

extern "C" size_t youth_codec_max_bytes(int width, int height) {
  if (width <= 0 || height <= 0) return 0;
  const size_t nb = ((size_t)width * height + YOUTH_CODEC_BLOCK_PIXELS - 1) / YOUTH_CODEC_BLOCK_PIXELS;
  return YOUTH_CODEC_HEADER_BYTES + nb + nb * YOUTH_CODEC_MAX_BLOCK_BYTES;
}

extern "C" void youth_codec_destroy(youth_codec* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_raw);
  cudaFree(c->d_packed);
  cudaFree(c->d_desc);
  cudaFree(c->d_ticket);
  cudaFree(c->d_frame_bytes);
  cudaFree(c->d_err);
  cudaFree(c->d_offsets);
  cudaFree(c->d_tile_sum);
  if (c->h_frame_bytes) cudaFreeHost(c->h_frame_bytes);
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->h_offsets) cudaFreeHost(c->h_offsets);
  if (c->e0) cudaEventDestroy(c->e0);
  if (c->e1) cudaEventDestroy(c->e1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

static int codec_init(youth_codec* c, int width, int height, int max_frames, int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("no CUDA device available (%s); the YD16 codec has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev) return fail("device %d out of range (have %d)", device, ndev);
  CU(cudaSetDevice(device));
  c->width = width;
  c->height = height;
  c->npix = width * height;
  c->device = device;
  c->max_frames = max_frames;
  c->nb = (uint32_t)((c->npix + YOUTH_CODEC_BLOCK_PIXELS - 1) / YOUTH_CODEC_BLOCK_PIXELS);
  c->tiles = (int)((c->nb + YC_TILE_BLOCKS - 1) / YC_TILE_BLOCKS);
  c->stride = (youth_codec_max_bytes(width, height) + 15) & ~(size_t)15;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(dalloc(&c->d_raw, (size_t)max_frames * c->npix + 8));
  CU(dalloc(&c->d_packed, (size_t)max_frames * c->stride + 64));
  CU(dalloc(&c->d_desc, (size_t)max_frames * c->tiles));
  CU(dalloc(&c->d_ticket, (size_t)1));
  CU(dalloc(&c->d_frame_bytes, (size_t)max_frames));
  CU(dalloc(&c->d_err, (size_t)1));
  CU(dalloc(&c->d_offsets, (size_t)max_frames + 1));
  CU(dalloc(&c->d_tile_sum, (size_t)max_frames * c->tiles));
  CU(cudaHostAlloc((void**)&c->h_frame_bytes, sizeof(uint32_t) * max_frames, cudaHostAllocDefault));
  CU(cudaHostAlloc((void**)&c->h_err, sizeof(unsigned int), cudaHostAllocDefault));
  CU(cudaHostAlloc((void**)&c->h_offsets, sizeof(unsigned long long) * ((size_t)max_frames + 1), cudaHostAllocDefault));
  CU(cudaEventCreate(&c->e0));
  CU(cudaEventCreate(&c->e1));
  return 1;
}

extern "C" int youth_codec_create(int width, int height, int max_frames, int device, youth_codec** out) {
  if (!out) return fail("null argument");
  *out = NULL;
  if (width <= 0 || height <= 0 || width > 65535 || height > 65535) return fail("width/height must be 1..65535");
  if ((size_t)width * height > ((size_t)1 << 30)) return fail("frame too large");
  if (max_frames < 1 || max_frames > 65535) return fail("max_frames must be 1..65535");
  youth_codec* c = new youth_codec();
  memset((void*)c, 0, sizeof(*c));
  c->device = device;
  if (!codec_init(c, width, height, max_frames, device)) {
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    youth_codec_destroy(c);
    memcpy(g_err, keep, sizeof(keep));
    return 0;
  }
  *out = c;
  return 1;
}

/* enqueue the encoder for n frames at d_in (device) into c->d_packed (fixed stride) */
static int codec_enqueue_encode(youth_codec* c, cudaStream_t st, const uint16_t* d_in, int n) {
  YcEncParams p;
  memset(&p, 0, sizeof(p));
  p.in = d_in;
  p.out = c->d_packed;
  p.out_stride = c->stride;
  p.width = c->width;
  p.height = c->height;
  p.npix = c->npix;
  p.nb = c->nb;
  p.tiles = c->tiles;
  p.n_frames = n;
  p.vec_ok = (c->npix % 8 == 0) && (((uintptr_t)d_in & 15) == 0);
  p.desc = c->d_desc;
  p.ticket = c->d_ticket;
  p.frame_bytes = c->d_frame_bytes;
  p.err = c->d_err;
  CU(cudaMemsetAsync(c->d_desc, 0, sizeof(unsigned long long) * (size_t)n * c->tiles, st));
  CU(cudaMemsetAsync(c->d_ticket, 0, sizeof(unsigned int), st));
  CU(cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), st));
  CU(cudaEventRecord(c->e0, st));
  k_yd16_encode<<<(unsigned)(n * c->tiles), YC_THREADS, 0, st>>>(p);
  CU(cudaEventRecord(c->e1, st));
  CU(cudaGetLastError());
  c->launches += 1;
  return 1;
}

/* enqueue the decoder: n streams inside d_in at d_offsets[0..n] -> n raw frames at d_out.
 * Clears nothing: the caller zeroes c->d_err once per group and reads it afterwards. */
static int codec_enqueue_decode(youth_codec* c, cudaStream_t st, const uint8_t* d_in, const unsigned long long* d_offsets,
                                uint32_t* d_tile_sum, int n, uint16_t* d_out) {
  YcDecParams p;
  memset(&p, 0, sizeof(p));
  p.in = d_in;
  p.offsets = d_offsets;
  p.out = d_out;
  p.width = c->width;
  p.height = c->height;
  p.npix = c->npix;
  p.nb = c->nb;
  p.tiles = c->tiles;
  p.n_frames = n;
  p.vec_ok = (c->npix % 8 == 0) && (((uintptr_t)d_out & 15) == 0);
  p.tile_sum = d_tile_sum;
  p.err = c->d_err;
  const dim3 grid((unsigned)c->tiles, (unsigned)n);
  k_yd16_tilesum<<<grid, YC_THREADS, 0, st>>>(p);
  k_yd16_decode<<<grid, YC_THREADS, 0, st>>>(p);
  CU(cudaGetLastError());
  c->launches += 2;
  return 1;
}

/* host-side header check of one stream (the kernels re-check what they rely on) */
static int codec_check_header(const youth_codec* c, const uint8_t* s, unsigned long long len, int i) {
  if (len < YOUTH_CODEC_HEADER_BYTES) return fail("stream %d: shorter than a header", i);
  uint32_t magic, nb, pay;
  uint16_t w, hh;
  memcpy(&magic, s, 4);
  memcpy(&w, s + 4, 2);
  memcpy(&hh, s + 6, 2);
  memcpy(&nb, s + 8, 4);
  memcpy(&pay, s + 12, 4);
  if (magic != YOUTH_CODEC_MAGIC) return fail("stream %d: bad magic", i);
  if (w != c->width || hh != c->height) return fail("stream %d: is %ux%u, codec is %dx%d", i, w, hh, c->width, c->height);
  if (nb != c->nb) return fail("stream %d: block count %u, expected %u", i, nb, c->nb);
  if (len != (unsigned long long)YOUTH_CODEC_HEADER_BYTES + nb + pay) return fail("stream %d: length does not match its header", i);
  if (len > youth_codec_max_bytes(c->width, c->height)) return fail("stream %d: longer than the worst case", i);
  return 1;
}

extern "C" int youth_codec_encode(youth_codec* c, const uint16_t* depth, int mem_kind, int n_frames, uint8_t* out,
                                  size_t out_capacity, uint64_t* offsets_out) {
  if (!c || !depth || !out || !offsets_out) return fail("null argument");
  if (n_frames < 1 || n_frames > c->max_frames) return fail("n_frames must be 1..%d", c->max_frames);
  CU(cudaSetDevice(c->device));
  const uint16_t* d_in = depth;
  if (mem_kind == YOUTH_MEM_HOST || mem_kind == YOUTH_MEM_HOST_PINNED) {
    CU(cudaMemcpyAsync(c->d_raw, depth, sizeof(uint16_t) * (size_t)n_frames * c->npix, cudaMemcpyHostToDevice, c->stream));
    d_in = c->d_raw;
  } else if (mem_kind != YOUTH_MEM_DEVICE) {
    return fail("unknown mem_kind %d", mem_kind);
  }
  if (!codec_enqueue_encode(c, c->stream, d_in, n_frames)) return 0;
  CU(cudaMemcpyAsync(c->h_frame_bytes, c->d_frame_bytes, sizeof(uint32_t) * n_frames, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->last_ms, c->e0, c->e1));
  if (*c->h_err) return fail("YD16 encoder reported device error 0x%x", *c->h_err);
  offsets_out[0] = 0;
  for (int i = 0; i < n_frames; ++i) offsets_out[i + 1] = offsets_out[i] + c->h_frame_bytes[i];
  if (offsets_out[n_frames] > out_capacity)
    return fail("output buffer too small: %llu bytes needed, %zu given", (unsigned long long)offsets_out[n_frames], out_capacity);
  for (int i = 0; i < n_frames; ++i)
    CU(cudaMemcpyAsync(out + offsets_out[i], c->d_packed + (size_t)i * c->stride, c->h_frame_bytes[i],
                       cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 1;
}

extern "C" int youth_codec_decode(youth_codec* c, const uint8_t* in, const uint64_t* offsets, int n_frames,
                                  uint16_t* depth_out, int mem_kind) {
  if (!c || !in || !offsets || !depth_out) return fail("null argument");
  if (n_frames < 1 || n_frames > c->max_frames) return fail("n_frames must be 1..%d", c->max_frames);
  if (mem_kind != YOUTH_MEM_HOST && mem_kind != YOUTH_MEM_HOST_PINNED && mem_kind != YOUTH_MEM_DEVICE)
    return fail("unknown mem_kind %d", mem_kind);
  for (int i = 0; i < n_frames; ++i) {
    if (offsets[i + 1] < offsets[i]) return fail("offsets must be non-decreasing");
    if (!codec_check_header(c, in + offsets[i], offsets[i + 1] - offsets[i], i)) return 0;
  }
  CU(cudaSetDevice(c->device));
  const uint64_t first = offsets[0], total = offsets[n_frames] - first;
  if (total > (uint64_t)c->max_frames * c->stride) return fail("streams larger than the context");
  for (int i = 0; i <= n_frames; ++i) c->h_offsets[i] = offsets[i] - first;
  CU(cudaMemcpyAsync(c->d_packed, in + first, total, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_offsets, c->h_offsets, sizeof(unsigned long long) * ((size_t)n_frames + 1), cudaMemcpyHostToDevice,
                     c->stream));
  CU(cudaMemsetAsync(c->d_err, 0, sizeof(unsigned int), c->stream));
  uint16_t* d_out = mem_kind == YOUTH_MEM_DEVICE ? depth_out : c->d_raw;
  CU(cudaEventRecord(c->e0, c->stream));
  if (!codec_enqueue_decode(c, c->stream, c->d_packed, c->d_offsets, c->d_tile_sum, n_frames, d_out)) return 0;
  CU(cudaEventRecord(c->e1, c->stream));
  CU(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
  if (mem_kind != YOUTH_MEM_DEVICE)
    CU(cudaMemcpyAsync(depth_out, c->d_raw, sizeof(uint16_t) * (size_t)n_frames * c->npix, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->last_ms, c->e0, c->e1));
  if (*c->h_err) return fail("malformed YD16 stream (device check 0x%x)", *c->h_err);
  return 1;
}

extern "C" float youth_codec_last_kernel_ms(const youth_codec* c) { return c ? c->last_ms : 0.0f; }
extern "C" uint64_t youth_codec_launch_count(const youth_codec* c) { return c ? c->launches : 0; }
