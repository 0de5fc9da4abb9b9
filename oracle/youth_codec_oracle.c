/*
 * youth_codec_oracle.c -- CPU statement of the YD16 depth codec (include/youth_codec.h).
 * TEST INFRASTRUCTURE ONLY (same rules as youth_oracle.c).  The reference has no codec -- it
 * records raw frames (Youth.Source/LoggingModule/loggingModule.c:118-127) -- so this file defines
 * the format; "parity" for this row means: the device encoder emits byte-identical streams and
 * both decoders return the original frame.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define YC_BLOCK 32

static uint32_t zigzag16(int16_t d) { return (uint32_t)(uint16_t)((d << 1) ^ (d >> 15)); }
static int16_t unzigzag16(uint32_t z) { return (int16_t)((z >> 1) ^ (uint32_t)(-(int32_t)(z & 1u))); }

size_t yc_max_bytes(int width, int height) {
  const size_t nb = ((size_t)width * height + YC_BLOCK - 1) / YC_BLOCK;
  return 16 + nb + nb * 69;
}

/* returns the stream length */
size_t yc_encode(const uint16_t* depth, int width, int height, uint8_t* out) {
  const size_t npix = (size_t)width * height;
  const uint32_t nb = (uint32_t)((npix + YC_BLOCK - 1) / YC_BLOCK);
  uint8_t* sizes = out + 16;
  uint8_t* pay = sizes + nb;
  size_t off = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    uint16_t v[YC_BLOCK];
    for (int i = 0; i < YC_BLOCK; ++i) {
      const size_t p = (size_t)b * YC_BLOCK + i;
      v[i] = p < npix ? depth[p] : 0;
    }
    uint32_t mask = 0;
    for (int i = 0; i < YC_BLOCK; ++i)
      if (v[i]) mask |= 1u << i;
    uint8_t* o = pay + off;
    memcpy(o, &mask, 4);
    size_t sz = 4;
    if (mask) {
      uint32_t zz[YC_BLOCK];
      int n = 0, have = 0;
      uint16_t prev = 0, first = 0;
      uint32_t zmax = 0;
      for (int i = 0; i < YC_BLOCK; ++i) {
        if (!v[i]) continue;
        if (!have) {
          first = v[i];
          have = 1;
        } else {
          zz[n] = zigzag16((int16_t)(uint16_t)(v[i] - prev));
          if (zz[n] > zmax) zmax = zz[n];
          ++n;
        }
        prev = v[i];
      }
      int bits = 0;
      while (zmax >> bits) ++bits;
      memcpy(o + 4, &first, 2);
      o[6] = (uint8_t)bits;
      const size_t nbytes = ((size_t)n * bits + 7) / 8;
      memset(o + 7, 0, nbytes);
      for (int k = 0; k < n; ++k) {
        const size_t bo = (size_t)k * bits;
        for (int t = 0; t < bits; ++t)
          if ((zz[k] >> t) & 1u) o[7 + ((bo + t) >> 3)] |= (uint8_t)(1u << ((bo + t) & 7));
      }
      sz = 7 + nbytes;
    }
    sizes[b] = (uint8_t)sz;
    off += sz;
  }
  const uint32_t magic = 0x36314459u, pay_bytes = (uint32_t)off;
  const uint16_t w16 = (uint16_t)width, h16 = (uint16_t)height;
  memcpy(out, &magic, 4);
  memcpy(out + 4, &w16, 2);
  memcpy(out + 6, &h16, 2);
  memcpy(out + 8, &nb, 4);
  memcpy(out + 12, &pay_bytes, 4);
  return 16 + (size_t)nb + off;
}

/* 1 = ok, 0 = malformed stream */
int yc_decode(const uint8_t* in, size_t len, int width, int height, uint16_t* depth) {
  const size_t npix = (size_t)width * height;
  uint32_t magic, nb, pay_bytes;
  uint16_t w16, h16;
  if (len < 16) return 0;
  memcpy(&magic, in, 4);
  memcpy(&w16, in + 4, 2);
  memcpy(&h16, in + 6, 2);
  memcpy(&nb, in + 8, 4);
  memcpy(&pay_bytes, in + 12, 4);
  if (magic != 0x36314459u || w16 != width || h16 != height) return 0;
  if (nb != (uint32_t)((npix + YC_BLOCK - 1) / YC_BLOCK) || len != 16 + (size_t)nb + pay_bytes) return 0;
  const uint8_t* sizes = in + 16;
  const uint8_t* pay = sizes + nb;
  size_t off = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    const uint8_t* o = pay + off;
    if (off + sizes[b] > pay_bytes || sizes[b] < 4) return 0;
    uint32_t mask;
    memcpy(&mask, o, 4);
    uint16_t first = 0;
    int bits = 0;
    if (mask) {
      if (sizes[b] < 7) return 0;
      memcpy(&first, o + 4, 2);
      bits = o[6];
      if (bits > 16) return 0;
      if (sizes[b] != 7 + ((size_t)(__builtin_popcount(mask) - 1) * bits + 7) / 8) return 0;
    } else if (sizes[b] != 4) {
      return 0;
    }
    uint16_t cur = first;
    int k = -1;
    for (int i = 0; i < YC_BLOCK; ++i) {
      const size_t p = (size_t)b * YC_BLOCK + i;
      uint16_t val = 0;
      if ((mask >> i) & 1u) {
        if (k >= 0) {
          uint32_t z = 0;
          const size_t bo = (size_t)k * bits;
          for (int t = 0; t < bits; ++t)
            if ((o[7 + ((bo + t) >> 3)] >> ((bo + t) & 7)) & 1u) z |= 1u << t;
          cur = (uint16_t)(cur + (uint16_t)unzigzag16(z));
        }
        ++k;
        val = cur;
      }
      if (p < npix) depth[p] = val;
    }
    off += sizes[b];
  }
  return off == pay_bytes;
}
