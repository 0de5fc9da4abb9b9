/*
 * youth_codec.h -- lossless GPU codec for 16-bit depth frames ("YD16"), the packed payload of
 * .bin records with frameType = FRAME_TYPE_DEPTH_PACKED.  Part of libyouth_cuda.so.
 *
 * Why: the reference records raw frames -- 614 400 B of depth per 640x480 frame, fwrite + fflush
 * per frame (Youth.Source/LoggingModule/loggingModule.c:118-127) -- and replays them with
 * readFrameFromFile (loggingModule.c:404-444); at the tracker's rate that is several GB/s of disk
 * and PCIe traffic.  Depth is smooth, so delta + bit-packing gets
 * about 3x without loss.  What each entry point stands next to in the reference:
 *
 *   youth_codec_encode           <- the depth fwrite of saveFrameToFile     loggingModule.c:118-121
 *   youth_codec_decode           <- the depth fread of readFrameFromFile    loggingModule.c:430-434
 *   youth_cuda_track_batch_packed<- processSlamFrame fed from a recording   SLAM.cpp:126-175
 *
 * Stream layout of ONE frame (little endian, byte granular):
 *   bytes 0..3   'Y' 'D' '1' '6'
 *   bytes 4..5   width   (uint16)      bytes 6..7  height (uint16)
 *   bytes 8..11  nblocks (uint32) = ceil(width*height / 32)
 *   bytes 12..15 payload bytes (uint32)
 *   nblocks x uint8   size of each block's payload (4, or 7..69)
 *   payload           blocks back to back
 * A block covers 32 consecutive pixels of the row-major frame (the last block is padded with 0):
 *   uint32 mask            bit i set = pixel i is non-zero
 *   if mask != 0:
 *     uint16 first         value of the first non-zero pixel
 *     uint8  b             bits per delta, 0..16 (the smallest width that holds every code)
 *     ceil((popcount(mask)-1) * b / 8) bytes: for the 2nd, 3rd, ... non-zero pixel the zig-zag code of
 *                          (int16)(value - previous non-zero value), b bits each, LSB first,
 *                          unused high bits of the last byte zero
 * Zero pixels (no reading) cost one mask bit and never disturb the deltas of their neighbours.
 * The encoding of a frame is unique (canonical), so "parity" is byte equality of streams.
 * A decoder rejects a stream whose header, size table or block headers are inconsistent.
 *
 * Return convention as in youth_cuda.h: 1 = success, 0 = failure (youth_cuda_last_error()).
 * No CPU fallback: every call fails without a CUDA device.
 */
#ifndef YOUTH_CODEC_H
#define YOUTH_CODEC_H

#include <stddef.h>
#include <stdint.h>

#include "youth_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define YOUTH_CODEC_MAGIC 0x36314459u /* "YD16" */
#define YOUTH_CODEC_BLOCK_PIXELS 32
#define YOUTH_CODEC_HEADER_BYTES 16
#define YOUTH_CODEC_MAX_BLOCK_BYTES 69 /* 4 + 2 + 1 + ceil(31*16/8) */
#define FRAME_TYPE_DEPTH_PACKED 2      /* FrameHeader.frameType of a record whose depth payload is YD16 */

typedef struct youth_codec youth_codec;

/* worst-case size of one packed frame */
size_t youth_codec_max_bytes(int width, int height);

/* a codec context sized for up to max_frames frames of width x height per call, on CUDA `device` */
int youth_codec_create(int width, int height, int max_frames, int device, youth_codec** out);
void youth_codec_destroy(youth_codec* c);

/* Pack n_frames tightly packed uint16 frames (host or device memory per mem_kind: YOUTH_MEM_HOST /
 * YOUTH_MEM_HOST_PINNED / YOUTH_MEM_DEVICE).  The n streams are written back to back into the
 * host buffer `out` (capacity out_capacity bytes; n_frames * youth_codec_max_bytes() always
 * suffices): stream i occupies out[offsets_out[i] .. offsets_out[i+1]).  offsets_out has
 * n_frames + 1 entries.  Blocking. */
int youth_codec_encode(youth_codec* c, const uint16_t* depth, int mem_kind, int n_frames, uint8_t* out,
                       size_t out_capacity, uint64_t* offsets_out);

/* Unpack n_frames streams laid out as above (host memory) into depth_out (host or device per
 * mem_kind), n_frames tightly packed frames.  Fails, leaving depth_out unspecified, when any stream
 * is malformed.  Blocking. */
int youth_codec_decode(youth_codec* c, const uint8_t* in, const uint64_t* offsets, int n_frames,
                       uint16_t* depth_out, int mem_kind);

/* device milliseconds spent in the kernels of the last encode / decode call (CUDA events on the
 * codec's stream), and the number of kernels that call launched */
float youth_codec_last_kernel_ms(const youth_codec* c);
uint64_t youth_codec_launch_count(const youth_codec* c);

/* Tracker fed from packed recordings: like youth_cuda_track_batch with YOUTH_MEM_HOST_PINNED /
 * YOUTH_MEM_HOST inputs, but streams[s] points at the n_frames back-to-back YD16 streams of
 * sequence s and offsets[s] at their n_frames + 1 offsets.  The packed bytes are copied to the
 * device and unpacked there, straight into the tracker's raw landing zone (a third of the PCIe
 * traffic of raw frames).  Host buffers must stay valid until the call returns (it returns after
 * the copies were enqueued from pageable memory, or when poses_out is non-NULL after the
 * group finished; pinned inputs follow the YOUTH_MEM_HOST_PINNED rule).  On a frame-to-model handle
 * (youth_model.h) the group is unpacked as a whole and then tracked frame after frame. */
int youth_cuda_track_batch_packed(youth_cuda_handle* h, const uint8_t* const* streams,
                                  const uint64_t* const* offsets, int n_frames, int mem_kind,
                                  const uint32_t* timestamps_ms, float* poses_out);

#ifdef __cplusplus
}
#endif
#endif
