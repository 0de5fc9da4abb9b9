/*
 * youth_frameio.c -- .bin record reader/writer and the mq chunk protocol.
 *
 * Record format (reference Youth.Source/LoggingModule/loggingModule.c:101-130 writer,
 * :404-444 reader): FrameHeader (28 B, native endian, raw fwrite) + depth payload +
 * colour payload; an optional terminating header with frameType = FRAME_TYPE_END_OF_FILE
 * (loggingModule.c:223-226).  Unlike the reference reader, whose caller caps payloads at
 * 1 MiB (loggingModule.c:528, :424-427) and therefore cannot replay 1280x960, the caps
 * here are whatever the caller's buffers hold.
 *
 * Chunk protocol (reference loggingModule.c:447-485, sensorModule.c:149-210): each mq
 * message = MessageHeader (292 B) + up to 7900 payload bytes; reassembly offset is
 * chunkIndex * 7900 (loggingModule.c:313); a frame is complete when the last depth chunk
 * AND the last colour chunk have been seen (loggingModule.c:354).
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <mqueue.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "youth_host.h"

static int write_record(FILE* f, uint32_t frame_id, uint32_t timestamp_ms, int width, int height, int frame_type,
                        const void* depth, uint32_t depth_bytes, const uint8_t* color) {
  if (!f || !depth || width <= 0 || height <= 0 || width > 65535 || height > 65535) return 0;
  FrameHeader h;
  memset(&h, 0, sizeof(h)); /* also zeroes the 2 padding bytes the reference leaves undefined */
  h.frameId = frame_id;
  h.timestamp = timestamp_ms;
  h.frameType = (uint16_t)frame_type;
  h.width = (uint16_t)width;
  h.height = (uint16_t)height;
  h.depthDataSize = depth_bytes;
  h.colorDataSize = (uint32_t)((size_t)width * height * 3);
  h.reserved = 0;
  if (fwrite(&h, sizeof(h), 1, f) != 1) return 0;
  if (fwrite(depth, 1, h.depthDataSize, f) != h.depthDataSize) return 0;
  if (color) {
    if (fwrite(color, 1, h.colorDataSize, f) != h.colorDataSize) return 0;
  } else {
    uint8_t row[4096];
    memset(row, 128, sizeof(row));
    size_t left = h.colorDataSize;
    while (left) {
      size_t n = left < sizeof(row) ? left : sizeof(row);
      if (fwrite(row, 1, n, f) != n) return 0;
      left -= n;
    }
  }
  return 1;
}

int youth_bin_write_frame(FILE* f, uint32_t frame_id, uint32_t timestamp_ms, int width, int height,
                          const uint16_t* depth, const uint8_t* color) {
  if (width <= 0 || height <= 0) return 0;
  return write_record(f, frame_id, timestamp_ms, width, height, FRAME_TYPE_DEPTH_COLOR, depth,
                      (uint32_t)((size_t)width * height * sizeof(uint16_t)), color);
}

int youth_bin_write_packed_frame(FILE* f, uint32_t frame_id, uint32_t timestamp_ms, int width, int height,
                                 const uint8_t* yd16, uint32_t yd16_bytes, const uint8_t* color) {
  if (yd16_bytes < YOUTH_CODEC_HEADER_BYTES) return 0;
  return write_record(f, frame_id, timestamp_ms, width, height, FRAME_TYPE_DEPTH_PACKED, yd16, yd16_bytes, color);
}

int youth_bin_write_eof(FILE* f) {
  if (!f) return 0;
  FrameHeader h;
  memset(&h, 0, sizeof(h));
  h.frameType = FRAME_TYPE_END_OF_FILE;
  return fwrite(&h, sizeof(h), 1, f) == 1;
}

int youth_bin_read_header(FILE* f, FrameHeader* hdr) {
  if (!f || !hdr) return 0;
  if (fread(hdr, sizeof(*hdr), 1, f) != 1) return 0;
  return hdr->frameType != FRAME_TYPE_END_OF_FILE;
}

int youth_bin_read_payload(FILE* f, const FrameHeader* hdr, void* depth, size_t depth_cap, void* color, size_t color_cap) {
  if (!f || !hdr || !depth) return 0;
  if (hdr->depthDataSize > depth_cap) return 0;
  if (color && hdr->colorDataSize > color_cap) return 0;
  if (fread(depth, 1, hdr->depthDataSize, f) != hdr->depthDataSize) return 0;
  if (color) {
    if (fread(color, 1, hdr->colorDataSize, f) != hdr->colorDataSize) return 0;
  } else if (hdr->colorDataSize) {
    if (fseek(f, (long)hdr->colorDataSize, SEEK_CUR) != 0) return 0;
  }
  return 1;
}

int youth_bin_read_frame(FILE* f, FrameHeader* hdr, void* depth, size_t depth_cap, void* color,
                         size_t color_cap) {
  if (!depth) return 0;
  return youth_bin_read_header(f, hdr) && youth_bin_read_payload(f, hdr, depth, depth_cap, color, color_cap);
}

/* ------------------------------------------------------------------ chunks */

int youth_chunk_count(size_t bytes) { return (int)YOUTH_CHUNKS_FOR(bytes); }

size_t youth_chunk_build(void* msg, int msg_type, int frame_id, uint32_t timestamp_ms, int width,
                         int height, const void* data, size_t data_bytes, int chunk) {
  MessageHeader h;
  memset(&h, 0, sizeof(h)); /* the reference playback path leaves ctrlCommand/filename as garbage */
  const int total = youth_chunk_count(data_bytes);
  size_t off = (size_t)chunk * YOUTH_CHUNK_PAYLOAD;
  size_t n = 0;
  if (data && chunk >= 0 && chunk < total) {
    n = data_bytes - off;
    if (n > (size_t)YOUTH_CHUNK_PAYLOAD) n = YOUTH_CHUNK_PAYLOAD;
  }
  h.msgType = msg_type;
  h.width = width;
  h.height = height;
  h.chunkIndex = chunk;
  h.totalChunks = total;
  h.dataSize = (int)n;
  h.frameId = frame_id;
  h.timestamp = timestamp_ms;
  memcpy(msg, &h, sizeof(h));
  if (n) memcpy((char*)msg + sizeof(h), (const char*)data + off, n);
  return sizeof(h) + n;
}

size_t youth_pose_msg_build(void* msg, int frame_id, uint32_t timestamp_ms, const float pose[12], uint32_t status,
                            uint32_t inliers) {
  MessageHeader h;
  YouthPoseMsg p;
  memset(&h, 0, sizeof(h));
  h.msgType = MSG_TYPE_POSE;
  h.totalChunks = 1;
  h.dataSize = (int)sizeof(p);
  h.frameId = frame_id;
  h.timestamp = timestamp_ms;
  memcpy(p.pose, pose, sizeof(p.pose));
  p.status = status;
  p.inliers = inliers;
  memcpy(msg, &h, sizeof(h));
  memcpy((char*)msg + sizeof(h), &p, sizeof(p));
  return sizeof(h) + sizeof(p);
}

int youth_pose_msg_parse(const void* msg, size_t len, int* frame_id, uint32_t* timestamp_ms, YouthPoseMsg* out) {
  MessageHeader h;
  if (!msg || !out || len < sizeof(h) + sizeof(*out)) return 0;
  memcpy(&h, msg, sizeof(h));
  if (h.msgType != MSG_TYPE_POSE || h.dataSize != (int)sizeof(*out)) return 0;
  memcpy(out, (const char*)msg + sizeof(h), sizeof(*out));
  if (frame_id) *frame_id = h.frameId;
  if (timestamp_ms) *timestamp_ms = h.timestamp;
  return 1;
}

struct youth_reasm {
  int width, height, frame_id;
  uint32_t timestamp;
  uint16_t* depth;
  uint8_t* color;
  size_t depth_bytes, color_bytes;
  int got_depth, got_color;
};

youth_reasm* youth_reasm_create(void) { return (youth_reasm*)calloc(1, sizeof(youth_reasm)); }

void youth_reasm_destroy(youth_reasm* r) {
  if (!r) return;
  free(r->depth);
  free(r->color);
  free(r);
}

static int reasm_resize(youth_reasm* r, int w, int h) {
  if (w <= 0 || h <= 0 || w > 65535 || h > 65535) return 0;
  if (w == r->width && h == r->height && r->depth) return 1;
  free(r->depth);
  free(r->color);
  r->depth_bytes = (size_t)w * h * 2;
  r->color_bytes = (size_t)w * h * 3;
  r->depth = (uint16_t*)calloc(1, r->depth_bytes);
  r->color = (uint8_t*)calloc(1, r->color_bytes);
  r->width = w;
  r->height = h;
  return r->depth && r->color;
}

int youth_reasm_feed(youth_reasm* r, const void* msg, size_t len) {
  if (!r || !msg || len < sizeof(MessageHeader)) return -1;
  MessageHeader h;
  memcpy(&h, msg, sizeof(h));
  const char* payload = (const char*)msg + sizeof(h);
  if (h.dataSize < 0 || (size_t)h.dataSize > len - sizeof(h)) return -1;
  switch (h.msgType) {
    case MSG_TYPE_METADATA:
      if (!reasm_resize(r, h.width, h.height)) return -1;
      r->frame_id = h.frameId;
      r->timestamp = h.timestamp;
      r->got_depth = r->got_color = 0;
      return 0;
    case MSG_TYPE_DEPTH_DATA:
    case MSG_TYPE_COLOR_DATA: {
      if (!r->depth) return -1;
      const int is_depth = h.msgType == MSG_TYPE_DEPTH_DATA;
      char* dst = is_depth ? (char*)r->depth : (char*)r->color;
      const size_t cap = is_depth ? r->depth_bytes : r->color_bytes;
      if (h.chunkIndex < 0) return -1;
      const size_t off = (size_t)h.chunkIndex * YOUTH_CHUNK_PAYLOAD;
      if (off + (size_t)h.dataSize > cap) return -1;
      memcpy(dst + off, payload, (size_t)h.dataSize);
      r->frame_id = h.frameId;
      r->timestamp = h.timestamp;
      if (h.chunkIndex == h.totalChunks - 1) {
        if (is_depth)
          r->got_depth = 1;
        else
          r->got_color = 1;
      }
      if (r->got_depth && r->got_color) {
        r->got_depth = r->got_color = 0;
        return 1;
      }
      return 0;
    }
    default:
      return 0; /* control traffic is not ours */
  }
}

const uint16_t* youth_reasm_depth(const youth_reasm* r) { return r ? r->depth : NULL; }
const uint8_t* youth_reasm_color(const youth_reasm* r) { return r ? r->color : NULL; }
void youth_reasm_info(const youth_reasm* r, int* w, int* h, int* id, uint32_t* ts) {
  if (!r) return;
  if (w) *w = r->width;
  if (h) *h = r->height;
  if (id) *id = r->frame_id;
  if (ts) *ts = r->timestamp;
}

/* Consumer of a viewer-side queue: the reference feeds MQ_LOGGER_TO_VIEWER from two places, the logger's
 * pass-through (loggingModule.c:284-288) and playbackThread (loggingModule.c:584-590), and its only reader
 * is the viewer's receive loop (viewerModule.c:160-250).  This is that loop for the tracker: receive,
 * reassemble (same completion test as the logger, loggingModule.c:354), hand every whole frame to `sink`
 * (the processSlamFrame signature) before the next message is received, so the reassembly buffer can be
 * reused -- the callee copies, SLAM.cpp:133-134. */
long youth_mq_consume(const char* queue, youth_frame_sink sink, volatile int* stop, int idle_timeout_ms) {
  if (!queue || !sink) return -1;
  mqd_t mq = mq_open(queue, O_RDONLY);
  if (mq == (mqd_t)-1) return -1;
  struct mq_attr attr;
  if (mq_getattr(mq, &attr) != 0 || attr.mq_msgsize <= 0) {
    mq_close(mq);
    return -1;
  }
  char* buf = (char*)malloc((size_t)attr.mq_msgsize);
  youth_reasm* r = youth_reasm_create();
  long frames = buf && r ? 0 : -1;
  int idle_ms = 0;
  while (frames >= 0 && !(stop && *stop)) {
    struct timespec to;
    clock_gettime(CLOCK_REALTIME, &to);
    to.tv_nsec += 20 * 1000 * 1000; /* poll period for `stop` */
    if (to.tv_nsec >= 1000000000L) {
      to.tv_sec += 1;
      to.tv_nsec -= 1000000000L;
    }
    const ssize_t n = mq_timedreceive(mq, buf, (size_t)attr.mq_msgsize, NULL, &to);
    if (n < 0) {
      if (errno == EINTR) continue;
      if (errno != ETIMEDOUT) {
        frames = -1;
        break;
      }
      idle_ms += 20;
      if (idle_timeout_ms > 0 && idle_ms >= idle_timeout_ms) break;
      continue;
    }
    idle_ms = 0;
    if (youth_reasm_feed(r, buf, (size_t)n) == 1) { /* malformed messages (-1) are dropped, like the viewer does */
      if (!sink((const int16_t*)r->depth, r->color, r->width, r->height, r->timestamp)) break;
      ++frames;
    }
  }
  youth_reasm_destroy(r);
  free(buf);
  mq_close(mq);
  return frames;
}
