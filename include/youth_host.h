/*
 * youth_host.h -- host-side C helpers around the tracker: the .bin record format, the
 * mq chunk protocol, the synthetic Astra-shaped sequence generator, YAML camera config
 * and TUM trajectory egress.  Plain C, no CUDA; linked into libAlgorithmModule.so.
 *
 * Reference interfaces mirrored here:
 *   youth_bin_write_frame  <- saveFrameToFile     Youth.Source/LoggingModule/loggingModule.c:101-130
 *   youth_bin_write_eof    <- EOF marker record   loggingModule.c:223-226, frameDefinitions.h:8
 *   youth_bin_read_frame   <- readFrameFromFile   loggingModule.c:404-444
 *   youth_chunk_*          <- sendDataInChunks    loggingModule.c:447-485 / sensorModule.c:149-210
 *   youth_reasm_*          <- loggerThread reassembly + completion test   loggingModule.c:292-357
 *   youth_mq_consume       <- the viewer's receive loop   ViewerModule/viewerModule.c:160-250 (dataReceiveThread)
 *   youth_config_from_yaml <- cv::FileStorage keys read by ORB-SLAM3 from
 *                             AlgorithmModule/config/astra_orb_slam3_rgbd.yaml:9-20,35
 *   youth_tum_write        <- SaveTrajectoryTUM call site   AlgorithmModule/SLAM.cpp:187-188
 */
#ifndef YOUTH_HOST_H
#define YOUTH_HOST_H

#include <stdint.h>
#include <stdio.h>

#include "frameDefinitions.h"
#include "youth_codec.h"
#include "youth_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- .bin records */
/* color may be NULL: a constant-128 RGB plane is written so the record stays format-valid */
int youth_bin_write_frame(FILE* f, uint32_t frame_id, uint32_t timestamp_ms, int width, int height,
                          const uint16_t* depth, const uint8_t* color);
/* the same record with the depth payload packed by the YD16 codec (include/youth_codec.h):
 * frameType = FRAME_TYPE_DEPTH_PACKED, depthDataSize = packed bytes; width/height say what it unpacks to */
int youth_bin_write_packed_frame(FILE* f, uint32_t frame_id, uint32_t timestamp_ms, int width, int height,
                                 const uint8_t* yd16, uint32_t yd16_bytes, const uint8_t* color);
int youth_bin_write_eof(FILE* f);
/* 1 = frame read, 0 = end of file / EOF marker / error / payload larger than the caps.
 * color may be NULL (payload skipped). */
int youth_bin_read_frame(FILE* f, FrameHeader* hdr, void* depth, size_t depth_cap, void* color,
                         size_t color_cap);
/* the same in two steps, for a reader that picks the destination after it has seen the header (e.g. a slot of the
 * tracker's page-locked frame ring, include/youth_slam_ext.h): header (0 at end of file / EOF marker), then the
 * payloads of that header */
int youth_bin_read_header(FILE* f, FrameHeader* hdr);
int youth_bin_read_payload(FILE* f, const FrameHeader* hdr, void* depth, size_t depth_cap, void* color, size_t color_cap);

/* ---------------------------------------------------------------- mq chunk protocol */
/* number of MAX_MSG_SIZE messages needed for `bytes` of payload */
int youth_chunk_count(size_t bytes);
/* build message `chunk` of a payload into msg[MAX_MSG_SIZE]; returns message length */
size_t youth_chunk_build(void* msg, int msg_type, int frame_id, uint32_t timestamp_ms, int width,
                         int height, const void* data, size_t data_bytes, int chunk);

/* pose egress (MSG_TYPE_POSE): build / parse one message; build returns the message length,
 * parse returns 1 when msg is a well-formed pose message */
size_t youth_pose_msg_build(void* msg, int frame_id, uint32_t timestamp_ms, const float pose[12], uint32_t status,
                            uint32_t inliers);
int youth_pose_msg_parse(const void* msg, size_t len, int* frame_id, uint32_t* timestamp_ms, YouthPoseMsg* out);

typedef struct youth_reasm youth_reasm;
youth_reasm* youth_reasm_create(void);
void youth_reasm_destroy(youth_reasm* r);
/* feed one mq message; returns 1 when this message completed a frame (depth and colour
 * both whole), 0 otherwise, -1 on a malformed message */
int youth_reasm_feed(youth_reasm* r, const void* msg, size_t len);
const uint16_t* youth_reasm_depth(const youth_reasm* r);
const uint8_t* youth_reasm_color(const youth_reasm* r);
void youth_reasm_info(const youth_reasm* r, int* width, int* height, int* frame_id, uint32_t* timestamp_ms);

/* Viewer-side queue consumer (the queue the logger's pass-through and playbackThread write,
 * loggingModule.c:284-288, :584-590; reader in the reference: viewerModule.c:160-250).  Opens `queue`
 * (e.g. MQ_LOGGER_TO_VIEWER) read-only, reassembles frames and calls `sink` -- processSlamFrame or anything
 * of its signature -- once per whole frame, on the calling thread.  Returns when *stop becomes non-zero
 * (stop may be NULL), when no message arrived for idle_timeout_ms (<= 0: never), or when sink returns 0.
 * Return value: frames delivered, -1 if the queue cannot be opened or read. */
typedef int (*youth_frame_sink)(const int16_t* depth, const uint8_t* color, int width, int height, uint32_t timestamp_ms);
long youth_mq_consume(const char* queue, youth_frame_sink sink, volatile int* stop, int idle_timeout_ms);

/* ---------------------------------------------------------------- synthetic sequences */
typedef struct youth_synth_config {
  int32_t width, height;
  double fx, fy, cx, cy;
  uint32_t seed;      /* 20261018 + sequence index */
  int32_t period;     /* frames per trajectory period (300) */
  double phase;       /* 2*pi*sequence/64 */
  double dropout;     /* fraction of pixels invalidated by the hash (0.02) */
  int32_t noise;      /* 1: add uniform {-2..2} mm */
  int32_t dmin_mm, dmax_mm; /* sensor working range, outside -> 0 */
} youth_synth_config;

void youth_synth_default(youth_synth_config* c, int width, int height, int sequence);
/* camera-to-world pose of frame i, row-major 3x4 double */
void youth_synth_pose(const youth_synth_config* c, int frame, double T_wc[12]);
/* ground-truth pose of frame i expressed in the frame-0 camera (what the tracker estimates) */
void youth_synth_gt(const youth_synth_config* c, int frame, double T_0c[12]);
/* ray-cast one uint16 depth frame (mm).  Rows are split over pthreads (YOUTH_SYNTH_THREADS overrides the count). */
void youth_synth_frame(const youth_synth_config* c, int frame, uint16_t* depth_out);
/* frames [first, first+n) tightly packed */
void youth_synth_sequence(const youth_synth_config* c, int first, int n, uint16_t* depth_out);

/* ---------------------------------------------------------------- config + egress */
/* start from youth_cuda_default_config, then override from YAML keys Camera.fx/fy/cx/cy/
 * width/height and DepthMapFactor.  path NULL/"" -> defaults.  0 if the file cannot be read. */
int youth_config_from_yaml(const char* path, youth_cuda_config* cfg);
/* rotation (row-major 3x4 [R|t]) -> unit quaternion x,y,z,w */
void youth_pose_to_quat(const float pose[12], double q_xyzw[4]);
/* TUM trajectory: "ts tx ty tz qx qy qz qw" per line, ts in seconds */
int youth_tum_write(const char* path, const float* poses, const uint32_t* timestamps_ms, int n);

#ifdef __cplusplus
}
#endif
#endif
