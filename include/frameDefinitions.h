/*
 * frameDefinitions.h -- shared frame ABI between Sensor / Logging / Algorithm / Viewer.
 *
 * Layout-compatible restatement of the reference's shared header
 * (reference: Youth.Source/frameDefinitions.h:7-64).  The struct and macro NAMES
 * and the binary layout are the contract (records are raw fwrite()s of these
 * structs: reference LoggingModule/loggingModule.c:118; mq chunks are raw
 * memcpy()s: reference SensorModule/sensorModule.c:149-176), so they are kept;
 * everything else (comments, compile-time layout checks, helper macros) is new.
 *
 * Compile-time checks below pin: sizeof(FrameHeader)==28, sizeof(MessageHeader)==292,
 * chunk payload == 7900 B.
 */
#ifndef FRAME_DEFINITIONS_H
#define FRAME_DEFINITIONS_H

#include <stddef.h>
#include <stdint.h>

/* ---- on-disk record kinds (FrameHeader.frameType) ---- */
#define FRAME_TYPE_DEPTH_COLOR 1
#define FRAME_TYPE_END_OF_FILE 0xFF

/* One record of a .bin capture: this header, then depthDataSize bytes of 16-bit
 * depth (mm, row-major, tightly packed), then colorDataSize bytes of RGB8. */
typedef struct {
  uint32_t frameId;       /* running frame counter                    (offset  0) */
  uint32_t timestamp;     /* milliseconds, wraps at 2^32              (offset  4) */
  uint16_t frameType;     /* FRAME_TYPE_*                             (offset  8) */
  uint16_t width;         /* pixels                                   (offset 10) */
  uint16_t height;        /* pixels                                   (offset 12) */
                          /* 2 bytes of natural padding               (offset 14) */
  uint32_t depthDataSize; /* bytes of depth payload                   (offset 16) */
  uint32_t colorDataSize; /* bytes of colour payload                  (offset 20) */
  uint32_t reserved;      /* zero in reference recordings             (offset 24) */
} FrameHeader;

/* Legacy metadata struct kept for source compatibility (unused on the hot path). */
typedef struct {
  int width;
  int height;
  int frameId;
  uint32_t timestamp;
  int dataSize;
} SensorDataMsg;

/* ---- message-queue chunk kinds (MessageHeader.msgType) ---- */
#define MSG_TYPE_METADATA 1
#define MSG_TYPE_DEPTH_DATA 2
#define MSG_TYPE_COLOR_DATA 3
#define MSG_TYPE_CONTROL 4

/* ---- control commands (MessageHeader.ctrlCommand) ---- */
#define CTRL_CMD_START_RECORD 1
#define CTRL_CMD_STOP_RECORD 2
#define CTRL_CMD_START_PLAYBACK 3
#define CTRL_CMD_STOP_PLAYBACK 4

/* Header that prefixes every POSIX-mq message; payload follows in the same message. */
typedef struct {
  int msgType;
  int width;
  int height;
  int chunkIndex;     /* reassembly offset = chunkIndex * YOUTH_CHUNK_PAYLOAD */
  int totalChunks;
  int dataSize;       /* payload bytes carried by THIS message */
  int frameId;
  uint32_t timestamp;
  int ctrlCommand;    /* only meaningful for MSG_TYPE_CONTROL */
  char filename[256]; /* only meaningful for MSG_TYPE_CONTROL */
} MessageHeader;

#define MQ_SENSOR_TO_LOGGER "/sensor_logger_queue"
#define MQ_LOGGER_TO_VIEWER "/logger_viewer_queue"
#define MQ_CONTROL_QUEUE "/control_queue"

#define MAX_MSG_SIZE 8192

/* ---- additions (not in the reference): pose egress message, derived constants, layout pins ---- */
/* Next free message type after MSG_TYPE_CONTROL: one message =
 * MessageHeader (frameId / timestamp of the tracked frame, dataSize = sizeof(YouthPoseMsg)) +
 * YouthPoseMsg.  Old viewers ignore unknown types (viewerModule.c:197-248 switches on 1..3). */
#define MSG_TYPE_POSE 5
typedef struct {
  float pose[12];   /* camera-to-world, row-major 3x4 [R|t], world = first camera frame */
  uint32_t status;  /* the frame's own YOUTH_STATUS_* bits (FIRST = 1: origin of a sequence, LOST = 2: pose not updated) */
  uint32_t inliers; /* correspondences of the last ICP iteration at the finest level, of the LAST frame of the launch
                       group this frame was tracked in (the tracker keeps one count per sequence) */
} YouthPoseMsg;

#define YOUTH_CHUNK_PAYLOAD ((int)(MAX_MSG_SIZE - sizeof(MessageHeader))) /* 7900 */
#define YOUTH_CHUNKS_FOR(bytes) (((bytes) + YOUTH_CHUNK_PAYLOAD - 1) / YOUTH_CHUNK_PAYLOAD)

#if defined(__cplusplus)
#define YOUTH_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define YOUTH_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
YOUTH_STATIC_ASSERT(sizeof(FrameHeader) == 28, "FrameHeader must be 28 bytes");
YOUTH_STATIC_ASSERT(offsetof(FrameHeader, frameType) == 8, "frameType offset");
YOUTH_STATIC_ASSERT(offsetof(FrameHeader, width) == 10, "width offset");
YOUTH_STATIC_ASSERT(offsetof(FrameHeader, height) == 12, "height offset");
YOUTH_STATIC_ASSERT(offsetof(FrameHeader, depthDataSize) == 16, "depthDataSize offset");
YOUTH_STATIC_ASSERT(offsetof(FrameHeader, colorDataSize) == 20, "colorDataSize offset");
YOUTH_STATIC_ASSERT(offsetof(FrameHeader, reserved) == 24, "reserved offset");
YOUTH_STATIC_ASSERT(sizeof(MessageHeader) == 292, "MessageHeader must be 292 bytes");
YOUTH_STATIC_ASSERT(MAX_MSG_SIZE - sizeof(MessageHeader) == 7900, "chunk payload is 7900 B");
YOUTH_STATIC_ASSERT(sizeof(YouthPoseMsg) == 56, "YouthPoseMsg must be 56 bytes");

#endif /* FRAME_DEFINITIONS_H */
