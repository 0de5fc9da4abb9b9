/*
 * ref_logging_shim.c -- TEST INFRASTRUCTURE.  Compiles the reference's LoggingModule from
 * the source file where it lies (the path arrives as -DREF_LOGGING_C="...": nothing is
 * copied into this repository) and exposes its file-static record writer / reader / chunk
 * sender so that tests can pin our restatements against the reference's own code:
 *
 *   ref_save_frames   -> saveFrameToFile      loggingModule.c:101-130
 *   ref_read_frames   -> readFrameFromFile    loggingModule.c:404-444
 *   ref_chunk_stream  -> sendMetadata + sendDataInChunks   loggingModule.c:447-502
 *
 * Built by oracle/Makefile into oracle/_ref/libref_logging.so only where /root/reference
 * is mounted; the prebuilt .so travels to the GPU box.
 */
#include REF_LOGGING_C

#include <stdint.h>

/* write n frames (depth u16 [n][h][w], colour u8 [n][h][w][3]) with the reference writer */
int ref_save_frames(const char* path, int n, int width, int height, const uint16_t* depth, const uint8_t* color,
                    const uint32_t* timestamps) {
  recordFile = fopen(path, "wb");
  if (!recordFile) return 0;
  currentWidth = width;
  currentHeight = height;
  const size_t dpx = (size_t)width * height;
  for (int i = 0; i < n; ++i) {
    depthBuffer = (char*)(depth + dpx * i);
    colorBuffer = (char*)(color + dpx * 3 * i);
    saveFrameToFile(i, timestamps ? timestamps[i] : (uint32_t)(33 * i));
  }
  fclose(recordFile);
  recordFile = NULL;
  depthBuffer = NULL;
  colorBuffer = NULL;
  return 1;
}

/* read up to max_frames with the reference reader; returns the number read.
 * headers_out: max_frames * sizeof(FrameHeader) bytes. */
int ref_read_frames(const char* path, int max_frames, int max_payload, void* headers_out, char* depth_out,
                    char* color_out, size_t depth_stride, size_t color_stride) {
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  int n = 0;
  while (n < max_frames) {
    FrameHeader h;
    if (!readFrameFromFile(f, &h, depth_out + depth_stride * n, color_out + color_stride * n, max_payload)) break;
    memcpy((char*)headers_out + sizeof(FrameHeader) * n, &h, sizeof(h));
    ++n;
  }
  fclose(f);
  return n;
}

struct chunk_job {
  mqd_t mq;
  int frame_id, width, height;
  uint32_t ts;
  const char* depth;
  const char* color;
};

static void* chunk_sender(void* arg) {
  struct chunk_job* j = (struct chunk_job*)arg;
  const int dsz = j->width * j->height * 2, csz = j->width * j->height * 3;
  sendMetadata(j->mq, j->frame_id, j->ts, j->width, j->height);
  sendDataInChunks(j->mq, MSG_TYPE_DEPTH_DATA, j->frame_id, j->ts, j->width, j->height, j->depth, dsz);
  sendDataInChunks(j->mq, MSG_TYPE_COLOR_DATA, j->frame_id, j->ts, j->width, j->height, j->color, csz);
  return NULL;
}

/* push one frame through a real POSIX mq with the reference sender; the receiver side
 * records every message verbatim: msgs_out[k * MAX_MSG_SIZE], lens_out[k].  Returns the
 * number of messages, -1 when mqueues are unavailable. */
int ref_chunk_stream(const char* mq_name, int frame_id, uint32_t ts, int width, int height, const char* depth,
                     const char* color, char* msgs_out, int* lens_out, int max_msgs) {
  struct mq_attr attr;
  memset(&attr, 0, sizeof(attr));
  attr.mq_maxmsg = 10; /* loggingModule.c:137-141 */
  attr.mq_msgsize = MAX_MSG_SIZE;
  mq_unlink(mq_name);
  mqd_t mq = mq_open(mq_name, O_CREAT | O_RDWR, 0644, &attr);
  if (mq == (mqd_t)-1) return -1;
  struct chunk_job job = {mq, frame_id, width, height, ts, depth, color};
  pthread_t th;
  pthread_create(&th, NULL, chunk_sender, &job);
  const int maxData = MAX_MSG_SIZE - (int)sizeof(MessageHeader);
  const int expect = 1 + (width * height * 2 + maxData - 1) / maxData + (width * height * 3 + maxData - 1) / maxData;
  int n = 0;
  while (n < expect && n < max_msgs) {
    struct timespec to;
    clock_gettime(CLOCK_REALTIME, &to);
    to.tv_sec += 5;
    ssize_t r = mq_timedreceive(mq, msgs_out + (size_t)n * MAX_MSG_SIZE, MAX_MSG_SIZE, NULL, &to);
    if (r < 0) break;
    lens_out[n++] = (int)r;
  }
  pthread_join(th, NULL);
  mq_close(mq);
  mq_unlink(mq_name);
  return n;
}
