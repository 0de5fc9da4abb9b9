/*
 * ref_logging_hooked_shim.c -- TEST INFRASTRUCTURE.  The reference's LoggingModule with the ONE-LINE
 * hook of INTEGRATION.md section 3 applied: oracle/Makefile pipes the reference source through sed into
 * a temporary file outside this repository (the only edit: the frame-complete test at
 * loggingModule.c:354 first hands the frame to processSlamFrame()), and this file #includes that
 * temporary (-DREF_LOGGING_HOOKED_C="..."), so the reference's own loggerThread -- its mq receive loop,
 * its pass-through, its chunk reassembly -- runs unmodified around the hook.  Nothing of the reference
 * is copied into the repository; the output is oracle/_ref/libref_logging_hooked.so.
 *
 * The test supplies the facade entry points as function pointers (no link dependency on the product),
 * a fake sensor that uses the reference's own sendMetadata / sendDataInChunks on the sensor queue, and a
 * fake viewer that drains the viewer queue.
 */
#include <stdint.h>

static int (*g_hook_running)(void);
static int (*g_hook_process)(const int16_t*, const uint8_t*, int, int, uint32_t);
static volatile long g_hook_calls;
#define isSlamModuleRunning() (g_hook_running ? g_hook_running() : 0)
#define processSlamFrame(d, c, w, h, t) (g_hook_calls++, g_hook_process((d), (c), (w), (h), (t)))

#include REF_LOGGING_HOOKED_C

void ref_hook_install(void* running, void* process) {
  g_hook_running = (int (*)(void))running;
  g_hook_process = (int (*)(const int16_t*, const uint8_t*, int, int, uint32_t))process;
  g_hook_calls = 0;
}
long ref_hook_calls(void) { return g_hook_calls; }

/* ---- fake viewer: drains MQ_LOGGER_TO_VIEWER so the logger's blocking pass-through never stalls ---- */
static pthread_t g_viewer;
static volatile int g_viewer_run;
static volatile long g_viewer_msgs;
static void* viewer_drain(void* arg) {
  (void)arg;
  mqd_t mq = mq_open(MQ_LOGGER_TO_VIEWER, O_RDONLY);
  if (mq == (mqd_t)-1) return NULL;
  char* buf = (char*)malloc(MAX_MSG_SIZE);
  while (g_viewer_run) {
    struct timespec to;
    clock_gettime(CLOCK_REALTIME, &to);
    to.tv_nsec += 50 * 1000 * 1000;
    if (to.tv_nsec >= 1000000000L) {
      to.tv_sec += 1;
      to.tv_nsec -= 1000000000L;
    }
    if (mq_timedreceive(mq, buf, MAX_MSG_SIZE, NULL, &to) > 0) g_viewer_msgs++;
  }
  free(buf);
  mq_close(mq);
  return NULL;
}

/* start the reference logging module (its queues, loggerThread, playbackThread) and the fake viewer */
int ref_pipeline_start(void) {
  mq_unlink(MQ_SENSOR_TO_LOGGER);
  mq_unlink(MQ_LOGGER_TO_VIEWER);
  mq_unlink(MQ_CONTROL_QUEUE);
  isPassThroughEnabled = 1; /* the module's statics as a fresh process has them (a finished playback leaves 0 behind) */
  isPlaybackActive = 0;
  initLoggingModule(); /* loggingModule.c:616 */
  g_viewer_run = 1;
  g_viewer_msgs = 0;
  if (pthread_create(&g_viewer, NULL, viewer_drain, NULL) != 0) return 0;
  usleep(100000); /* the reference threads open their queues asynchronously */
  return mqFromSensor != (mqd_t)-1;
}

/* fake sensor: one frame as SensorModule sends it (sensorModule.c:128-210), with the reference's own
 * chunk sender on the sensor -> logger queue; blocks while the 10-slot queue is full */
int ref_fake_sensor_send(int frame_id, uint32_t ts, int width, int height, const char* depth, const char* color) {
  mqd_t mq = mq_open(MQ_SENSOR_TO_LOGGER, O_WRONLY);
  if (mq == (mqd_t)-1) return 0;
  sendMetadata(mq, frame_id, ts, width, height);
  sendDataInChunks(mq, MSG_TYPE_DEPTH_DATA, frame_id, ts, width, height, depth, width * height * 2);
  sendDataInChunks(mq, MSG_TYPE_COLOR_DATA, frame_id, ts, width, height, color, width * height * 3);
  mq_close(mq);
  return 1;
}

long ref_viewer_messages(void) { return g_viewer_msgs; }

/* ---- playback direction: the reference's playbackThread (loggingModule.c:505-611) writes a recording into
 * MQ_LOGGER_TO_VIEWER; the test puts the product's queue consumer in the viewer's seat ---- */
/* retire the fake viewer so that somebody else can read the viewer queue */
void ref_viewer_stop(void) {
  if (!g_viewer_run) return;
  g_viewer_run = 0;
  pthread_join(g_viewer, NULL);
}
/* the reference's own startPlayback (loggingModule.c:710-720) + sensor messages: the command sits in the
 * control queue until loggerThread comes round to polling it, and loggerThread blocks on the sensor queue
 * (SURVEY appendix A), so a metadata message wakes it */
int ref_start_playback(const char* filename) {
  if (!startPlayback(filename)) return 0;
  mqd_t mq = mq_open(MQ_SENSOR_TO_LOGGER, O_WRONLY);
  if (mq == (mqd_t)-1) return 0;
  sendMetadata(mq, -1, 0, 8, 8);
  sendMetadata(mq, -1, 0, 8, 8); /* the first one may be consumed before the command is in the control queue */
  mq_close(mq);
  return 1;
}
int ref_is_playing_back(void) { return isPlayingBack(); }

/* orderly stop: the reference's loggerThread blocks in mq_receive on the sensor queue,
 * so the flag is cleared first and one metadata message wakes it up */
void ref_pipeline_stop(void) {
  loggingIsRunning = 0;
  mqd_t mq = mq_open(MQ_SENSOR_TO_LOGGER, O_WRONLY | O_NONBLOCK);
  if (mq != (mqd_t)-1) {
    sendMetadata(mq, -1, 0, currentWidth > 0 ? currentWidth : 8, currentHeight > 0 ? currentHeight : 8);
    mq_close(mq);
  }
  stopLoggingModule(); /* loggingModule.c:668 */
  ref_viewer_stop();
  mq_unlink(MQ_SENSOR_TO_LOGGER);
  mq_unlink(MQ_LOGGER_TO_VIEWER);
  mq_unlink(MQ_CONTROL_QUEUE);
}
