/*
 * youth_slam_ext.h -- additions to the AlgorithmModule facade (include/SLAM.h) for headless drivers and for
 * producers that already assemble frames themselves.  None of these exist in the reference
 * (Youth.Source/AlgorithmModule/SLAM.h:11-38 has seven functions); the seven reference entry points keep their
 * behaviour whether or not these are used.  Same 1 = success / 0 = failure convention.
 */
#ifndef YOUTH_SLAM_EXT_H
#define YOUTH_SLAM_EXT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Call before initSlamModule().  lossless: 1 = the producer blocks when the ring is full instead of the
 * reference's drop-oldest policy (SLAM.cpp:163-167), for replay / benchmarks; batch: frames per launch group
 * (1..512).  -1 keeps the environment (YOUTH_SLAM_LOSSLESS, YOUTH_SLAM_BATCH) or default choice. */
void youthSlamSetOptions(int lossless, int batch);

/* Block until every accepted frame has been tracked. */
void youthSlamDrain(void);

/* Zero-copy form of processSlamFrame(): reserve the next slot of the page-locked host frame ring, fill it with
 * width*height uint16 depth values (mm), publish it.  A producer that assembles frames anyway -- the logger's
 * chunk reassembly (reference loggingModule.c:303-327), a file reader -- writes them where the GPU's copy
 * engine reads them, instead of into a buffer of its own that processSlamFrame() copies once more
 * (SLAM.cpp:133-134: 614 KB per VGA frame on the caller's thread).  One slot is outstanding at a time.
 * Acquire returns NULL when no frame can be taken (not running / stopping / wrong size / tracker failed; in the
 * lossy mode with a physically full ring the frame counts as accepted and dropped).  Commit returns 1 when the
 * frame was queued. */
uint16_t* youthSlamAcquireSlot(int width, int height);
int youthSlamCommitSlot(uint32_t timestamp_ms);
void youthSlamAbortSlot(void); /* give the acquired slot back unpublished (the frame could not be completed) */

/* n frames back to back that already live in page-locked host memory (youth_cuda_host_alloc): no CPU copy, the
 * GPU's copy engine reads them in place, launch groups of the configured size, two in flight.  Ordered after
 * everything queued before; synchronous (returns when the n frames are tracked). */
int youthSlamProcessPinnedFrames(const uint16_t* frames, int n, int width, int height, const uint32_t* timestamps);

/* Replay of FRAME_TYPE_DEPTH_PACKED records (include/youth_codec.h): n YD16 streams back to back, offsets[n+1]. */
int youthSlamProcessPackedFrames(const uint8_t* streams, const uint64_t* offsets, int n, int width, int height,
                                 const uint32_t* timestamps);

/* Copy out up to max_frames poses (12 floats each) / timestamps / YOUTH_STATUS_* words; returns the count. */
int youthSlamGetTrajectory(float* poses_out, uint32_t* timestamps_out, uint32_t* status_out, int max_frames);

/* Frames accepted by processSlamFrame / commit, dropped by the lossy back-pressure, tracked. */
void youthSlamStats(long* accepted, long* dropped, long* tracked);

#ifdef __cplusplus
}
#endif
#endif /* YOUTH_SLAM_EXT_H */
