/*
 * SLAM.h -- the AlgorithmModule C facade (drop-in for reference
 * Youth.Source/AlgorithmModule/SLAM.h:11-38).  Same seven exported C symbols, same
 * argument meaning, same 1 = success / 0 = failure convention (SLAM.h:21,26,30).
 *
 * Differences from the reference header, all source-compatible:
 *   - includes <stdint.h> (the reference forgets it and does not compile from C;
 *     SURVEY.md section 0, fact F3);
 *   - prototypes use (void) so they are real C prototypes;
 *   - the engine behind the facade is the dense frame-to-frame ICP tracker in
 *     libyouth_cuda.so (include/youth_cuda.h) instead of ORB-SLAM3.
 */
#ifndef SLAM_H
#define SLAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Start the tracker. config_file: YAML with Camera.fx/fy/cx/cy/width/height and
 * DepthMapFactor (reference config/astra_orb_slam3_rgbd.yaml:9-20,35); NULL or ""
 * selects the Astra defaults.  vocabulary_file is accepted and ignored (an ORB
 * vocabulary has no meaning for dense ICP).  Outcome via isSlamModuleRunning().
 * Replaces reference SLAM.cpp:67-95. */
void initSlamModule(const char* config_file, const char* vocabulary_file);

/* Drain queued frames, stop the worker, release the device.  SLAM.cpp:97-124. */
void stopSlamModule(void);

/* Hand one frame to the tracker.  depth_data: width*height 16-bit millimetres,
 * row-major, 0 = no reading; color_data: RGB8 (ignored by the tracker, may be
 * NULL); timestamp in ms.  The depth buffer is copied before returning, so the
 * caller may reuse it immediately (reference SLAM.cpp:133-134).  Returns 1 when
 * queued, 0 when the module is not running or the size does not match the
 * configured camera.  Replaces reference SLAM.cpp:126-175. */
int processSlamFrame(const int16_t* depth_data, const uint8_t* color_data, int width, int height,
                     uint32_t timestamp);

/* Write <map_file>_trajectory.txt and <map_file>_keyframes.txt in TUM format
 * ("ts tx ty tz qx qy qz qw", reference SLAM.cpp:187-188).  1 / 0. */
int saveSlamMap(const char* map_file);

/* 1 while the module is initialised and its worker is alive.  SLAM.cpp:200-202. */
int isSlamModuleRunning(void);

/* Reference returns the ORB map-point count (SLAM.cpp:204-218); the dense
 * tracker's closest analogue is the inlier correspondence count of the last
 * tracked frame at the finest level. */
int getSlamMapPoints(void);

/* Forget the trajectory and the previous frame; next frame becomes the origin.
 * SLAM.cpp:220-228. */
void resetSlam(void);

#ifdef __cplusplus
}
#endif

#endif /* SLAM_H */
