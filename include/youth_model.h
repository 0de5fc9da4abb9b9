/*
 * youth_model.h -- frame-to-MODEL tracking on top of libyouth_cuda.so.
 *
 * The reference's engine call, `slam_system->TrackRGBD(...)` (Youth.Source/AlgorithmModule/
 * SLAM.cpp:54), tracks against a map, not against the previous frame; the dense equivalent is a
 * truncated signed distance volume fused from every tracked frame and ray-cast from the last pose
 * into the vertex / normal maps that stage 3 aligns the next frame to.  Stages 1-5 (and their
 * kernels) are unchanged: the ray-cast maps take the place of the previous frame's maps.
 *
 *   youth_cuda_enable_model   <- the `new ORB_SLAM3::System(...)` map construction   SLAM.cpp:78-83
 *   youth_cuda_track[_batch]  <- TrackRGBD against the map                          SLAM.cpp:54
 *   youth_cuda_reset          <- `slam_system->Reset()` (also clears the volume)     SLAM.cpp:226
 *
 * Per tracked frame, per sequence, with no host synchronisation (one captured CUDA graph):
 *   k_ingest, k_normals              stages 1-2 of the new frame
 *   k_icp x (iterations)             stages 3-5 against the model maps
 *   k_compose                        pose chain
 *   k_tsdf_integrate                 fuse the frame at its new pose (skipped when flagged LOST)
 *   k_tsdf_raycast, k_normals        model vertex maps of every pyramid level from the new pose (each ray
 *                                    starts just in front of the depth the fused frame measured at its
 *                                    pixel) and their cross-product normals (stage 2b, unchanged)
 * A sequence is a chain here (frame t needs the pose of t-1), so throughput comes from tracking
 * several independent sequences per launch (cfg.n_streams), each with its own volume.
 *
 * Arithmetic contract: oracle/youth_tsdf_oracle.c defines it; the device is bit-identical.
 * Return convention as in youth_cuda.h (1 / 0, youth_cuda_last_error()).
 */
#ifndef YOUTH_MODEL_H
#define YOUTH_MODEL_H

#include <stddef.h>
#include <stdint.h>

#include "youth_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct youth_tsdf_config {
  int32_t dim[3];     /* voxels along x, y, z (x fastest in memory); each 8..1024, dim[0] % 8 == 0 */
  float voxel_m;      /* edge length of a voxel in metres                                          */
  float origin[3];    /* world (= first camera frame) position of the corner of voxel (0,0,0)      */
  float trunc_m;      /* truncation distance mu; the ray-cast step is 0.8 mu                       */
  int32_t max_weight; /* cap of the per-voxel observation count, 1..32767                          */
  float near_m, far_m;/* depth range searched by the ray cast                                      */
} youth_tsdf_config;

/* 256 x 128 x 256 voxels of 25 mm around the first camera (x -3.2..3.2, y -1.6..1.6, z -1.2..5.2 m:
 * the synthetic room of SURVEY.md section 8(d) fits), mu = 0.10 m, weight cap 64, range 0.4..8 m. */
int youth_tsdf_default_config(youth_tsdf_config* cfg);

/* Switch a freshly initialised (or reset) handle to frame-to-model tracking: allocates one
 * volume (4 B per voxel) and one set of model maps per sequence.  Cannot be undone. */
int youth_cuda_enable_model(youth_cuda_handle* h, const youth_tsdf_config* cfg);

/* 1 when the handle tracks against a model */
int youth_cuda_model_enabled(const youth_cuda_handle* h);

/* Size of the map of sequence `stream`: observed voxels whose tsdf changes sign towards an observed +x, +y
 * or +z neighbour -- the voxels the fused surface passes through; the dense counterpart of
 * `GetAllMapPoints().size()` (reference SLAM.cpp:212-217).  Blocking; -1 on error. */
long long youth_cuda_model_surface_voxels(youth_cuda_handle* h, int stream);

/* ---- parity hooks (blocking) ---- */
/* volume of sequence `stream`: int16 [dim2][dim1][dim0][2] = (tsdf * 32767, weight) */
int youth_cuda_debug_read_volume(youth_cuda_handle* h, int stream, int16_t* dst, size_t dst_bytes);
/* model maps of `level`: what = YOUTH_DBG_VERTEX / YOUTH_DBG_NORMAL, float [h][w][4] = x, y, z, valid */
int youth_cuda_debug_read_model(youth_cuda_handle* h, int what, int stream, int level, float* dst, size_t dst_bytes);
/* run ONE kernel outside the tracking schedule: fuse resident frame `frame` of `stream` at `pose`
 * (camera-to-world 3x4) / ray-cast every level from `pose` into the model maps.
 * hint_frame >= 0: search every ray only from 1.25 mu in front of to 2 mu behind the depth that resident
 * frame measured at the pixel (what the tracker does with the frame it has just fused); -1: whole range */
int youth_cuda_debug_integrate(youth_cuda_handle* h, int stream, int frame, const float pose[12]);
int youth_cuda_debug_raycast(youth_cuda_handle* h, int stream, const float pose[12], int hint_frame);

#ifdef __cplusplus
}
#endif
#endif
