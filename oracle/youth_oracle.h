/*
 * youth_oracle.h -- CPU oracle of the frame-to-frame depth tracking path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under slam-rgbd_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do, and only as the checker or the reported CPU baseline.
 *
 * PARITY UNPINNED for stages 1-5: the reference (SeunghwanByun/SLAM-RGBD) contains no
 * dense ICP tracker (its AlgorithmModule forwards to un-vendored, un-pinned ORB-SLAM3:
 * Youth.Source/AlgorithmModule/SLAM.cpp:54,78-83) and ships no tests or golden vectors,
 * so this file DEFINES the expected arithmetic instead of restating it.  The parts the
 * reference does pin are followed and cited where they are used:
 *   - depth unit: raw / 1000 -> metres        ViewerModule/viewerModule.c:343, SLAM.cpp:153-155
 *   - validity: raw depth > 0                  viewerModule.c:341
 *   - pinhole back-projection, no distortion   viewerModule.c:344-345
 *   - intrinsics 570.3 / 320 / 240 @ 640x480   AlgorithmModule/config/astra_orb_slam3_rgbd.yaml:9-12,19-20
 *   - record format FrameHeader + payloads     frameDefinitions.h:11-20, loggingModule.c:101-130,404-444
 *     (that one IS pinned, against the reference's own reader/writer: oracle/_ref)
 *
 * Arithmetic rules (the device path obeys the same ones, which is what makes the
 * comparison bit-exact): float unless stated, one rounding per written operation
 * (compile with -ffp-contract=off) EXCEPT where the source says fmaf(), which is a
 * single-rounding IEEE fused multiply-add on both sides (stage 3 only); IEEE division
 * and sqrt; no libm in anything that depends on the data (tables of exp() are built once
 * from the config; fmaf is exact by definition); reductions in the fixed order documented
 * at yo_icp_sums().
 */
#ifndef YOUTH_ORACLE_H
#define YOUTH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YO_MAX_LEVELS 4
#define YO_ICP_LANES 32
#define YO_SUM_SLOTS 32
/* slot layout of the sums (see icp_pixel in youth_oracle.c) */
#define YO_SUMS_B0 24
#define YO_SUMS_RR 30
#define YO_SUMS_COUNT 31
#define YO_BILATERAL_RADIUS 3
#define YO_RANGE_LUT_MAX 1024

typedef struct yo_config {
  int32_t width, height;
  float fx, fy, cx, cy;
  float depth_factor;
  int32_t levels;
  int32_t iters[YO_MAX_LEVELS];
  int32_t depth_min_mm, depth_max_mm;
  int32_t bilateral;
  float sigma_space_px, sigma_range_mm;
  float dist_thresh_m, cos_thresh;
  int32_t min_inliers;
  int32_t icp_ppt;
} yo_config;

/* per-level geometry derived from the config */
typedef struct yo_level {
  int32_t w, h;
  float fx, fy, cx, cy;
} yo_level;

/* one preprocessed frame: per level depth (raw units, 0 invalid), pyramid sample count,
 * vertex map and normal map as float4 (x,y,z,valid) */
typedef struct yo_frame {
  float* depth[YO_MAX_LEVELS];
  uint8_t* pyrcnt[YO_MAX_LEVELS];
  float* vmap[YO_MAX_LEVELS];
  float* nmap[YO_MAX_LEVELS];
} yo_frame;

void yo_default_config(yo_config* cfg);
int yo_level_geometry(const yo_config* cfg, int level, yo_level* out);

yo_frame* yo_frame_alloc(const yo_config* cfg);
void yo_frame_free(yo_frame* f);

/* stage 1a: validity + 7x7 bilateral on raw depth -> float depth in raw units */
void yo_bilateral(const yo_config* cfg, const uint16_t* raw, float* depth0);
/* stage 1b: one pyramid step (w,h = source size) */
void yo_pyrdown(const yo_config* cfg, int w, int h, const float* src, float* dst, uint8_t* cnt);
/* stage 2: vertex + normal maps of one level */
void yo_vertex_normal(const yo_config* cfg, int level, const float* depth, float* vmap, float* nmap);
/* the normal half of stage 2 on its own (also used on ray-cast model vertex maps) */
void yo_normals_from_vertices(int w, int h, const float* vmap, float* nmap);
/* stages 1+2 for all levels */
void yo_preprocess(const yo_config* cfg, const uint16_t* raw, yo_frame* out);

/* stage 3+4: association, residual, Jacobian, fixed-order reduction.  pose = prev<-cur,
 * row-major 3x4 float.  sums[YO_SUM_SLOTS] double.  corr nullable int32[w*h]. */
void yo_icp_sums(const yo_config* cfg, int level, const yo_frame* cur, const yo_frame* prev,
                 const float pose[12], double* sums, int32_t* corr);
/* stage 5: 6x6 solve + SE(3) update of pose_d (double 3x4) and its float copy.
 * returns 1 when the pose was updated, 0 when the iteration was skipped. */
int yo_solve_update(const yo_config* cfg, const double* sums, double pose_d[12], float pose_f[12]);
/* full schedule for one frame pair; rel_out = prev<-cur (double 3x4); returns status bits,
 * *inliers_out = inlier count of the last iteration at level 0 */
uint32_t yo_track_pair(const yo_config* cfg, const yo_frame* cur, const yo_frame* prev,
                       double rel_out[12], int32_t* inliers_out);
/* world_new = world_prev * rel (double 3x4) */
void yo_compose(const double world_prev[12], const double rel[12], double world_new[12]);

/* sequence tracker mirroring youth_cuda_track */
typedef struct yo_tracker yo_tracker;
yo_tracker* yo_tracker_create(const yo_config* cfg);
void yo_tracker_destroy(yo_tracker* t);
void yo_tracker_reset(yo_tracker* t);
/* returns status bits; pose_out = camera-to-world float 3x4 */
uint32_t yo_tracker_track(yo_tracker* t, const uint16_t* raw, float pose_out[12]);
int32_t yo_tracker_last_inliers(const yo_tracker* t);
const yo_frame* yo_tracker_frame(const yo_tracker* t, int which /*0 = newest, 1 = previous*/);

/* track n frames of one sequence, poses_out[n][12], status_out[n] nullable; returns
 * seconds spent inside the tracker (CLOCK_MONOTONIC), excludes nothing else */
double yo_track_sequence(const yo_config* cfg, const uint16_t* frames, int n, float* poses_out,
                         uint32_t* status_out);

#define YO_STATUS_FIRST 1u
#define YO_STATUS_LOST 2u

#define YO_REJ_CUR_INVALID (-1)
#define YO_REJ_BEHIND (-2)
#define YO_REJ_OUT_OF_IMAGE (-3)
#define YO_REJ_PREV_INVALID (-4)
#define YO_REJ_DISTANCE (-5)
#define YO_REJ_ANGLE (-6)

#ifdef __cplusplus
}
#endif
#endif
