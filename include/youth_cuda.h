/*
 * youth_cuda.h -- C ABI of libyouth_cuda.so, the B200 (sm_100a) dense frame-to-frame
 * depth tracker that sits behind the AlgorithmModule facade (include/SLAM.h).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * What each entry point replaces in the reference:
 *
 *   youth_cuda_init          <- `new ORB_SLAM3::System(...)`        Youth.Source/AlgorithmModule/SLAM.cpp:78-83
 *   youth_cuda_track[_batch] <- `slam_system->TrackRGBD(...)`       SLAM.cpp:54 (+ the int16->float metres
 *                                                                   conversion of SLAM.cpp:133-134,153-155,
 *                                                                   fused into the ingest kernel)
 *   youth_cuda_get_trajectory<- `SaveTrajectoryTUM` pose egress     SLAM.cpp:187-188
 *   youth_cuda_last_inliers  <- `GetAllMapPoints().size()`          SLAM.cpp:212-217
 *   youth_cuda_reset         <- `slam_system->Reset()`              SLAM.cpp:226
 *   youth_cuda_destroy       <- `slam_system->Shutdown()`           SLAM.cpp:110
 *   back-projection formula  <- display_3d_color()                  Youth.Source/ViewerModule/viewerModule.c:341-345
 *
 * Return convention follows the facade (SLAM.h:21,26,30): 1 = success, 0 = failure;
 * youth_cuda_last_error() returns a thread-local description of the last failure.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Threading: a handle is not thread-safe -- drive each handle from one thread (the facade's
 * worker does) or lock around it; different handles (e.g. one per GPU) are independent.
 * Groups of at most 8 frames per sequence fed from host memory run as one captured CUDA graph
 * (the launch-bound live path); YOUTH_CUDA_GRAPHS=0 in the environment disables that.
 *
 * Numerical contract (frozen by oracle/youth_oracle.c): fused
 * multiply-add only where the specification names it (stage 3), never by compiler
 * contraction; IEEE division/sqrt; fixed-order reductions -- the device results are
 * bit-identical to the CPU oracle.
 */
#ifndef YOUTH_CUDA_H
#define YOUTH_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOUTH_CUDA_ABI_VERSION 1
#define YOUTH_MAX_LEVELS 4
#define YOUTH_ICP_LANES 32    /* lanes per ICP run (reduction geometry, part of the spec) */
#define YOUTH_SUM_SLOTS 32    /* 21 JtJ + 6 Jtr + sum r^2 + inlier count + 3 duplicates */
/* Layout of the 32 sums (youth_cuda_debug_icp): 16 pairs sized for one packed
 * fma.rn.f32x2 each.  A[i][j] (i <= j) lives at: row 0 -> 0..5; row 1 -> 7..11; row 2 -> 12..15;
 * row 3 -> 17..19; row 4 -> 20..21; row 5 -> 23; slots 6, 16, 22 duplicate A10, A32, A54. */
#define YOUTH_SUMS_B0 24      /* J^T r: slots 24..29 */
#define YOUTH_SUMS_RR 30      /* sum of r^2          */
#define YOUTH_SUMS_COUNT 31   /* inlier count        */

typedef struct youth_cuda_handle youth_cuda_handle;

typedef struct youth_cuda_config {
  int32_t width, height;  /* level-0 image size; both divisible by 2^(levels-1), width % 8 == 0 */
  float fx, fy, cx, cy;   /* level-0 pinhole intrinsics (Camera.fx/fy/cx/cy)                  */
  float depth_factor;     /* raw units per metre (DepthMapFactor, 1000 for Astra)             */
  int32_t levels;         /* pyramid levels, 1..YOUTH_MAX_LEVELS                              */
  int32_t iters[YOUTH_MAX_LEVELS]; /* ICP iterations per level, index = level (0 = finest)    */
  int32_t depth_min_mm;   /* raw depth valid iff depth_min_mm <= d <= depth_max_mm            */
  int32_t depth_max_mm;
  int32_t bilateral;      /* 1: 7x7 bilateral filter on raw depth, 0: pass-through            */
  float sigma_space_px;   /* bilateral spatial sigma (pixels)                                 */
  float sigma_range_mm;   /* bilateral range sigma (raw units); pyramid gate = 3 sigma        */
  float dist_thresh_m;    /* correspondence rejection: |T v - v'| > dist_thresh               */
  float cos_thresh;       /* correspondence rejection: (R n) . n' < cos_thresh                */
  int32_t min_inliers;    /* an iteration with fewer inliers leaves the pose unchanged        */
  int32_t icp_ppt;        /* pixels per lane per ICP run at level 0 (power of two, 1..256); a run is
                             32*ppt pixels (strided sweep) reduced by one warp; level l uses max(1, ppt >> 2l) */
  int32_t n_streams;      /* independent sequences tracked by this handle (>= 1)              */
  int32_t batch;          /* max frames per sequence per youth_cuda_track_batch call (>= 1)   */
  int32_t traj_capacity;  /* max poses kept per sequence                                      */
  int32_t device;         /* CUDA device ordinal                                              */
  void* stream;           /* cudaStream_t to launch on; NULL = library-owned stream           */
} youth_cuda_config;

/* where the depth frames handed to youth_cuda_track_batch live */
#define YOUTH_MEM_HOST 0        /* pageable host memory: staged through a pinned buffer, copied before return */
#define YOUTH_MEM_DEVICE 1      /* device memory on cfg.device: consumed in place                              */
#define YOUTH_MEM_HOST_PINNED 2 /* page-locked host memory (youth_cuda_host_alloc): async H2D, caller keeps it
                                   alive until the next youth_cuda_sync / blocking call                        */

/* debug read-back selectors (parity tests) */
#define YOUTH_DBG_DEPTH 1  /* float[h*w]    filtered / downsampled depth in raw units, 0 = invalid */
#define YOUTH_DBG_VERTEX 2 /* float[h*w*4]  x,y,z,valid(1/0)                                      */
#define YOUTH_DBG_NORMAL 3 /* float[h*w*4]  nx,ny,nz,valid(1/0)                                   */
#define YOUTH_DBG_MASK 4   /* uint8[h*w]    bit0 vertex valid, bit1 normal valid                  */
#define YOUTH_DBG_PYRCNT 5 /* uint8[h*w]    samples averaged by the pyramid (level >= 1)          */

/* per-frame status bits (youth_cuda_get_status) */
#define YOUTH_STATUS_FIRST 1u    /* first frame of a sequence: pose = identity, no ICP        */
#define YOUTH_STATUS_LOST 2u     /* >= 1 iteration skipped (too few inliers / singular system) */

/* Fill *cfg with the Astra defaults: 640x480, fx=fy=570.3, cx=320, cy=240, factor 1000
 * (reference config/astra_orb_slam3_rgbd.yaml:9-20,35), 3 levels with 10/5/4 iterations
 * (fine->coarse), bilateral on (7x7, 4.5 px, 30 mm), 0.10 m / cos 20 deg gates. */
int youth_cuda_default_config(youth_cuda_config* cfg);

int youth_cuda_init(const youth_cuda_config* cfg, youth_cuda_handle** out);
void youth_cuda_destroy(youth_cuda_handle* h);

/* Track ONE frame of sequence 0 (host memory, copied before return).  pose_out (nullable)
 * receives the camera-to-world pose, row-major 3x4 [R|t], world = first camera frame;
 * when it is NULL the call returns after enqueueing the work. */
int youth_cuda_track(youth_cuda_handle* h, const uint16_t* depth_mm, uint32_t timestamp_ms,
                     float pose_out[12]);

/* Track n_frames (1..cfg.batch) consecutive frames of EVERY sequence in one launch group.
 * depth[s] points at n_frames tightly packed frames of sequence s (mem_kind says where).
 * timestamps_ms: [n_frames] shared by all sequences, nullable.  poses_out: nullable
 * [n_streams][n_frames][12]; non-NULL makes the call blocking. */
int youth_cuda_track_batch(youth_cuda_handle* h, const uint16_t* const* depth, int n_frames,
                           int mem_kind, const uint32_t* timestamps_ms, float* poses_out);

/* Launch schedule of stages 3-5 for groups of more than `pairs_per_group` frame pairs: pairs are
 * independent, so the whole coarse-to-fine iteration schedule of `pairs_per_group` pairs is run
 * before the next pairs are touched (their maps stay in L2 from iteration to iteration), groups
 * alternating over `queues` (1..8) CUDA streams.  pairs_per_group = 0 restores one launch per
 * iteration over all pairs.  Results do not depend on the schedule. */
int youth_cuda_set_icp_schedule(youth_cuda_handle* h, int pairs_per_group, int queues);

/* Block until everything enqueued on the handle has finished. */
int youth_cuda_sync(youth_cuda_handle* h);

/* Forget trajectory and previous frame of sequence `stream` (-1 = all). */
int youth_cuda_reset(youth_cuda_handle* h, int stream);

/* Number of frames tracked so far in sequence `stream`. */
int youth_cuda_frame_count(youth_cuda_handle* h, int stream);

/* Copy up to max_frames poses (12 floats each) / timestamps / status words of sequence
 * `stream`, starting at frame `first`; returns the number copied (blocking), -1 on error. */
int youth_cuda_get_trajectory(youth_cuda_handle* h, int stream, int first, int max_frames,
                              float* poses_out, uint32_t* timestamps_out, uint32_t* status_out);

/* Stream-ordered variant: enqueues the copies behind everything submitted so far and returns the number of
 * frames that will be copied (-1 on error) without waiting; *ticket_out names their completion.
 * poses_out / status_out should be page-locked (youth_cuda_host_alloc) for the copy to be asynchronous.
 * youth_cuda_wait_ticket blocks until that read-back has landed (tickets complete in order).  Keeping
 * two groups in flight -- submit g+1, then wait for g -- hides the host-to-device copy of g+1 under the
 * kernels of g. */
int youth_cuda_read_trajectory_async(youth_cuda_handle* h, int stream, int first, int max_frames,
                                     float* poses_out, uint32_t* status_out, uint64_t* ticket_out);
int youth_cuda_wait_ticket(youth_cuda_handle* h, uint64_t ticket);
/* Stream-ordered youth_cuda_last_inliers: covered by the ticket of a youth_cuda_read_trajectory_async
 * issued after it. */
int youth_cuda_read_last_inliers_async(youth_cuda_handle* h, int stream, int* inliers_out);

/* Inlier correspondences of the last tracked frame of `stream` at the finest level
 * (last iteration); blocking. */
int youth_cuda_last_inliers(youth_cuda_handle* h, int stream);

/* Device pointer of the trajectory of sequence `stream` (float[traj_capacity][12]) for
 * zero-copy hand-off to a collective (NCCL gather of per-sequence trajectories). */
void* youth_cuda_trajectory_device_ptr(youth_cuda_handle* h, int stream);

/* Device helpers for C hosts that do not link the CUDA runtime themselves (host/multi_gpu.c gives these buffers
 * to NCCL): device count, current device of the calling thread, device memory, blocking copy to the host, and a
 * whole-device synchronisation. */
int youth_cuda_device_count(void);
void* youth_cuda_stream(youth_cuda_handle* h); /* cudaStream_t the handle launches on (for stream-ordered collectives) */
int youth_cuda_set_device(int device);
void* youth_cuda_device_alloc(size_t bytes);
void youth_cuda_device_free(void* p);
int youth_cuda_copy_to_host(void* dst, const void* src_device, size_t bytes);
int youth_cuda_copy_to_device(void* dst_device, const void* src, size_t bytes);
int youth_cuda_device_sync(void);

/* Page-locked host memory for YOUTH_MEM_HOST_PINNED inputs. */
void* youth_cuda_host_alloc(size_t bytes);
void youth_cuda_host_free(void* p);

/* Parity hooks.  `frame` is the absolute frame index inside the sequence; it must still
 * be resident in the ring (the last cfg.batch+1 frames are).
 * YOUTH_DBG_DEPTH and YOUTH_DBG_PYRCNT read the filtered float depth pyramid and the pyramid sample counts, which
 * nothing on the product path consumes: the tracker only stores them (and only allocates their buffers) after
 * youth_cuda_debug_enable_maps(), to be called before the frames of interest are tracked (YOUTH_DEBUG_MAPS=1 in
 * the environment does the same at init; frame-to-model tracking enables them itself, its ray cast starts from
 * the depth pyramid).  Vertex / normal maps and masks are always readable. */
int youth_cuda_debug_enable_maps(youth_cuda_handle* h);
int youth_cuda_debug_read(youth_cuda_handle* h, int what, int stream, int frame, int level,
                          void* dst, size_t dst_bytes);
/* One association + reduction pass (stage 3+4) of frame `frame` against frame-1 at `level`
 * with the given prev<-cur pose; sums_out[YOUTH_SUM_SLOTS] (double), corr_out nullable
 * int32[h*w]: matched pixel index in the previous frame, or a negative reject code. */
int youth_cuda_debug_icp(youth_cuda_handle* h, int stream, int frame, int level,
                         const float pose[12], double* sums_out, int32_t* corr_out);

/* Stage 3 divides by v'.z through a reciprocal written out for positive normal floats (no range test,
 * no slow path).  Returns how many floats with bit patterns in [lo_bits, hi_bits] get a different
 * result from the correctly rounded reciprocal (must be 0 over the positive normals below 2^126),
 * -1 on error. */
long long youth_cuda_debug_rcp_check(youth_cuda_handle* h, uint32_t lo_bits, uint32_t hi_bits);

/* Elapsed device milliseconds between two marks on the handle's stream. */
int youth_cuda_timer_start(youth_cuda_handle* h);
int youth_cuda_timer_stop(youth_cuda_handle* h, float* ms_out);

/* Per-kernel-class device timing: while enabled every launch is bracketed by CUDA events on
 * the launching stream (serialises nothing, but adds event records, so throughput numbers
 * are taken with it off).  ms_out / launches_out: [YOUTH_PROF_CLASSES] totals since enable. */
#define YOUTH_PROF_INGEST 0  /* k_ingest: depth ingest + bilateral + pyramid + vertex maps */
#define YOUTH_PROF_NORMALS 1 /* k_normals */
#define YOUTH_PROF_ICP0 2    /* k_icp at level 0 (ICP0 + l = level l) */
#define YOUTH_PROF_SOLVE 6   /* (was k_solve: the solve moved into k_icp's last run) now k_tsdf_integrate */
#define YOUTH_PROF_INTEGRATE YOUTH_PROF_SOLVE
#define YOUTH_PROF_MISC 7    /* k_compose, YD16 unpacking */
#define YOUTH_PROF_RAYCAST 8 /* k_tsdf_raycast */
#define YOUTH_PROF_CLASSES 9
int youth_cuda_profile_enable(youth_cuda_handle* h, int on);
int youth_cuda_profile_read(youth_cuda_handle* h, double* ms_out, uint64_t* launches_out);

/* Kernels launched by this handle since init (for bench.py's gpu_launches). */
uint64_t youth_cuda_launch_count(youth_cuda_handle* h);

const char* youth_cuda_last_error(void);
int youth_cuda_abi_version(void);

/* correspondence reject codes written by youth_cuda_debug_icp */
#define YOUTH_REJ_CUR_INVALID (-1)
#define YOUTH_REJ_BEHIND (-2)
#define YOUTH_REJ_OUT_OF_IMAGE (-3)
#define YOUTH_REJ_PREV_INVALID (-4)
#define YOUTH_REJ_DISTANCE (-5)
#define YOUTH_REJ_ANGLE (-6)

#ifdef __cplusplus
}
#endif

#endif /* YOUTH_CUDA_H */
