/*
 * rmd_b200.h -- C-ABI of the Blackwell-native REMODE depth-filter hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ types, no
 * exceptions, no torch.  Every entry point replaces one member of the
 * reference's device-side classes (paths relative to the reference tree):
 *
 *   rmd_seeds_*      rmd::SeedMatrix        include/rmd/seed_matrix.cuh:45-109
 *                                           src/seed_matrix.cu:28-230
 *   rmd_denoiser_*   rmd::DepthmapDenoiser  include/rmd/depthmap_denoiser.cuh:27-54
 *                                           src/depthmap_denoiser.cu:124-229
 *   rmd_reduce_*     rmd::ImageReducer<T>   include/rmd/reduction.cuh:27-62
 *                                           src/reduction.cu:22-187
 *   rmd_image_*      rmd::DeviceImage<T>    include/rmd/device_image.cuh:34-180
 *
 * include/rmd/ holds header-compatible C++ classes (same names and signatures
 * as the reference's) that forward to these functions and turn non-zero
 * return codes back into rmd::CudaException, so rmd::Depthmap / the ROS node
 * compile against them unchanged.
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a cudaError_t value or
 *     one of the RMD_ERR_* codes below; rmd_last_error_string() describes the
 *     last failure of the calling thread;
 *   - host images are densely packed row-major (width*sizeof(T) pitch), gray
 *     value / 255 in [0,1] for float frames, as the reference requires
 *     (device_image.cuh:95-102, src/depthmap.cpp:105);
 *   - poses are SE3 3x4 row-major [R|t] float[12] (include/rmd/se3.cuh:27-142),
 *     world -> camera ("T_curr_world"), exactly what SeedMatrix takes;
 *   - a handle owns its device memory and a CUDA stream on the device it was
 *     created for; handles share no process-global state, so several can run
 *     concurrently on one GPU or one per GPU;
 *   - calls on one handle must come from one thread at a time (as with the
 *     reference: src/main_ros.cpp:43-48).
 */
#ifndef RMD_B200_H
#define RMD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMD_B200_ABI_VERSION 1

/* own error codes (cudaError_t values are small positive integers) */
#define RMD_ERR_INVALID_ARGUMENT (-1)
#define RMD_ERR_NOT_INITIALISED  (-2)  /* e.g. update before set_reference */
#define RMD_ERR_UNSUPPORTED      (-3)
#define RMD_ERR_DEVICE_WAIT       (-4)  /* a bounded device-side wait of a chained launch expired (a bug, not a state) */

/* rmd::ConvergenceStates, include/rmd/seed_matrix.cuh:31-43 */
enum rmd_convergence_state {
  RMD_UPDATE = 0,
  RMD_CONVERGED = 1,
  RMD_BORDER = 2,
  RMD_DIVERGED = 3,
  RMD_NO_MATCH = 4,
  RMD_NOT_VISIBLE = 5
};

/* Per-pixel fields of a seed matrix (download / upload / device_ptr). */
enum rmd_field {
  RMD_FIELD_MU = 0,                /* float   SeedMatrix::downloadDepthmap    seed_matrix.cu:160 */
  RMD_FIELD_SIGMA_SQ = 1,          /* float   downloadSigmaSq                 :206 */
  RMD_FIELD_A = 2,                 /* float   downloadA                       :210 */
  RMD_FIELD_B = 3,                 /* float   downloadB                       :214 */
  RMD_FIELD_CONVERGENCE = 4,       /* int32   downloadConvergence             :165 */
  RMD_FIELD_SUM_TEMPL = 5,         /* float   downloadSumTempl                :218 */
  RMD_FIELD_CONST_TEMPL_DENOM = 6, /* float   downloadConstTemplDenom         :222 */
  RMD_FIELD_EPIPOLAR_MATCHES = 7,  /* float2  downloadEpipolarMatches         :226 */
  RMD_FIELD_REF_IMG = 8,           /* float   the reference image as uploaded */
  RMD_FIELD_DEBUG_TIMELINE = 100   /* int64[16] per tile, see RMD_OPT_DEBUG_TIMELINE */
};

enum rmd_seeds_option {
  /* 1: keep the best epipolar match of every pixel (RMD_FIELD_EPIPOLAR_MATCHES);
   * off by default because nothing downstream of update() reads it
   * (it exists in the reference for tests: seed_matrix.cuh:76-83). */
  RMD_OPT_RECORD_MATCHES = 0,
  /* 0 (default): staged kernel (TMA -> shared memory, balanced candidate work
   * list), 1: direct kernel (one thread per pixel, global loads); same
   * arithmetic, bit-identical results. */
  RMD_OPT_KERNEL_VARIANT = 1,
  /* fractional bits of the bilinear weights of the current-image taps
   * (8 = what the texture unit of the reference path uses; 0 = exact fp32). */
  RMD_OPT_TEX_FRAC_BITS = 2,
  /* debug: the staged kernel records 16 x int64 per 32x8 tile (clock64 at its
   * phase boundaries, SM id, work items, strip coverage ...); read them back
   * with rmd_seeds_download(h, RMD_FIELD_DEBUG_TIMELINE, dst) where dst holds
   * 16 * ceil(w/32) * ceil(h/8) int64 values. */
  RMD_OPT_DEBUG_TIMELINE = 3,
  /* 1: a host frame given to rmd_seeds_update* / set_reference* that lies in
   * page-locked memory (cudaHostAlloc / cudaHostRegister, detected with
   * cudaPointerGetAttributes) is DMA'd straight from the caller's buffer:
   * no staging copy into the library's pinned ring.  The caller then must
   * leave the frame untouched until rmd_seeds_sync() (or any download)
   * returns -- the reference's "reusable on return" guarantee
   * (device_image.cuh:93-106) no longer holds for such buffers.  Pageable
   * buffers still take the staged path.  Default 0. */
  RMD_OPT_PINNED_INPUT = 4,
  /* Frames per launch of rmd_seeds_update_device_batch (1..8, default 1): with n > 1, n consecutive frames of
   * the keyframe are chained inside ONE persistent launch -- a tile moves on to frame k+1 as soon as its own
   * frame k is final, so frames overlap on the GPU.  Same results as one launch per frame.  Opt-in: it paid
   * 9 % with the 80-register kernel at VGA, nothing with the 128-register kernel that is now the default, and it
   * costs 8-25 % at 720p / 1080p (profiles/r02_occupancy_ab.txt). */
  RMD_OPT_CHAIN_FRAMES = 5,
  /* Seed-major mode: once at most this percentage of the pixels is still being updated (0 = never, the default)
   * the handle keeps the live seeds as a compact list and a launch walks every listed seed through its frames
   * warp by warp -- no tiles, no per-frame synchronisation (csrc/depth_filter_seeds.cu).  Same results.  Off by
   * default: on the bench workloads 10-20 % of the seeds stay live (NO_MATCH seeds keep searching) and at that
   * density the tile organisation is faster (profiles/r02_tune_probe.txt); it pays for sparser live sets. */
  RMD_OPT_SEED_MODE_PCT = 6,
  /* tuning knobs of the staged kernel's busy-tile splitting and sparse-tile
   * path (defaults in csrc/staged_maps.cuh); results never depend on them */
  RMD_OPT_TUNE_SPLIT_MAX = 10,            /* most CTAs sharing one busy tile (1 = never split; default 16) */
  RMD_OPT_TUNE_SPLIT_MIN_ITEMS = 11,      /* tiles with fewer work items are never split (512) */
  RMD_OPT_TUNE_SPLIT_ITEMS_PER_CTA = 12,  /* least work items a CTA of a split tile gets (384) */
  RMD_OPT_TUNE_SPARSE_MAX_SEEDS = 13,     /* tiles with at most this many seeds to update skip TMA staging (16; 0 = off) */
  RMD_OPT_TUNE_HEAVY_MIN_ITEMS = 14,      /* tiles with at least this many items are dispatched first (128) */
  RMD_OPT_TUNE_SPLIT_AVG_PCT = 15,        /* target items per CTA of a split tile, in % of the frame's items per resident CTA slot (50) */
  RMD_OPT_TUNE_PDL = 16,           /* 1 (default): programmatic dependent launch of consecutive frames */
  RMD_OPT_TUNE_WARP_TILE_SEEDS = 17, /* tiles with at most this many seeds to update (and a few dozen candidates) are processed by one warp, eight tiles per CTA (8 = the most; 0 = off) */
  RMD_OPT_TUNE_GRID_CTAS = 18,      /* size of the persistent grid (0, the default: one CTA per resident slot, SMs x occupancy) */
  RMD_OPT_TUNE_CTAS_PER_SM = 19,    /* 5x5 staged kernel: the 128-register build (2, = 0, the default) or the 80-register build (3; sized for 3 CTAs per SM, of which two fit next to the 60 KB strips) */
  RMD_OPT_TUNE_WARP_TILE_CANDS = 20  /* ... and at most this many candidates in all (64 = the most) */
};

typedef struct rmd_seeds rmd_seeds_t;
typedef struct rmd_denoiser rmd_denoiser_t;
typedef struct rmd_multi rmd_multi_t;

/* ------------------------------------------------------------------ misc */
int rmd_abi_version(void);
const char *rmd_last_error_string(void);
int rmd_device_count(int *count);

/* ----------------------------------------------------------- seed matrix */

/* SeedMatrix::SeedMatrix(width, height, PinholeCamera(fx,fy,cx,cy))
 * seed_matrix.cu:28-80.  patch_side = RMD_CORR_PATCH_SIDE (5 or 7; a compile
 * time macro in the reference, CMakeLists.txt:51).  device < 0: current. */
int rmd_seeds_create(int width, int height, float fx, float fy, float cx,
                     float cy, int patch_side, int device, rmd_seeds_t **out);
int rmd_seeds_destroy(rmd_seeds_t *s);

/* Run this handle's work on a caller-owned cudaStream_t (NULL = back to the
 * handle's own stream).  The reference uses the legacy default stream only. */
int rmd_seeds_set_stream(rmd_seeds_t *s, void *cuda_stream);
int rmd_seeds_get_stream(rmd_seeds_t *s, void **cuda_stream);
int rmd_seeds_set_option(rmd_seeds_t *s, int option, int value);

/* SeedMatrix::setReferenceImage, seed_matrix.cu:87-118.  Host buffer. */
int rmd_seeds_set_reference(rmd_seeds_t *s, const float *host_img,
                            const float *T_curr_world, float min_depth,
                            float max_depth);
/* Same with the image already in device memory (pitch in bytes). */
int rmd_seeds_set_reference_device(rmd_seeds_t *s, const float *dev_img,
                                   size_t pitch_bytes,
                                   const float *T_curr_world, float min_depth,
                                   float max_depth);
/* Same from an 8-bit gray frame; value * (1/255.f) is applied on the GPU
 * (what rmd::Depthmap::inputImage does on the CPU, src/depthmap.cpp:105). */
int rmd_seeds_set_reference_u8(rmd_seeds_t *s, const uint8_t *host_img,
                               const float *T_curr_world, float min_depth,
                               float max_depth);

/* SeedMatrix::update, seed_matrix.cu:120-158: convergence check, epipolar NCC
 * search, triangulation and Bayesian update of every seed -- one fused kernel.
 * The host buffer may be reused as soon as the call returns (it is copied to
 * a pinned ring); the GPU work is enqueued on the handle's stream and, like
 * the reference's last kernel, may still be running on return. */
int rmd_seeds_update(rmd_seeds_t *s, const float *host_img,
                     const float *T_curr_world);
int rmd_seeds_update_u8(rmd_seeds_t *s, const uint8_t *host_img,
                        const float *T_curr_world);

/* Several live reference keyframes against one incoming frame (SURVEY.md 8f
 * row 2; the reference node keeps a single rmd::Depthmap and re-keyframes,
 * src/depthmap_node.cpp:125-157).  The frame is staged and uploaded ONCE
 * (through handles[0]) and the keyframes are updated by ONE launch per group
 * of up to 8 (on handles[0]'s stream; the other handles' streams wait for
 * it): the launch's work list is the concatenation of the keyframes' lists,
 * so their dependent chains interleave from the first cycle -- steady frames
 * of one keyframe leave most issue slots idle.  Handles on
 * the direct variant or with another patch size are enqueued one by one.
 * Same result, bit for bit, as calling rmd_seeds_update[_u8] on each handle.
 * All handles must have the same image size and device and a reference
 * frame; at most 64 per call. */
int rmd_seeds_update_many(rmd_seeds_t *const *handles, int n, const float *host_img,
                          const float *T_curr_world);
int rmd_seeds_update_many_u8(rmd_seeds_t *const *handles, int n, const uint8_t *host_img,
                             const float *T_curr_world);

/* Frame ingest with lens undistortion.
 * rmd::Depthmap::initUndistortionMap(k1, k2, r1, r2), src/depthmap.cpp:45-61:
 * builds the fixed-point maps of cv::initUndistortRectifyMap(K, D, I, K, size,
 * CV_16SC2) for the handle's camera.  From then on the *_u8 entry points
 * (set_reference_u8, update_u8) run rmd::Depthmap::inputImage
 * (src/depthmap.cpp:95-106) on the GPU, fused in one kernel: cv::remap(...,
 * CV_INTER_LINEAR) of the 8-bit frame, then convertTo(CV_32F, 1/255.f).  Float
 * entry points are not affected (the reference's SeedMatrix takes undistorted
 * float images).  k1 = k2 = r1 = r2 = 0 is still a remap (identity up to the
 * map's rounding); rmd_seeds_clear_undistortion_map() switches it off. */
int rmd_seeds_init_undistortion_map(rmd_seeds_t *s, float k1, float k2, float r1, float r2);
int rmd_seeds_clear_undistortion_map(rmd_seeds_t *s);
/* The maps as OpenCV lays them out: xy = CV_16SC2 (2*w*h int16: x, y of the
 * top-left source pixel), frac = CV_16UC1 (w*h: (fy << 5) | fx, 5-bit
 * fractions).  Either pointer may be NULL. */
int rmd_seeds_get_undistortion_map(rmd_seeds_t *s, int16_t *host_xy, uint16_t *host_frac);
/* img_undistorted_8uc1_ of rmd::Depthmap (src/depthmap.cpp:99): the remapped
 * 8-bit frame itself, host to host (the reference keeps it as the intensity
 * source of the published point cloud).  Synchronous. */
int rmd_seeds_undistort_u8(rmd_seeds_t *s, const uint8_t *host_src, uint8_t *host_dst);
/* Frame already resident in device memory (must stay valid until the stream
 * has consumed it).  pitch_bytes must be a multiple of 16. */
int rmd_seeds_update_device(rmd_seeds_t *s, const float *dev_img,
                            size_t pitch_bytes, const float *T_curr_world);

/* n_frames consecutive updates from frames resident in device memory:
 * frame i starts at dev_frames + i*frame_stride_bytes (pitch_bytes per row),
 * its pose is T_curr_world + 12*i.  One call, n_frames fused launches, no
 * host round trip in between (the sequence is strictly ordered: frame k+1
 * searches around the posterior frame k left, epipolar_match.cu:60-75). */
int rmd_seeds_update_device_batch(rmd_seeds_t *s, const float *dev_frames,
                                  size_t frame_stride_bytes,
                                  size_t pitch_bytes, int n_frames,
                                  const float *T_curr_world);

int rmd_seeds_sync(rmd_seeds_t *s);

/* download* accessors: dst is width*height elements, densely packed. */
int rmd_seeds_download(rmd_seeds_t *s, int field, void *host_dst);
/* Test / checkpoint hook the reference lacks: overwrite one field
 * (MU, SIGMA_SQ, A, B, CONVERGENCE). */
int rmd_seeds_upload_state(rmd_seeds_t *s, int field, const void *host_src);
/* getMu()/getSigmaSq()/getA()/getB()/getConvergence(): a pitched planar
 * device image of the field, valid until the next call on this handle. */
int rmd_seeds_device_ptr(rmd_seeds_t *s, int field, void **dev_ptr,
                         size_t *pitch_bytes);
/* Copy a field into caller-owned device memory, ASYNCHRONOUSLY on the handle's
 * stream: a consumer on any other stream (including the legacy default stream
 * of rmd_reduce_* and the denoiser's own stream) must rmd_seeds_sync() first. */
int rmd_seeds_copy_field_to_device(rmd_seeds_t *s, int field, void *dev_dst,
                                   size_t dst_pitch_bytes);

/* SeedMatrix::getConvergedCount, seed_matrix.cu:195-198: number of pixels the
 * last update() classified CONVERGED (the fused kernel counts them; no extra
 * pass).  Before the first update it is 0. */
int rmd_seeds_converged_count(rmd_seeds_t *s, size_t *count);
/* SeedMatrix::getDistFromRef, seed_matrix.cu:200-203. */
int rmd_seeds_dist_from_ref(rmd_seeds_t *s, float *dist);
int rmd_seeds_size(rmd_seeds_t *s, int *width, int *height, int *patch_side);
/* Number of fused depth-filter kernels / all kernels this handle launched. */
int rmd_seeds_launch_count(rmd_seeds_t *s, uint64_t *fused, uint64_t *total);
/* Device time of the last fused depth-filter kernel in milliseconds, measured
 * with CUDA events on the handle's stream (blocks until it finished). */
int rmd_seeds_last_kernel_ms(rmd_seeds_t *s, float *ms);
int rmd_seeds_enable_kernel_timing(rmd_seeds_t *s, int on);
/* Debug: host-side time of the streaming ingest, accumulated over all handles
 * of the process when the environment variable RMD_HOST_PROFILE is set:
 * out[0] wait for a free ring slot, [1] copy into pinned memory, [2] enqueue
 * the host-to-device copy, [3] encode TMA descriptors, [4] launch,
 * [5] whole rmd_seeds_update* call (all seconds), [6] number of calls.
 * reset != 0 clears the accumulators after reading. */
int rmd_debug_host_profile(double out[8], int reset);

/* Point-cloud extraction.
 * rmd::Publisher::publishPointCloud, src/publisher.cpp:54-86: every pixel whose
 * state is CONVERGED becomes a point  T_world_ref * (normalize((x-cx)/fx,
 * (y-cy)/fy, 1) * depth(y, x))  with the 8-bit intensity of the reference image,
 * in row-major pixel order.  The reference downloads the depth and convergence
 * maps and loops on the CPU; here the points are compacted on the device (same
 * order, IEEE arithmetic in the reference's operation order) and only they are
 * copied out.  `dev_depth` is a pitched device image (e.g. the denoised map of
 * rmd_denoiser_run_seeds_to_device) or NULL for the seeds' own depth estimate.
 * Points are 4 floats (x, y, z, intensity).  *count is always the number of
 * CONVERGED pixels; at most `capacity_points` points are written.  Synchronous. */
int rmd_seeds_point_cloud(rmd_seeds_t *s, const float *dev_depth, size_t depth_pitch_bytes,
                          float *host_xyzi, size_t capacity_points, size_t *count);
/* Same into device memory (16-byte aligned). */
int rmd_seeds_point_cloud_device(rmd_seeds_t *s, const float *dev_depth, size_t depth_pitch_bytes,
                                 float *dev_xyzi, size_t capacity_points, size_t *count);

/* -------------------------------------------------------------- denoiser */

/* DepthmapDenoiser(width, height), depthmap_denoiser.cu:143-177 */
int rmd_denoiser_create(int width, int height, int device,
                        rmd_denoiser_t **out);
int rmd_denoiser_destroy(rmd_denoiser_t *d);
int rmd_denoiser_set_stream(rmd_denoiser_t *d, void *cuda_stream);
/* setLargeSigmaSq(depth_range), :226-229 */
int rmd_denoiser_set_large_sigma_sq(rmd_denoiser_t *d, float depth_range);
/* denoise(mu, sigma_sq, a, b, host_denoised, lambda, iterations), :179-224.
 * Inputs are pitched planar float images in device memory (pitch in bytes).
 * Returns RMD_ERR_NOT_INITIALISED if set_large_sigma_sq was never called
 * (the reference prints to stderr and returns, :189-193). */
int rmd_denoiser_run(rmd_denoiser_t *d, const float *mu, size_t mu_pitch,
                     const float *sigma_sq, size_t sigma_sq_pitch,
                     const float *a, size_t a_pitch, const float *b,
                     size_t b_pitch, float *host_denoised, float lambda,
                     int iterations);
/* Same, reading the seed state of `s` directly (no planar export). */
int rmd_denoiser_run_seeds(rmd_denoiser_t *d, rmd_seeds_t *s,
                           float *host_denoised, float lambda, int iterations);
/* Same, leaving the result in caller-owned device memory (no D2H copy).
 * Asynchronous on the denoiser's stream.  The seeds handle is ordered after it:
 * a following rmd_seeds_point_cloud(s, dev_out, ...), update or set_reference
 * on `s` waits (on the device) for the denoised map / for the read of the seed
 * state.  Any OTHER consumer of dev_out must rmd_denoiser_sync() first. */
int rmd_denoiser_run_seeds_to_device(rmd_denoiser_t *d, rmd_seeds_t *s,
                                     float *dev_out, size_t out_pitch_bytes,
                                     float lambda, int iterations);
int rmd_denoiser_sync(rmd_denoiser_t *d);
int rmd_denoiser_launch_count(rmd_denoiser_t *d, uint64_t *total);

/* ------------------------------------------------------------ reductions */

/* ImageReducer<T>::sum / countEqual, reduction.cu:81-184.  Device pointer,
 * stride in ELEMENTS (as the reference), legacy default stream, blocking. */
int rmd_reduce_sum_f32(const float *dev_img, size_t stride, size_t width,
                       size_t height, float *out);
int rmd_reduce_sum_i32(const int32_t *dev_img, size_t stride, size_t width,
                       size_t height, int32_t *out);
int rmd_reduce_count_eq_i32(const int32_t *dev_img, size_t stride,
                            size_t width, size_t height, int32_t value,
                            size_t *out);
/* extras the north star asks for (not in the reference) */
int rmd_reduce_min_max_f32(const float *dev_img, size_t stride, size_t width,
                           size_t height, float *out_min, float *out_max);

/* -------------------------------------------------------------- multi-GPU */

/* Independent reference keyframes, one rmd_seeds_t per GPU.  The
 * depth filter never reads a neighbour's or another keyframe's state
 * (src/seed_check.cu, src/epipolar_match.cu, src/seed_update.cu), so there is
 * no data-path collective: the only exchange is the FINAL GATHER of every
 * keyframe's depth (f32) and convergence (i32) map to one root GPU -- grouped
 * ncclSend / ncclRecv over NVLink.  The reference has no multi-GPU path (one
 * SeedMatrix on the current device, src/check_cuda_device.cu:109).  NCCL
 * (libnccl.so.2) is loaded at run time; without it these functions return
 * RMD_ERR_UNSUPPORTED.
 *
 * One process driving n GPUs (e.g. a node owning several rmd::Depthmap
 * objects): rank i <-> devices[i]; ncclCommInitAll. */
int rmd_multi_create(const int *devices, int n, int width, int height, rmd_multi_t **out);
/* One process per GPU: rank 0 calls rmd_multi_unique_id, ships the 128 bytes to
 * the other ranks by its own means, every rank calls rmd_multi_create_rank
 * (collective: ncclCommInitRank).  device < 0: current. */
int rmd_multi_unique_id(char id[128]);
int rmd_multi_create_rank(const char id[128], int n_ranks, int rank, int device, int width,
                          int height, rmd_multi_t **out);
int rmd_multi_destroy(rmd_multi_t *m);
int rmd_multi_size(rmd_multi_t *m, int *n_ranks, int *n_local, int *first_rank);
/* The final gather.  seeds[i]: keyframe of local member i (n entries after
 * rmd_multi_create, one after rmd_multi_create_rank), whose queued updates are
 * waited for on the device.  dev_depth (may be NULL, entries may be NULL): a
 * pitched device image to send instead of the seeds' own depth estimate, e.g.
 * the denoised map of rmd_denoiser_run_seeds_to_device (complete when this is
 * called: rmd_denoiser_sync).  On the process that holds rank `root`,
 * host_depth / host_conv receive n_ranks * width * height elements each, rank
 * after rank; elsewhere they may be NULL.  Collective over all ranks; returns
 * when the exchange has finished. */
int rmd_multi_gather_maps(rmd_multi_t *m, rmd_seeds_t *const *seeds,
                          const float *const *dev_depth, const size_t *dev_depth_pitch,
                          int root, float *host_depth, int32_t *host_conv);

/* ---------------------------------------------------------- device image */

/* DeviceImage<T>(width,height) = cudaMallocPitch, device_image.cuh:37-50 */
int rmd_image_alloc(size_t width, size_t height, size_t elem_size,
                    void **dev_ptr, size_t *pitch_bytes);
int rmd_image_free(void *dev_ptr);
/* setDevData / getDevData / zero / operator=, device_image.cuh:93-171 */
int rmd_image_upload(void *dev_ptr, size_t pitch_bytes, const void *host_src,
                     size_t width, size_t height, size_t elem_size);
int rmd_image_download(const void *dev_ptr, size_t pitch_bytes, void *host_dst,
                       size_t width, size_t height, size_t elem_size);
int rmd_image_zero(void *dev_ptr, size_t pitch_bytes, size_t width,
                   size_t height, size_t elem_size);
int rmd_image_copy(void *dst, size_t dst_pitch, const void *src,
                   size_t src_pitch, size_t width, size_t height,
                   size_t elem_size);

#ifdef __cplusplus
}
#endif
#endif /* RMD_B200_H */
