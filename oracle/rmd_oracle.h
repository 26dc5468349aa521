/*
 * rmd_oracle.h -- CPU ORACLE for the REMODE depth-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 * The product (rpg_open_remode_b200/librmd_b200.so) never links or calls it.
 *
 * It is a plain-C, IEEE-fp32 restatement of the reference's CUDA kernels
 * (file:line cited at each function in rmd_oracle.c).  The reference has no
 * CPU implementation and stores no golden outputs, so
 * this oracle is pinned two ways:
 *   1. against the reference's own known-answer tests re-hosted on synthetic
 *      frames (seedMatrixInit / seedMatrixCheck / epipolarMatchTest /
 *      reduction_test), tests/test_oracle_pins.py, CPU;
 *   2. against the reference's own, unmodified CUDA kernels rebuilt for
 *      sm_100a (oracle/_ref/librmd_ref.so, recipe oracle/Makefile), on the
 *      GPU, tests/test_ref_cuda_parity.py.
 */
#ifndef RMD_ORACLE_H
#define RMD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numbering as rmd::ConvergenceStates, include/rmd/seed_matrix.cuh:31-43 */
enum {
  RMD_O_UPDATE = 0,
  RMD_O_CONVERGED = 1,
  RMD_O_BORDER = 2,
  RMD_O_DIVERGED = 3,
  RMD_O_NO_MATCH = 4,
  RMD_O_NOT_VISIBLE = 5
};

typedef struct rmd_oracle_seeds rmd_oracle_seeds;

/* patch = RMD_CORR_PATCH_SIDE (5 or 7, any odd >= 3). */
rmd_oracle_seeds *rmd_oracle_seeds_create(int width, int height, float fx,
                                          float fy, float cx, float cy,
                                          int patch);
void rmd_oracle_seeds_destroy(rmd_oracle_seeds *s);

/* Texture-unit model for the bilinear taps of the current image:
 * frac_bits = number of fractional bits the filter weights are quantised to
 * (8 on NVIDIA hardware); frac_bits <= 0 means exact fp32 weights. */
void rmd_oracle_seeds_set_tex_model(rmd_oracle_seeds *s, int frac_bits);
void rmd_oracle_set_threads(int n);
int rmd_oracle_get_threads(void);

/* img: densely packed row-major float [h*w] in [0,1]; T: SE3 3x4 row major,
 * world -> camera (the reference's T_curr_world). */
int rmd_oracle_seeds_set_reference(rmd_oracle_seeds *s, const float *img,
                                   const float *T_curr_world, float min_depth,
                                   float max_depth);
int rmd_oracle_seeds_update(rmd_oracle_seeds *s, const float *img,
                            const float *T_curr_world);

/* Stage-level entry points (one reference kernel each) for kernel-level
 * parity tests.  They operate on the state held in s. */
void rmd_oracle_stage_check(rmd_oracle_seeds *s);
void rmd_oracle_stage_match(rmd_oracle_seeds *s, const float *curr_img,
                            const float *T_curr_ref);
void rmd_oracle_stage_update(rmd_oracle_seeds *s, const float *T_ref_curr);

/* field ids shared with the product's C-ABI (include/rmd_b200.h) */
enum {
  RMD_O_F_MU = 0,
  RMD_O_F_SIGMA_SQ = 1,
  RMD_O_F_A = 2,
  RMD_O_F_B = 3,
  RMD_O_F_CONVERGENCE = 4, /* int32 */
  RMD_O_F_SUM_TEMPL = 5,
  RMD_O_F_CONST_TEMPL_DENOM = 6,
  RMD_O_F_EPIPOLAR_MATCHES = 7, /* float2 */
  RMD_O_F_REF_IMG = 8
};
/* Direct pointer to the densely packed field (w*h elements; float2 for 7). */
void *rmd_oracle_seeds_field(rmd_oracle_seeds *s, int field);
size_t rmd_oracle_seeds_converged_count(const rmd_oracle_seeds *s);
float rmd_oracle_seeds_dist_from_ref(const rmd_oracle_seeds *s);
void rmd_oracle_seeds_T_curr_ref(const rmd_oracle_seeds *s, float *out12);

/* TV-L1 denoiser, deterministic Jacobi two-phase semantics (see .c). */
int rmd_oracle_denoise(const float *mu, const float *sigma_sq, const float *a,
                       const float *b, int width, int height,
                       float depth_range, float lambda, int iterations,
                       float *out);

/* Reductions.  *_ref_order reproduce the reference's 4x4 grid of 16x16 blocks
 * + tree order exactly; *_f64 is the double-accumulated value the reference's
 * test compares with. */
float rmd_oracle_sum_f32_ref_order(const float *img, size_t stride, size_t w,
                                   size_t h);
double rmd_oracle_sum_f32_f64(const float *img, size_t stride, size_t w,
                              size_t h);
int rmd_oracle_sum_i32(const int *img, size_t stride, size_t w, size_t h);
size_t rmd_oracle_count_equal_i32(const int *img, size_t stride, size_t w,
                                  size_t h, int value);

/* SE3 helpers (include/rmd/se3.cuh:81-97,146-162) exposed for host-logic
 * tests. */
void rmd_oracle_se3_inv(const float *T, float *out);
void rmd_oracle_se3_mul(const float *A, const float *B, float *out);
void rmd_oracle_se3_from_quat(float qw, float qx, float qy, float qz, float tx,
                              float ty, float tz, float *out);

/* Frame ingest (rmd_oracle_ingest.c): rmd::Depthmap::initUndistortionMap and
 * inputImage, src/depthmap.cpp:45-61,95-106 -- OpenCV's initUndistortRectifyMap
 * (CV_16SC2 maps), remap (INTER_LINEAR, 8-bit) and convertTo(CV_32F, 1/255.f). */
void rmd_oracle_undistort_maps(int width, int height, float fx, float fy, float cx,
                               float cy, float k1, float k2, float p1, float p2,
                               int16_t *map1, uint16_t *map2);
void rmd_oracle_remap_u8(const uint8_t *src, int width, int height, const int16_t *map1,
                         const uint16_t *map2, uint8_t *dst);
void rmd_oracle_u8_to_float(const uint8_t *src, size_t n, float *dst);

/* Point cloud (rmd_oracle_pointcloud.c): rmd::Publisher::publishPointCloud,
 * src/publisher.cpp:54-86 -- dense maps in, 4 floats per CONVERGED pixel out
 * (NULL: count only); returns the number of points. */
size_t rmd_oracle_point_cloud(const float *depth, const int *conv, const uint8_t *ref_u8,
                              int width, int height, float fx, float fy, float cx, float cy,
                              const float *T_world_ref, float *out);

#ifdef __cplusplus
}
#endif
#endif
