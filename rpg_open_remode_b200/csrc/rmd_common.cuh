// rmd_common.cuh -- shared definitions of the sm_100a depth-filter library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/rmd_b200.h"

namespace rmdb
{

// ---------------------------------------------------------------- errors
void set_last_error(const std::string &msg);
int fail(int code, const char *what);
int fail_cuda(cudaError_t err, const char *what);

#define RMD_CUDA_TRY(expr)                                   \
  do {                                                       \
    const cudaError_t rmd_err_ = (expr);                     \
    if(rmd_err_ != cudaSuccess)                              \
      return ::rmdb::fail_cuda(rmd_err_, #expr);             \
  } while(0)

#define RMD_REQUIRE(cond, what)                              \
  do {                                                       \
    if(!(cond))                                              \
      return ::rmdb::fail(RMD_ERR_INVALID_ARGUMENT, what);   \
  } while(0)

// RAII device switch: handles are pinned to the device they were created on.
struct DeviceGuard
{
  explicit DeviceGuard(int device) : prev_(-1), changed_(false)
  {
    if(cudaGetDevice(&prev_) == cudaSuccess && prev_ != device)
    {
      changed_ = (cudaSetDevice(device) == cudaSuccess);
    }
  }
  ~DeviceGuard()
  {
    if(changed_) cudaSetDevice(prev_);
  }
  int prev_;
  bool changed_;
};

// ------------------------------------------------------------ geometry
// SE3 as 12 floats, row-major 3x4 [R|t] (the reference's layout,
// include/rmd/se3.cuh:142 "Matrix<Type,3,4> data").
struct Pose
{
  float m[12];
};

inline Pose pose_from(const float *v)
{
  Pose p;
  for(int i = 0; i < 12; ++i) p.m[i] = v[i];
  return p;
}
Pose pose_inverse(const Pose &T);                 // se3.cuh:81-97
Pose pose_compose(const Pose &a, const Pose &b);  // se3.cuh:146-162

struct Camera
{
  float fx, fy, cx, cy;
};

// Everything the fused depth-filter kernel needs, passed by value as a
// __grid_constant__ parameter (replaces the reference's mvs::DeviceData block
// that lives in device memory and is pointer-chased by every thread,
// include/rmd/mvs_device_data.cuh:47-106).
struct FilterParams
{
  int width, height;
  // seed state, array-of-structures: one float4 (mu, sigma_sq, a, b) per pixel
  float4 *seed;
  int seed_stride;      // in float4 elements
  // NCC template statistics: float2 (sum_templ, const_templ_denom)
  const float2 *templ;
  int templ_stride;     // in float2 elements
  int *conv;            // ConvergenceState per pixel
  int conv_stride;      // in ints
  const float *ref;     // reference image
  int ref_stride;       // in floats
  const float *curr;    // current image
  int curr_stride;      // in floats
  float2 *matches;      // optional (RMD_OPT_RECORD_MATCHES), may be null
  int match_stride;     // in float2 elements
  Camera cam;
  Pose T_curr_ref;
  Pose T_ref_curr;
  float eta_inlier, eta_outlier, epsilon;
  float depth_range;
  float one_pix_angle;
  float tex_quant;      // 2^frac_bits of the bilinear weights, 0 = exact
  int trust_conv;       // 1: absorbing states recorded in conv are final
  unsigned int *converged_now;   // counter of this frame
  unsigned int *converged_next;  // counter to clear for the next frame
  long long *timeline;           // debug: 8 clock64() stamps per CTA of the staged kernel, or null
  // staged kernel, busy-tile splitting (depth_filter_staged.cu)
  int split_max;                         // grid.z; 1 = never split
  int split_min_items, split_items_per_cta, sparse_max_seeds;  // tuning (staged_maps.cuh defaults)
  int cta_slots;                         // resident CTAs of the whole GPU (SMs x CTAs per SM)
  unsigned long long *tile_keys;         // [tiles][256] partial arg-max keys of split tiles
  unsigned int *tile_arrivals;           // [tiles] CTAs of a split tile that finished searching
  int n_tiles, tiles_x;                  // tile grid of the image
  const int *tile_zeff_cur;              // [tiles] CTAs sharing each tile in this frame (0/1 = one)
  int *tile_zeff_next;                   // [tiles] ... decided now for the next frame
  const unsigned int *helper_list_cur;   // helper CTAs of this frame: tile | share << 20 | zeff << 26
  unsigned int *helper_list_next;
  const unsigned int *helper_count_cur;  // entries of helper_list_cur
  unsigned int *helper_count_next, *helper_count_zero;
  int helper_cap;                        // capacity of the helper lists = helper CTAs launched
  const unsigned int *frame_items_prev;  // work items of the whole previous frame
  unsigned int *frame_items_next;        // ... accumulated this frame
  unsigned int *frame_items_zero;        // slot to clear for the next frame
};

} // namespace rmdb
