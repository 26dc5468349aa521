// reduction.cuh -- launch interface of the image reductions.
#pragma once

#include "rmd_common.cuh"

namespace rmdb
{

// All reductions run on `stream`, leave the result in `result` (device memory,
// 16 bytes) and are deterministic: per-block partials are combined in block
// order by the last block to finish.
struct ReduceScratch
{
  double *partials;      // >= max_blocks * 2 doubles
  unsigned int *ticket;  // zero before the launch; reset by the kernel
  void *result;          // 16 bytes
  int max_blocks;
};

cudaError_t launch_sum_f32(const float *img, size_t stride, size_t w, size_t h,
                           const ReduceScratch &s, cudaStream_t stream);   // result: double
cudaError_t launch_sum_i32(const int *img, size_t stride, size_t w, size_t h,
                           const ReduceScratch &s, cudaStream_t stream);   // result: long long
cudaError_t launch_count_eq_i32(const int *img, size_t stride, size_t w, size_t h, int value,
                                const ReduceScratch &s, cudaStream_t stream);  // result: long long
cudaError_t launch_min_max_f32(const float *img, size_t stride, size_t w, size_t h,
                               const ReduceScratch &s, cudaStream_t stream);   // result: float2 (min, max)

} // namespace rmdb
