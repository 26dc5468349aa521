// denoiser.cuh -- launch interface of the TV-L1 denoiser kernels.
#pragma once

#include <cuda.h>

#include "rmd_common.cuh"

namespace rmdb
{

// Temporal blocking of the primal-dual iteration (denoiser.cu): iterations per launch, output tile,
// and the staged block (tile + one halo pixel per iteration).
constexpr int DENOISE_T = 8;
constexpr int DENOISE_TILE_W = 48, DENOISE_TILE_H = 24;
constexpr int DENOISE_EXT_W = DENOISE_TILE_W + 2 * DENOISE_T;   // 64
constexpr int DENOISE_EXT_H = DENOISE_TILE_H + 2 * DENOISE_T;   // 40

struct DenoiseSetupParams
{
  int width, height;
  // input A: planar pitched images (strides in floats)
  const float *mu; int mu_stride;
  const float *sigma_sq; int sigma_sq_stride;
  const float *a; int a_stride;
  const float *b; int b_stride;
  // input B: the seed matrix' own float4 records
  const float4 *seed; int seed_stride;
  float large_sigma_sq;
  // outputs: planar images of the solver, common stride in floats
  float *g, *noisy;            // weight, noisy depth (constant over the iterations)
  float *u, *u_head, *p_x, *p_y;
  int stride;
};

// One launch = n_it <= DENOISE_T Jacobi iterations from the `in_*` planes (TMA descriptors with box
// DENOISE_EXT_W x DENOISE_EXT_H) to the `out_*` planes.
struct alignas(64) DenoiseBlockParams
{
  CUtensorMap in_u, in_uh, in_px, in_py, g, mu;
  float *out_u, *out_uh, *out_px, *out_py;
  int width, height, stride;
  int n_it;
  float sigma, tau, theta, lambda;
};

cudaError_t launch_denoise_setup(const DenoiseSetupParams &P, bool from_seeds, cudaStream_t stream);
cudaError_t launch_denoise_block(const DenoiseBlockParams &P, cudaStream_t stream);
size_t denoise_block_smem_bytes();

// cuTensorMapEncodeTiled for a pitched 2-D float image (depth_filter_staged.cu); returns 0 or an error code
// and sets the thread's last error string.
int encode_tensor_map_2d_f32(CUtensorMap *map, const void *base, int width, int height, int stride_floats,
                             int box_w, int box_h);

} // namespace rmdb
