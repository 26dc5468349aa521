// denoiser.cuh -- launch interface of the TV-L1 denoiser kernels.
#pragma once

#include "rmd_common.cuh"

namespace rmdb
{

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
  // outputs, common stride in elements
  float2 *gmu;    // (g, noisy depth)
  float4 *state;  // (u, u_head, p.x, p.y)
  int stride;
};

struct DenoiseStepParams
{
  int width, height, stride;
  const float4 *in;
  float4 *out;
  const float2 *gmu;
  float sigma, tau, theta, lambda;
};

cudaError_t launch_denoise_setup(const DenoiseSetupParams &P, bool from_seeds, cudaStream_t stream);
cudaError_t launch_denoise_step(const DenoiseStepParams &P, cudaStream_t stream);

} // namespace rmdb
