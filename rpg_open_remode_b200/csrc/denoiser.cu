// denoiser.cu -- weighted TV-L1 (Chambolle-Pock) smoothing of the depth map.
//
// Mathematics: src/depthmap_denoiser.cu:46-59 (weights), :62-118 (one
// primal-dual step), :124-141 (L = sqrt(8), tau = 0.02, sigma = 1/(L^2 tau),
// theta = 0.5), :226-229 (large sigma^2 = range^2 / 72).
//
// Organisation (ours): the iterate (u, u_head, p.x, p.y) is one float4 per
// pixel, ping-ponged between two buffers, so an iteration is a pure function
// of the previous one -- a deterministic Jacobi sweep.  The reference updates
// p/u/u_head in place with only an intra-block barrier and is therefore
// racy across 16x16 tile seams; inside a tile it is the
// same Jacobi sweep.  Each CTA stages its tile plus a one-pixel halo of the
// old iterate in shared memory, computes the new dual for the tile and its
// west/north halo there, and then the primal update from shared memory.
#include "denoiser.cuh"

namespace rmdb
{

namespace
{
constexpr int TILE_W = 32;
constexpr int TILE_H = 8;
constexpr int EXT_W = TILE_W + 2;  // x0-1 .. x0+TILE_W
constexpr int EXT_H = TILE_H + 2;  // y0-1 .. y0+TILE_H
}

// computeWeightsKernel (:46-59) + "u_ = mu; u_head_ = u_; p_.zero()" (:215-217)
template<bool FROM_SEEDS>
__global__ void __launch_bounds__(256) denoise_setup_kernel(const DenoiseSetupParams P)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x >= P.width || y >= P.height)
    return;
  float mu, sigma_sq, a, b;
  if(FROM_SEEDS)
  {
    const float4 s = P.seed[(size_t)y * P.seed_stride + x];
    mu = s.x; sigma_sq = s.y; a = s.z; b = s.w;
  }
  else
  {
    mu = P.mu[(size_t)y * P.mu_stride + x];
    sigma_sq = P.sigma_sq[(size_t)y * P.sigma_sq_stride + x];
    a = P.a[(size_t)y * P.a_stride + x];
    b = P.b[(size_t)y * P.b_stride + x];
  }
  const float E_pi = a / (a + b);
  const float g = fmaxf((E_pi * sigma_sq + (1.0f - E_pi) * P.large_sigma_sq) / P.large_sigma_sq, 1.0f);
  P.gmu[(size_t)y * P.stride + x] = make_float2(g, mu);
  P.state[(size_t)y * P.stride + x] = make_float4(mu, mu, 0.0f, 0.0f);
}

// updateTVL1PrimalDualKernel (:62-118), Jacobi form.
__global__ void __launch_bounds__(TILE_W * TILE_H) denoise_step_kernel(const DenoiseStepParams P)
{
  __shared__ float4 s_old[EXT_H][EXT_W];          // (u, u_head, p.x, p.y) of iteration n
  __shared__ float s_g[EXT_H][EXT_W];
  __shared__ float2 s_p[TILE_H + 1][TILE_W + 1];  // p of iteration n+1, [0][*] north, [*][0] west

  const int x0 = blockIdx.x * TILE_W;
  const int y0 = blockIdx.y * TILE_H;
  const int tid = threadIdx.y * TILE_W + threadIdx.x;

  // stage tile + halo; clamped coordinates implement the reference's
  // min(width-1, x+1) / max(0, x-1) neighbour addressing (:79-80, :87-88)
  for(int k = tid; k < EXT_W * EXT_H; k += TILE_W * TILE_H)
  {
    const int ey = k / EXT_W, ex = k - ey * EXT_W;
    const int gx = min(max(x0 - 1 + ex, 0), P.width - 1);
    const int gy = min(max(y0 - 1 + ey, 0), P.height - 1);
    s_old[ey][ex] = P.in[(size_t)gy * P.stride + gx];
    s_g[ey][ex] = P.gmu[(size_t)gy * P.stride + gx].x;
  }
  __syncthreads();

  // dual ascent + projection for the tile and its west / north halo (:70-83)
  for(int k = tid; k < (TILE_W + 1) * (TILE_H + 1); k += TILE_W * TILE_H)
  {
    const int py = k / (TILE_W + 1), px = k - py * (TILE_W + 1);
    // element (px, py) of s_p is pixel (x0-1+px, y0-1+py) = s_old[py][px]
    const float4 c = s_old[py][px];
    const float g = s_g[py][px];
    // east / south neighbours; at the image edge the clamp makes them the
    // pixel itself, which staging already resolved except when the pixel
    // itself is the last column/row *inside* the tile:
    const int gx = x0 - 1 + px, gy = y0 - 1 + py;
    const float uh_e = (gx >= P.width - 1) ? c.y : s_old[py][px + 1].y;
    const float uh_s = (gy >= P.height - 1) ? c.y : s_old[py + 1][px].y;
    const float grad_x = uh_e - c.x;
    const float grad_y = uh_s - c.x;
    const float tx = g * grad_x * P.sigma + c.z;
    const float ty = g * grad_y * P.sigma + c.w;
    const float len = sqrtf(tx * tx + ty * ty);
    const float d = fmaxf(1.0f, len);
    s_p[py][px] = make_float2(tx / d, ty / d);
  }
  __syncthreads();

  // divergence, primal shrink towards the noisy depth, over-relaxation (:86-115)
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if(x < P.width && y < P.height)
  {
    const float2 own = s_p[threadIdx.y + 1][threadIdx.x + 1];
    float2 cp = own;
    float2 wp = (x == 0) ? own : s_p[threadIdx.y + 1][threadIdx.x];
    float2 np = (y == 0) ? own : s_p[threadIdx.y][threadIdx.x + 1];
    if(x == 0) wp.x = 0.0f;
    else if(x >= P.width - 1) cp.x = 0.0f;
    if(y == 0) np.y = 0.0f;
    else if(y >= P.height - 1) cp.y = 0.0f;
    const float divergence = cp.x - wp.x + cp.y - np.y;

    const float4 c = s_old[threadIdx.y + 1][threadIdx.x + 1];
    const float g = s_g[threadIdx.y + 1][threadIdx.x + 1];
    const float noisy = P.gmu[(size_t)y * P.stride + x].y;
    const float old_u = c.x;
    const float temp_u = old_u + P.tau * g * divergence;
    float new_u;
    if((temp_u - noisy) > (P.tau * P.lambda))
      new_u = temp_u - P.tau * P.lambda;
    else if((temp_u - noisy) < (-P.tau * P.lambda))
      new_u = temp_u + P.tau * P.lambda;
    else
      new_u = noisy;
    const float new_uh = new_u + P.theta * (new_u - old_u);
    P.out[(size_t)y * P.stride + x] = make_float4(new_u, new_uh, own.x, own.y);
  }
}

static inline dim3 grid_for(int width, int height, dim3 block)
{
  return dim3((width + block.x - 1) / block.x, (height + block.y - 1) / block.y);
}

cudaError_t launch_denoise_setup(const DenoiseSetupParams &P, bool from_seeds, cudaStream_t stream)
{
  const dim3 block(32, 8);
  const dim3 grid = grid_for(P.width, P.height, block);
  if(from_seeds)
    denoise_setup_kernel<true><<<grid, block, 0, stream>>>(P);
  else
    denoise_setup_kernel<false><<<grid, block, 0, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_denoise_step(const DenoiseStepParams &P, cudaStream_t stream)
{
  const dim3 block(TILE_W, TILE_H);
  denoise_step_kernel<<<grid_for(P.width, P.height, block), block, 0, stream>>>(P);
  return cudaGetLastError();
}

} // namespace rmdb
