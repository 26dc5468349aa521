// depth_filter.cu -- kernels of the seed matrix: initialisation and the fused
// per-frame update (convergence check + epipolar NCC search + triangulation +
// Bayesian update), "direct" variant: one thread per pixel, global loads.
// The staged variant (TMA -> shared memory, balanced work list) lives in
// depth_filter_staged.cu and shares depth_filter_math.cuh, so both variants
// execute the same arithmetic per candidate.
//
// Replaces the reference's four launches per frame with two host syncs
// (src/seed_matrix.cu:139-155: seedCheckKernel, seedEpipolarMatchKernel,
// seedUpdateKernel) by one launch and no sync.
#include "depth_filter.cuh"
#include "depth_filter_math.cuh"

namespace rmdb
{

// ------------------------------------------------------------------ init

// seedInitKernel, src/seed_init.cu:28-61: template statistics of the PSxPS
// patch around every pixel (clamped addressing at the image edge, like the
// reference's texture fetches) and the prior of every seed.  Additionally
// writes the initial convergence map (BORDER ring / UPDATE interior) so that
// the map is defined before the first update (the reference leaves it
// uninitialised until then).
template<int PS>
__global__ void __launch_bounds__(256) seed_init_kernel(const InitParams P)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x >= P.width || y >= P.height)
    return;

  float sum_templ = 0.0f, sum_templ_sq = 0.0f;
#pragma unroll
  for(int py = 0; py < PS; ++py)
  {
    const int yy = min(max(y - PS / 2 + py, 0), P.height - 1);
#pragma unroll
    for(int px = 0; px < PS; ++px)
    {
      const int xx = min(max(x - PS / 2 + px, 0), P.width - 1);
      const float t = __ldg(P.ref + (size_t)yy * P.ref_stride + xx);
      sum_templ += t;
      sum_templ_sq += t * t;
    }
  }
  const float denom = (float)((double)(PS * PS) * (double)sum_templ_sq -
                              (double)sum_templ * (double)sum_templ);
  P.templ[(size_t)y * P.templ_stride + x] = make_float2(sum_templ, denom);
  P.seed[(size_t)y * P.seed_stride + x] = make_float4(P.avg_depth, P.sigma_sq_max, 10.0f, 10.0f);
  const bool border = (x > P.width - PS - 1) || (y > P.height - PS - 1) || (x < PS) || (y < PS);
  P.conv[(size_t)y * P.conv_stride + x] = border ? RMD_BORDER : RMD_UPDATE;
}

// --------------------------------------------------- fused update, direct

template<int PS>
__global__ void __launch_bounds__(256) depth_filter_direct_kernel(const __grid_constant__ FilterParams P)
{
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;

  if(blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0)
    *P.converged_next = 0u;  // nobody else touches the next frame's counter now

  bool converged = false;
  if(x < P.width && y < P.height)
  {
    int *conv_ptr = P.conv + (size_t)y * P.conv_stride + x;
    const int prev = *conv_ptr;
    int state;
    if(P.trust_conv && (prev == RMD_BORDER || prev == RMD_CONVERGED || prev == RMD_DIVERGED))
    {
      // BORDER never changes; CONVERGED / DIVERGED seeds are never written
      // again (seed_update.cu:54-56), so seed_check.cu would re-derive the
      // same state from the same numbers: skip the 16-byte seed read.
      state = prev;
    }
    else if((x > P.width - PS - 1) || (y > P.height - PS - 1) || (x < PS) || (y < PS))
    {
      state = RMD_BORDER;  // seed_check.cu:37-42
    }
    else
    {
      float4 *seed_ptr = P.seed + (size_t)y * P.seed_stride + x;
      float4 seed = *seed_ptr;
      state = classify_seed(P, seed);
      if(state == RMD_UPDATE)
      {
        // template (reference patch) and its statistics
        float templ[PS * PS];
#pragma unroll
        for(int py = 0; py < PS; ++py)
#pragma unroll
          for(int px = 0; px < PS; ++px)
            templ[py * PS + px] =
                __ldg(P.ref + (size_t)(y - PS / 2 + py) * P.ref_stride + (x - PS / 2 + px));
        const float2 stats = __ldg(P.templ + (size_t)y * P.templ_stride + x);

        const EpiSegment seg = epipolar_segment(P, x, y, seed.x, seed.y);
        float best_ncc = -1.0f;
        float2 best_px = make_float2(0.0f, 0.0f);
        for(float l = -seg.half_len; l <= seg.half_len; l += RMD_EPIPOLAR_STEP)
        {
          const float2 px = candidate_px(seg.mean.x, seg.mean.y, seg.dir.x, seg.dir.y, l);
          if(candidate_rejected<PS>(px, P.width, P.height))
            continue;
          const TapFrame frame = tap_frame<PS>(px, P.tex_quant);
          const GlobalTaps taps(P.curr, P.curr_stride, frame);
          const float ncc = ncc_score<PS>(taps, frame, templ, stats.x, stats.y);
          if(ncc > best_ncc)
          {
            best_px = px;
            best_ncc = ncc;
          }
        }
        if(best_ncc < RMD_NCC_ACCEPT)
        {
          state = RMD_NO_MATCH;  // epipolar_match.cu:131-134
          seed.w += 1.0f;        // seed_update.cu:113-117
          *seed_ptr = seed;
        }
        else
        {
          if(P.matches)
            P.matches[(size_t)y * P.match_stride + x] = best_px;
          if(bayes_update(P, x, y, best_px, seed))
            *seed_ptr = seed;
        }
      }
    }
    if(state != prev)
      *conv_ptr = state;
    converged = (state == RMD_CONVERGED);
  }

  const unsigned int ballot = __ballot_sync(0xffffffffu, converged);
  if(threadIdx.x == 0 && ballot)
    atomicAdd(P.converged_now, (unsigned int)__popc(ballot));
}

// ----------------------------------------------------------- small kernels

// u8 -> float * (1/255.f): the GPU side of rmd::Depthmap::inputImage
// (src/depthmap.cpp:105, cv::Mat::convertTo(CV_32F, 1.0f/255.0f)).
__global__ void __launch_bounds__(256) u8_to_float_kernel(
    const uint8_t *__restrict__ src, int src_stride, float *__restrict__ dst, int dst_stride,
    int width, int height)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x < width && y < height)
    dst[(size_t)y * dst_stride + x] = (float)src[(size_t)y * src_stride + x] * (1.0f / 255.0f);
}

// seed / template planes <-> planar images (getMu() etc. and the test hooks)
__global__ void __launch_bounds__(256) export_plane_kernel(
    const float *__restrict__ src, int src_stride_floats, int comps, int comp,
    float *__restrict__ dst, int dst_stride, int width, int height)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x < width && y < height)
    dst[(size_t)y * dst_stride + x] = src[(size_t)y * src_stride_floats + (size_t)x * comps + comp];
}

__global__ void __launch_bounds__(256) import_plane_kernel(
    const float *__restrict__ src, int src_stride, float *__restrict__ dst,
    int dst_stride_floats, int comps, int comp, int width, int height)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x < width && y < height)
    dst[(size_t)y * dst_stride_floats + (size_t)x * comps + comp] = src[(size_t)y * src_stride + x];
}

__global__ void __launch_bounds__(256) fill_u64_kernel(unsigned long long *dst, size_t n, unsigned long long value)
{
  for(size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = value;
}

// ---------------------------------------------------------------- launchers

static inline dim3 grid_for(int width, int height, dim3 block)
{
  return dim3((width + block.x - 1) / block.x, (height + block.y - 1) / block.y);
}

cudaError_t launch_seed_init(const InitParams &P, int patch_side, cudaStream_t stream)
{
  const dim3 block(32, 8);
  const dim3 grid = grid_for(P.width, P.height, block);
  if(patch_side == 5)
    seed_init_kernel<5><<<grid, block, 0, stream>>>(P);
  else if(patch_side == 7)
    seed_init_kernel<7><<<grid, block, 0, stream>>>(P);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_depth_filter_direct(const FilterParams &P, int patch_side, cudaStream_t stream)
{
  const dim3 block(32, 8);
  const dim3 grid = grid_for(P.width, P.height, block);
  if(patch_side == 5)
    depth_filter_direct_kernel<5><<<grid, block, 0, stream>>>(P);
  else if(patch_side == 7)
    depth_filter_direct_kernel<7><<<grid, block, 0, stream>>>(P);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_fill_u64(unsigned long long *dst, size_t n, unsigned long long value, cudaStream_t stream)
{
  const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  fill_u64_kernel<<<blocks > 0 ? blocks : 1, 256, 0, stream>>>(dst, n, value);
  return cudaGetLastError();
}

// Work list of the staged kernel's first frame: every tile once, image order.
__global__ void worklist_init_kernel(unsigned int *light, unsigned int *counts, int n_tiles)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n_tiles)
    light[i] = (unsigned int)i | (1u << 26);   // share 0 of 1
  if(i == 0)
  {
    counts[0] = 0u; counts[1] = (unsigned int)n_tiles; counts[2] = 0u; counts[3] = 0u;
    counts[4] = (unsigned int)n_tiles; counts[5] = 0u;   // every tile listed once; see FilterParams::counts_cur
    counts[6] = 0u; counts[7] = 0u;
  }
}

cudaError_t launch_worklist_init(unsigned int *light, unsigned int *counts, int n_tiles, cudaStream_t stream)
{
  worklist_init_kernel<<<(n_tiles + 255) / 256, 256, 0, stream>>>(light, counts, n_tiles);
  return cudaGetLastError();
}

cudaError_t launch_u8_to_float(const uint8_t *src, int src_stride, float *dst, int dst_stride,
                               int width, int height, cudaStream_t stream)
{
  const dim3 block(32, 8);
  u8_to_float_kernel<<<grid_for(width, height, block), block, 0, stream>>>(
      src, src_stride, dst, dst_stride, width, height);
  return cudaGetLastError();
}

cudaError_t launch_export_plane(const float *src, int src_stride_floats, int comps, int comp,
                                float *dst, int dst_stride, int width, int height,
                                cudaStream_t stream)
{
  const dim3 block(32, 8);
  export_plane_kernel<<<grid_for(width, height, block), block, 0, stream>>>(
      src, src_stride_floats, comps, comp, dst, dst_stride, width, height);
  return cudaGetLastError();
}

cudaError_t launch_import_plane(const float *src, int src_stride, float *dst, int dst_stride_floats,
                                int comps, int comp, int width, int height, cudaStream_t stream)
{
  const dim3 block(32, 8);
  import_plane_kernel<<<grid_for(width, height, block), block, 0, stream>>>(
      src, src_stride, dst, dst_stride_floats, comps, comp, width, height);
  return cudaGetLastError();
}

} // namespace rmdb
