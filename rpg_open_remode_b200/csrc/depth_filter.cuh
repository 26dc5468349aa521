// depth_filter.cuh -- launch interface of the seed-matrix kernels.
#pragma once

#include "rmd_common.cuh"

namespace rmdb
{

struct InitParams
{
  int width, height;
  const float *ref;
  int ref_stride;
  float4 *seed;
  int seed_stride;
  float2 *templ;
  int templ_stride;
  int *conv;
  int conv_stride;
  float avg_depth, sigma_sq_max;
};

cudaError_t launch_seed_init(const InitParams &P, int patch_side, cudaStream_t stream);

// One thread per pixel, global loads.
cudaError_t launch_depth_filter_direct(const FilterParams &P, int patch_side, cudaStream_t stream);

// TMA-staged shared-memory variant (depth_filter_staged.cu).  `maps` is the
// host-side descriptor set built by StagedMaps::encode for this frame.
struct StagedMaps;
// One launch for n keyframes (n = 1: the single-keyframe instantiation).  Persistent grid: one CTA per
// resident slot, each pulling entries of the keyframes' work lists through `cursor` (2 zero-initialised
// uints owned by the caller; the kernel leaves them at zero).
// chain = 0: P[0..n) are n keyframes updated by the same frame.  chain = 1: P[0..n) are n CONSECUTIVE FRAMES of
// one keyframe; a tile moves on to frame k+1 as soon as its own frame k is final (FilterParams::tile_done), so
// frames overlap on the GPU and the launch gap between frames disappears.  `cursor`: STAGED_CURSOR_WORDS
// zero-initialised uints (staged_maps.cuh).
cudaError_t launch_depth_filter_staged(const FilterParams *const *P, const StagedMaps *const *maps, int n, int chain,
                                       unsigned int *cursor, int patch_side, cudaStream_t stream);
// Resident CTAs of the staged kernel on the current device (SMs x occupancy), 0 on error.
int staged_cta_slots(int patch_side, int ctas_per_sm);

// Seed-major variant for the steady state of a keyframe (depth_filter_seeds.cu): the seeds that are still
// updated are kept as a compact list; one launch walks every listed seed through up to SEED_FRAMES_MAX
// consecutive frames (p[f] = parameters of frame f; the work-list members of FilterParams are unused).
// ctl: {listed seeds (ping), listed seeds (pong), group cursor, CONVERGED seeds of the keyframe, CTAs done}.
constexpr int SEED_FRAMES_MAX = 16;
struct alignas(16) SeedModeBatch
{
  FilterParams p[SEED_FRAMES_MAX];
  const unsigned int *list_cur;    // x | y << 16 of the listed seeds, ctl[cur] of them
  unsigned int *list_next;         // survivors are appended here, ctl[cur ^ 1] counts them (zero at launch)
  unsigned int *ctl;
  int cur;
  int n_frames;
};
cudaError_t launch_seed_list_build(const int *conv, int conv_stride, int width, int height, unsigned int *list,
                                   unsigned int *ctl, cudaStream_t stream);
// max_listed: an upper bound of ctl[cur] known to the host (sizes the persistent grid only)
cudaError_t launch_depth_filter_seeds(const SeedModeBatch &B, int max_listed, int patch_side, cudaStream_t stream);

// dst[i] = value for i < n (64-bit pattern fill)
cudaError_t launch_fill_u64(unsigned long long *dst, size_t n, unsigned long long value, cudaStream_t stream);

// light[i] = tile i (share 0 of 1) for i < n_tiles; counts = {0, n_tiles, 0, 0}
cudaError_t launch_worklist_init(unsigned int *light, unsigned int *counts, int n_tiles, cudaStream_t stream);

cudaError_t launch_u8_to_float(const uint8_t *src, int src_stride, float *dst, int dst_stride,
                               int width, int height, cudaStream_t stream);
cudaError_t launch_export_plane(const float *src, int src_stride_floats, int comps, int comp,
                                float *dst, int dst_stride, int width, int height,
                                cudaStream_t stream);
cudaError_t launch_import_plane(const float *src, int src_stride, float *dst, int dst_stride_floats,
                                int comps, int comp, int width, int height, cudaStream_t stream);

} // namespace rmdb
