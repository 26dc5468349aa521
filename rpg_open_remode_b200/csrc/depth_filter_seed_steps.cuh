// depth_filter_seed_steps.cuh -- the per-seed steps of the fused depth filter, shared by every kernel that
// organises the work differently (CTA per tile, warp per sparse tile, warp per group of listed seeds):
// classification, candidate counting with the reference's own float accumulation, the exact accepted
// candidate range, and the final match -> triangulation -> Bayesian update.  Identical code, identical results.
#pragma once

#include <limits.h>

#include "depth_filter_math.cuh"
#include "staged_maps.cuh"

namespace rmdb
{

using namespace staged;

struct SearchRec  // one active seed's search, 32 bytes
{
  float mean_x, mean_y, dir_x, dir_y, half_len, sum_templ, denom;
  int n;
};

// float -> unsigned with the same ordering (for atomicMax on NCC scores)
__device__ __forceinline__ unsigned int orderable(float f)
{
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- frame chaining: release / acquire on global flags, bounded waits
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_release(unsigned int *p, const unsigned int v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Waits until *p >= want.  Bounded (~0.2 s): whatever is waited for has already been picked by a resident
// CTA, so the bound is never reached in a correct run; if it is, the launch reports it instead of hanging.
__device__ __forceinline__ bool wait_at_least(const unsigned int *p, const unsigned int want)
{
  for(unsigned int it = 0; it < (1u << 22); ++it)
  {
    if(ld_acquire(p) >= want)
      return true;
    __nanosleep(32);
  }
  return false;
}


// ------------------------------------------------------------ per-seed steps shared by the CTA and warp paths

constexpr unsigned long long K_NO_MATCH = 0x407FFFFF00000000ull;   // orderable(-1.0f) << 32: "no candidate scored"

// src/seed_check.cu:37-66 for one pixel; `prev` is the state the previous frame left (absorbing states are final
// when P.trust_conv says the map agrees with the parameters).
template<int PS>
__device__ __forceinline__ int classify_pixel(const FilterParams &P, const int x, const int y, const int prev,
                                              const float4 seed, bool &active)
{
  active = false;
  if(P.trust_conv && (prev == RMD_BORDER || prev == RMD_CONVERGED || prev == RMD_DIVERGED))
    return prev;
  if((x > P.width - PS - 1) || (y > P.height - PS - 1) || (x < PS) || (y < PS))
    return RMD_BORDER;
  const int state = classify_seed(P, seed);
  active = (state == RMD_UPDATE);
  return state;
}

// Phase 4 for one seed that was searched: NO_MATCH (b += 1, seed_update.cu:113-117) or triangulation + Bayesian
// update from the best candidate (key = (orderable(ncc) << 32) | ~index, K_NO_MATCH if nothing was scored).
// `seed` is updated in place (and stored); returns the seed's new state (stored when it differs from `prev`).
template<int PS>
__device__ __forceinline__ int apply_match(const FilterParams &P, const int x, const int y, const EpiSegment &seg,
                                            const int n_cand, const unsigned long long key, const float *ckpt,
                                            float4 &seed, float4 *seed_ptr, const int prev, int *conv_ptr)
{
  int state = RMD_UPDATE;
  if(key == K_NO_MATCH || !(n_cand > 0))
  {
    state = RMD_NO_MATCH;
  }
  else
  {
    const unsigned int hi = (unsigned int)(key >> 32);
    const float best_ncc = __uint_as_float((hi & 0x80000000u) ? (hi & 0x7fffffffu) : ~hi);
    if(best_ncc < RMD_NCC_ACCEPT)
    {
      state = RMD_NO_MATCH;
    }
    else
    {
      const int best_idx = (int)(0xffffffffu - (unsigned int)(key & 0xffffffffull));
      const float l = candidate_l(ckpt, best_idx);
      const float2 best_px = candidate_px(seg.mean.x, seg.mean.y, seg.dir.x, seg.dir.y, l);
      if(P.matches)
        P.matches[(size_t)y * P.match_stride + x] = best_px;
      if(bayes_update(P, x, y, best_px, seed))
        *seed_ptr = seed;
    }
  }
  if(state == RMD_NO_MATCH)
  {
    seed.w += 1.0f;  // seed_update.cu:113-117
    *seed_ptr = seed;
  }
  if(state != prev)
    *conv_ptr = state;
  return state;
}

} // namespace rmdb
