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

__device__ __forceinline__ int to_int_clamped(float v)
{
  return (int)fminf(fmaxf(v, -1.0e6f), 1.0e6f);  // NaN -> -1e6 (fmaxf drops NaN)
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

// The candidate positions follow the reference's own float accumulation of l (epipolar_match.cu:88).  One cheap
// pass counts them and records l every 16th candidate in ckpt[] (work items and the final match restart from a
// checkpoint, bit-identically).  l only grows, so a block of 16 additions needs one end test, not 16: the same
// additions in the same order as the reference's loop, a fifth of the instructions of the naive transcription.
__device__ __forceinline__ int count_candidates(const float half_len, float *ckpt)
{
  int k = 0;
  float l = -half_len;
  while(l <= half_len && k < L_CHECKPOINT_STEP * L_CHECKPOINTS)
  {
    ckpt[k / L_CHECKPOINT_STEP] = l;
    float l_blk = l;
#pragma unroll
    for(int t = 0; t < L_CHECKPOINT_STEP; ++t) l_blk += RMD_EPIPOLAR_STEP;
    if(l_blk <= half_len)
    {
      l = l_blk;              // candidates k .. k+16 all exist
      k += L_CHECKPOINT_STEP;
      continue;
    }
    int t = 1;                // the last candidate is k + t - 1, 1 <= t <= 16
    for(l += RMD_EPIPOLAR_STEP; t < L_CHECKPOINT_STEP && l <= half_len; l += RMD_EPIPOLAR_STEP) ++t;
    k += t;
    break;
  }
  return k;
}

// l of candidate k: restart from the checkpoint, at most 15 of the reference's additions.  Kept a rolled loop:
// unrolled at each of its ~20 inlined call sites it made up 1069 of the staged kernel's 7872 instructions and
// was no faster (profiles/r02_tune_probe.txt).
__device__ __forceinline__ float candidate_l(const float *ckpt, const int k)
{
  float l = ckpt[k / L_CHECKPOINT_STEP];
#pragma unroll 1
  for(int t = 0; t < (k & (L_CHECKPOINT_STEP - 1)); ++t) l += RMD_EPIPOLAR_STEP;
  return l;
}

// The candidates that pass the image-bounds test (epipolar_match.cu:91-97) form one contiguous index range
// [k_lo, k_hi] -- the segment is a straight line, the accepted region convex and float rounding monotone -- so
// everything outside it is skipped wholesale and seeds whose projection left the image cost no work.  The range
// is estimated in closed form and then fixed EXACTLY by testing the real candidates around the estimate.
// k_hi < 0: none.
template<int PS>
__device__ __forceinline__ void accepted_range(const FilterParams &P, const EpiSegment &seg, const int n_cand,
                                               const float *ckpt, int &k_lo, int &k_hi)
{
  k_lo = INT_MAX; k_hi = -1;
  if(n_cand <= 0)
    return;
  auto accepted = [&](int k) -> bool
  {
    const float l = candidate_l(ckpt, k);
    const float2 px = candidate_px(seg.mean.x, seg.mean.y, seg.dir.x, seg.dir.y, l);
    return !candidate_rejected<PS>(px, P.width, P.height);
  };
  // l-interval in which P <= mean + l*dir < size - P holds, per axis
  float la = -1.0e30f, lb = 1.0e30f;
  bool none = false, exact_scan = false;
  {
    const float lo_x = (float)PS, hi_x = (float)(P.width - PS), lo_y = (float)PS, hi_y = (float)(P.height - PS);
    const float m[2] = {seg.mean.x, seg.mean.y}, d[2] = {seg.dir.x, seg.dir.y};
    const float lo[2] = {lo_x, lo_y}, hi[2] = {hi_x, hi_y};
#pragma unroll
    for(int ax = 0; ax < 2; ++ax)
    {
      if(!(fabsf(m[ax]) < 1.0e7f) || !(fabsf(d[ax]) <= 2.0f))
        exact_scan = true;                       // NaN / inf: no shortcut
      else if(fabsf(d[ax]) < 1.0e-6f)
      {
        // the segment does not move along this axis: inside, outside, or too close to call
        if((fabsf(m[ax] - lo[ax]) <= 0.5f) || (fabsf(m[ax] - hi[ax]) <= 0.5f))
          exact_scan = true;
        else if((m[ax] < lo[ax]) || (m[ax] >= hi[ax]))
          none = true;
      }
      else
      {
        const float t0 = (lo[ax] - m[ax]) / d[ax], t1 = (hi[ax] - m[ax]) / d[ax];
        la = fmaxf(la, fminf(t0, t1));
        lb = fminf(lb, fmaxf(t0, t1));
      }
    }
  }
  if(exact_scan)
  {
    for(int k = 0; k < n_cand; ++k)
      if(accepted(k)) { k_lo = min(k_lo, k); k_hi = k; }
  }
  else if(!none)
  {
    // estimated index range, widened by 2 candidates on both sides
    const float fa = (la + seg.half_len) / RMD_EPIPOLAR_STEP, fb = (lb + seg.half_len) / RMD_EPIPOLAR_STEP;
    const int a = max(0, to_int_clamped(ceilf(fa)) - 2), b = min(n_cand - 1, to_int_clamped(floorf(fb)) + 2);
    if(a <= b)
    {
      int first = -1, last = -1;
      for(int k = a; k <= min(a + 4, b); ++k)
        if(accepted(k)) { first = k; break; }
      for(int k = b; k >= max(b - 4, a); --k)
        if(accepted(k)) { last = k; break; }
      if(first >= 0 && last >= 0)
      {
        // the estimate must have bracketed the true ends; if an end sits on
        // the widened border (and is not the segment's end) scan further
        while(first > 0 && first == a && accepted(first - 1)) { --first; }
        while(last < n_cand - 1 && last == b && accepted(last + 1)) { ++last; }
        k_lo = first; k_hi = last;
      }
      else if(b - a > 4)
      {
        // an end was not found next to its estimate: be exact over the whole window
        for(int k = a; k <= b; ++k)
          if(accepted(k)) { k_lo = min(k_lo, k); k_hi = k; }
      }
    }
  }
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
