// candidate_range.cuh -- where the candidates of an epipolar search are: the reference's float accumulation of l,
// checkpoints to restart it, and the exact index range of the candidates that pass the image-bounds test.  Pure
// arithmetic on a segment, no memory traffic: compiled for the device by nvcc (unchanged code: these functions were
// moved here from depth_filter_math.cuh / depth_filter_seed_steps.cuh) and for the HOST by plain g++, where
// tests/cpp/candidate_range_test.cpp checks them against the naive transcription of the reference's loop
// (src/epipolar_match.cu:85-97) on a CPU-only box.
#pragma once

#include <limits.h>
#include <math.h>

#include "rmd_common.cuh"
#include "staged_maps.cuh"

#if defined(__CUDACC__)
#define RMD_HD __host__ __device__ __forceinline__
#else
#define RMD_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define RMD_MIN(a, b) min(a, b)
#define RMD_MAX(a, b) max(a, b)
#else
#define RMD_MIN(a, b) ((a) < (b) ? (a) : (b))
#define RMD_MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

namespace rmdb
{

#define RMD_EPIPOLAR_STEP 0.7f          // src/epipolar_match.cu:88

// a * b + c with one rounding: the intrinsic on the device, C99 fmaf on the host
RMD_HD float fma_rn(const float a, const float b, const float c)
{
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}

// The epipolar search segment of one seed, src/epipolar_match.cu:60-75.
struct EpiSegment
{
  float2 mean;      // projection of the depth estimate
  float2 dir;       // unit direction of the segment
  float half_len;   // half of min(segment length, 100 px)
};

// True when the candidate patch centre is outside the searchable interior,
// src/epipolar_match.cu:91-97 (NaN coordinates pass, as there).
template<int PS>
RMD_HD bool candidate_rejected(const float2 px, const int width, const int height)
{
  return (px.x >= (float)(width - PS)) || (px.y >= (float)(height - PS)) ||
         (px.x < (float)PS) || (px.y < (float)PS);
}

// Candidate l of a segment: px_mean + l * epi_dir (src/epipolar_match.cu:90), one fused multiply-add per axis.
RMD_HD float2 candidate_px(const float mean_x, const float mean_y, const float dir_x,
                                               const float dir_y, const float l)
{
  return make_float2(fma_rn(l, dir_x, mean_x), fma_rn(l, dir_y, mean_y));
}

RMD_HD int to_int_clamped(float v)
{
  return (int)fminf(fmaxf(v, -1.0e6f), 1.0e6f);  // NaN -> -1e6 (fmaxf drops NaN)
}


// The candidate positions follow the reference's own float accumulation of l (epipolar_match.cu:88).  One cheap
// pass counts them and records l every 16th candidate in ckpt[] (work items and the final match restart from a
// checkpoint, bit-identically).  l only grows, so a block of 16 additions needs one end test, not 16: the same
// additions in the same order as the reference's loop, a fifth of the instructions of the naive transcription.
RMD_HD int count_candidates(const float half_len, float *ckpt)
{
  int k = 0;
  float l = -half_len;
  while(l <= half_len && k < staged::L_CHECKPOINT_STEP * staged::L_CHECKPOINTS)
  {
    ckpt[k / staged::L_CHECKPOINT_STEP] = l;
    float l_blk = l;
#pragma unroll
    for(int t = 0; t < staged::L_CHECKPOINT_STEP; ++t) l_blk += RMD_EPIPOLAR_STEP;
    if(l_blk <= half_len)
    {
      l = l_blk;              // candidates k .. k+16 all exist
      k += staged::L_CHECKPOINT_STEP;
      continue;
    }
    int t = 1;                // the last candidate is k + t - 1, 1 <= t <= 16
    for(l += RMD_EPIPOLAR_STEP; t < staged::L_CHECKPOINT_STEP && l <= half_len; l += RMD_EPIPOLAR_STEP) ++t;
    k += t;
    break;
  }
  return k;
}

// l of candidate k: restart from the checkpoint, at most 15 of the reference's additions.  Kept a rolled loop:
// unrolled at each of its ~20 inlined call sites it made up 1069 of the staged kernel's 7872 instructions and
// was no faster (profiles/r02_tune_probe.txt).
RMD_HD float candidate_l(const float *ckpt, const int k)
{
  float l = ckpt[k / staged::L_CHECKPOINT_STEP];
#pragma unroll 1
  for(int t = 0; t < (k & (staged::L_CHECKPOINT_STEP - 1)); ++t) l += RMD_EPIPOLAR_STEP;
  return l;
}

// The candidates that pass the image-bounds test (epipolar_match.cu:91-97) form one contiguous index range
// [k_lo, k_hi] -- the segment is a straight line, the accepted region convex and float rounding monotone -- so
// everything outside it is skipped wholesale and seeds whose projection left the image cost no work.  The range
// is estimated in closed form and then fixed EXACTLY by testing the real candidates around the estimate.
// k_hi < 0: none.
template<int PS>
RMD_HD void accepted_range(const FilterParams &P, const EpiSegment &seg, const int n_cand,
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
      if(accepted(k)) { k_lo = RMD_MIN(k_lo, k); k_hi = k; }
  }
  else if(!none)
  {
    // estimated index range, widened by 2 candidates on both sides
    const float fa = (la + seg.half_len) / RMD_EPIPOLAR_STEP, fb = (lb + seg.half_len) / RMD_EPIPOLAR_STEP;
    const int a = RMD_MAX(0, to_int_clamped(ceilf(fa)) - 2), b = RMD_MIN(n_cand - 1, to_int_clamped(floorf(fb)) + 2);
    if(a <= b)
    {
      int first = -1, last = -1;
      for(int k = a; k <= RMD_MIN(a + 4, b); ++k)
        if(accepted(k)) { first = k; break; }
      for(int k = b; k >= RMD_MAX(b - 4, a); --k)
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
          if(accepted(k)) { k_lo = RMD_MIN(k_lo, k); k_hi = k; }
      }
    }
  }
}

} // namespace rmdb
