// depth_filter_math.cuh -- per-pixel arithmetic of the fused depth filter.
//
// What is computed follows the reference (cited per function); how it is
// organised does not: one thread-level "seed" record (mu, sigma_sq, a, b) in
// registers, pose and camera in constant kernel parameters, bilinear taps
// from a 6x6 (P+1 x P+1) texel neighbourhood filtered separably in registers
// instead of 25 texture fetches per candidate.
#pragma once

#include <float.h>
#include <math_constants.h>

#include "rmd_common.cuh"

namespace rmdb
{

#define RMD_MAX_EPIPOLAR_EXTENT 100.0f  // RMD_MAX_EXTENT_EPIPOLAR_SEARCH, CMakeLists.txt:53
#define RMD_EPIPOLAR_STEP 0.7f          // src/epipolar_match.cu:88
#define RMD_NCC_ACCEPT 0.5f             // src/epipolar_match.cu:131

__device__ __forceinline__ float dot3(const float3 a, const float3 b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z;
}

__device__ __forceinline__ float3 scaled(const float3 v, const float s)
{
  return make_float3(v.x * s, v.y * s, v.z * s);
}

// normalize(cam.cam2world(px)): pinhole_camera.cuh:40-46, helper_math.h:1309-1313
__device__ __forceinline__ float3 unit_bearing(const Camera &cam, const float u, const float v)
{
  const float3 p = make_float3((u - cam.cx) / cam.fx, (v - cam.cy) / cam.fy, 1.0f);
  return scaled(p, rsqrtf(dot3(p, p)));
}

// se3.cuh:111-117
__device__ __forceinline__ float3 rotate(const Pose &T, const float3 p)
{
  return make_float3(T.m[0] * p.x + T.m[1] * p.y + T.m[2] * p.z,
                     T.m[4] * p.x + T.m[5] * p.y + T.m[6] * p.z,
                     T.m[8] * p.x + T.m[9] * p.y + T.m[10] * p.z);
}

// se3.cuh:164-168
__device__ __forceinline__ float3 transform(const Pose &T, const float3 p)
{
  const float3 r = rotate(T, p);
  return make_float3(r.x + T.m[3], r.y + T.m[7], r.z + T.m[11]);
}

// pinhole_camera.cuh:48-53
__device__ __forceinline__ float2 project(const Camera &cam, const float3 p)
{
  return make_float2(cam.fx * p.x / p.z + cam.cx, cam.fy * p.y / p.z + cam.cy);
}

// src/seed_check.cu:53-66 (border handled by the caller)
__device__ __forceinline__ int classify_seed(const FilterParams &P, const float4 seed)
{
  const float sigma_sq = seed.y, a = seed.z, b = seed.w;
  if(((a / (a + b)) > P.eta_inlier) && (sigma_sq < P.epsilon))
    return RMD_CONVERGED;
  if((a - 1.0f) / (a + b - 2.0f) < P.eta_outlier)
    return RMD_DIVERGED;
  return RMD_UPDATE;
}

// The epipolar search segment of one seed, src/epipolar_match.cu:60-75.
struct EpiSegment
{
  float2 mean;      // projection of the depth estimate
  float2 dir;       // unit direction of the segment
  float half_len;   // half of min(segment length, 100 px)
};

__device__ __forceinline__ EpiSegment epipolar_segment(
    const FilterParams &P, const int x, const int y, const float mu, const float sigma_sq)
{
  const float sigma = sqrtf(sigma_sq);
  const float3 f = unit_bearing(P.cam, (float)x, (float)y);
  EpiSegment s;
  s.mean = project(P.cam, transform(P.T_curr_ref, scaled(f, mu)));
  const float2 lo = project(P.cam, transform(P.T_curr_ref, scaled(f, fmaxf(mu - 3.0f * sigma, 0.01f))));
  const float2 hi = project(P.cam, transform(P.T_curr_ref, scaled(f, mu + (3.0f * sigma))));
  const float2 line = make_float2(hi.x - lo.x, hi.y - lo.y);
  const float len_sq = line.x * line.x + line.y * line.y;
  const float inv_len = rsqrtf(len_sq);
  s.dir = make_float2(line.x * inv_len, line.y * inv_len);
  // Two defined deviations, both where the reference samples its texture at NaN coordinates
  // (hardware-defined results):
  // 1. an exactly zero-length segment (exact identity motion) makes the reference divide 0 by 0 (SURVEY.md 8a
  //    note 4); here: one candidate at the mean projection;
  // 2. a NaN or infinite segment -- in practice sigma_sq < 0: the posterior variance
  //    c1 (s^2 + m^2) + c2 (sigma^2 + mu^2) - mu'^2 (seed_update.cu:105) of a seed whose variance has collapsed
  //    is a rounding residue of either sign, and sqrtf of a negative one is NaN.  fminf drops the NaN, so the
  //    reference walks 143 candidates at NaN coordinates; measured on B200 (tools/parity_diag.py) it reports
  //    NO_MATCH for such seeds frame after frame (b += 1, mu and sigma_sq frozen).  Here: no candidates, hence
  //    NO_MATCH -- the same fate, without the 143 wasted patches.
  if(len_sq == 0.0f)
    s.dir = make_float2(0.0f, 0.0f);
  s.half_len = 0.5f * fminf(sqrtf(len_sq), RMD_MAX_EPIPOLAR_EXTENT);
  if(!(fabsf(len_sq) < CUDART_INF_F))
    s.half_len = CUDART_NAN_F;     // `l <= half_len` is never true: the search loop does not run
  return s;
}

// True when the candidate patch centre is outside the searchable interior,
// src/epipolar_match.cu:91-97 (NaN coordinates pass, as there).
template<int PS>
__device__ __forceinline__ bool candidate_rejected(const float2 px, const int width, const int height)
{
  return (px.x >= (float)(width - PS)) || (px.y >= (float)(height - PS)) ||
         (px.x < (float)PS) || (px.y < (float)PS);
}

// Integer origin and (optionally quantised) bilinear weights of a candidate.
// The reference samples curr_img_tex at px + d + 0.5 with the texture unit's
// linear filter (epipolar_match.cu:111-114): xB = px + d, i = floor(xB),
// alpha = frac(xB) held in 1.8 fixed point.  All PSxPS taps of a candidate
// share alpha/beta, so the patch is a (PS+1)x(PS+1) texel block at (i0, j0).
struct TapFrame
{
  int i0, j0;
  float wx0, wx1, wy0, wy1;
};

template<int PS>
__device__ __forceinline__ TapFrame tap_frame(const float2 px, const float quant)
{
  const float bx = px.x + (float)(-(PS / 2));
  const float by = px.y + (float)(-(PS / 2));
  TapFrame t;
  float al, be;
  if(quant > 0.0f)
  {
    const float inv_q = 1.0f / quant;
    const float tx = floorf(bx * quant + 0.5f);
    const float ty = floorf(by * quant + 0.5f);
    const float fi = floorf(tx * inv_q);
    const float fj = floorf(ty * inv_q);
    t.i0 = (int)fi;
    t.j0 = (int)fj;
    al = (tx - fi * quant) * inv_q;
    be = (ty - fj * quant) * inv_q;
  }
  else
  {
    const float fi = floorf(bx), fj = floorf(by);
    t.i0 = (int)fi;
    t.j0 = (int)fj;
    al = bx - fi;
    be = by - fj;
  }
  t.wx1 = al; t.wx0 = 1.0f - al;
  t.wy1 = be; t.wy0 = 1.0f - be;
  return t;
}

// NCC of the reference template against the bilinearly resampled PSxPS patch
// of the current image at `frame`: src/epipolar_match.cu:99-123.
// `taps.at(j, i)` returns texel (frame.i0 + i, frame.j0 + j), 0 <= i,j <= PS,
// from global memory or from the shared-memory strip.
//
// Taps::kEdgeFirst (global memory): the first and last texel of all PS+1 rows
// are loaded before anything else.  A row is 24 or 32 bytes, so those two loads
// touch every 32-byte sector the row lives in: 2*(PS+1) independent L2 round
// trips in flight at once, after which the remaining taps are L1 hits.  Left to
// itself the compiler (register-limited) trickles the (PS+1)^2 loads between
// the arithmetic and a candidate costs ~18 serial L2 latencies (measured:
// ~4500 cycles against ~1500 from the shared-memory strip).
template<int PS, typename Taps>
__device__ __forceinline__ float ncc_score(
    const Taps &taps, const TapFrame &t, const float (&templ)[PS * PS],
    const float sum_templ, const float const_templ_denom)
{
  float sum_img = 0.0f, sum_img_sq = 0.0f, sum_img_templ = 0.0f;
  float first[Taps::kEdgeFirst ? PS + 1 : 1], last[Taps::kEdgeFirst ? PS + 1 : 1];
  if(Taps::kEdgeFirst)
  {
#pragma unroll
    for(int j = 0; j <= PS; ++j)
    {
      first[j] = taps.at(j, 0);
      last[j] = taps.at(j, PS);
    }
  }
  float upper[PS];  // horizontally filtered row j (weights wx0/wx1)
#pragma unroll
  for(int j = 0; j <= PS; ++j)
  {
    float v[PS + 1];
#pragma unroll
    for(int i = 0; i <= PS; ++i)
    {
      if(Taps::kEdgeFirst && i == 0) v[i] = first[j];
      else if(Taps::kEdgeFirst && i == PS) v[i] = last[j];
      else v[i] = taps.at(j, i);
    }
    float lower[PS];
#pragma unroll
    for(int i = 0; i < PS; ++i) lower[i] = t.wx0 * v[i] + t.wx1 * v[i + 1];
    if(j > 0)
    {
#pragma unroll
      for(int i = 0; i < PS; ++i)
      {
        const float img = t.wy0 * upper[i] + t.wy1 * lower[i];
        const float tv = templ[(j - 1) * PS + i];
        sum_img += img;
        sum_img_sq += img * img;
        sum_img_templ += img * tv;
      }
    }
#pragma unroll
    for(int i = 0; i < PS; ++i) upper[i] = lower[i];
  }
  const float area = (float)(PS * PS);
  const float numerator = area * sum_img_templ - sum_img * sum_templ;
  const float denominator = (area * sum_img_sq - sum_img * sum_img) * const_templ_denom;
  return numerator * rsqrtf(denominator + FLT_MIN);
}

// Texel block of a candidate read straight from a pitched global image
// (read-only path).  Candidates are confined to [PS, size-PS) so every tap is
// in bounds (epipolar_match.cu:91-97).
struct GlobalTaps
{
  static constexpr bool kEdgeFirst = true;
  const float *origin;
  int stride;
  __device__ __forceinline__ GlobalTaps(const float *img, const int img_stride, const TapFrame &t)
    : origin(img + (size_t)t.j0 * img_stride + t.i0), stride(img_stride) {}
  __device__ __forceinline__ float at(const int j, const int i) const
  {
    return __ldg(origin + j * stride + i);
  }
};

// src/triangulation.cu:30-50
__device__ __forceinline__ float3 triangulate_midpoint(
    const float3 f_ref, const float3 f_curr, const Pose &T_ref_curr)
{
  const float3 t = make_float3(T_ref_curr.m[3], T_ref_curr.m[7], T_ref_curr.m[11]);
  const float3 f2 = rotate(T_ref_curr, f_curr);
  const float bx = dot3(t, f_ref);
  const float by = dot3(t, f2);
  const float a0 = dot3(f_ref, f_ref);
  const float a2 = dot3(f_ref, f2);
  const float a1 = -a2;
  const float a3 = dot3(make_float3(-f2.x, -f2.y, -f2.z), f2);
  const float det = a0 * a3 - a1 * a2;
  const float l1 = (a3 * bx - a1 * by) / det;
  const float l2 = (-a2 * bx + a0 * by) / det;
  const float3 xm = scaled(f_ref, l1);
  const float3 xn = make_float3(t.x + l2 * f2.x, t.y + l2 * f2.y, t.z + l2 * f2.z);
  return make_float3((xm.x + xn.x) / 2.0f, (xm.y + xn.y) / 2.0f, (xm.z + xn.z) / 2.0f);
}

// src/triangulation.cu:53-68
__device__ __forceinline__ float triangulation_uncertainty(
    const float z, const float3 f_ref, const float3 t, const float one_pix_angle)
{
  const float3 a = make_float3(f_ref.x * z - t.x, f_ref.y * z - t.y, f_ref.z * z - t.z);
  const float t_norm = sqrtf(dot3(t, t));
  const float a_norm = sqrtf(dot3(a, a));
  const float alpha = acosf(dot3(f_ref, t) / t_norm);
  const float beta = acosf((-dot3(a, t)) / (t_norm * a_norm));
  const float beta_plus = beta + one_pix_angle;
  const float gamma_plus = (float)(CUDART_PI - (double)alpha - (double)beta_plus);
  const float z_plus = t_norm * sinf(beta_plus) / sinf(gamma_plus);
  return z_plus - z;
}

// src/seed_update.cu:31-37
__device__ __forceinline__ float normal_pdf(const float x, const float mu, const float sigma_sq)
{
  return expf(-(x - mu) * (x - mu) / (2.0f * sigma_sq)) *
         rsqrtf((float)(2.0 * CUDART_PI * (double)sigma_sq));
}

// Vogiatzis-Hernandez posterior update of one seed from the matched pixel,
// src/seed_update.cu:58-110.  Returns false (seed untouched) when the
// triangulated point is behind the camera (:77-80) or the update is NaN
// (:100-103).
__device__ __forceinline__ bool bayes_update(
    const FilterParams &P, const int x, const int y, const float2 match, float4 &seed)
{
  const float mu = seed.x, sigma_sq = seed.y, a = seed.z, b = seed.w;
  const float3 f_ref = unit_bearing(P.cam, (float)x, (float)y);
  const float3 f_match = unit_bearing(P.cam, match.x, match.y);
  const float3 pt = triangulate_midpoint(f_ref, f_match, P.T_ref_curr);
  if(pt.z < 0.0f)
    return false;
  const float depth = sqrtf(dot3(pt, pt));
  const float3 t = make_float3(P.T_ref_curr.m[3], P.T_ref_curr.m[7], P.T_ref_curr.m[11]);
  const float tau = triangulation_uncertainty(depth, f_ref, t, P.one_pix_angle);
  const float tau_sq = tau * tau;
  const float s_sq = (tau_sq * sigma_sq) / (tau_sq + sigma_sq);
  const float m = s_sq * (mu / sigma_sq + depth / tau_sq);
  float c1 = (a / (a + b)) * normal_pdf(depth, mu, sigma_sq + tau_sq);
  float c2 = (b / (a + b)) * (1.0f / P.depth_range);
  const float norm_const = c1 + c2;
  c1 = c1 / norm_const;
  c2 = c2 / norm_const;
  const float f = c1 * ((a + 1.0f) / (a + b + 1.0f)) + c2 * (a / (a + b + 1.0f));
  const float e = c1 * (((a + 1.0f) * (a + 2.0f)) / ((a + b + 1.0f) * (a + b + 2.0f))) +
                  c2 * (a * (a + 1.0f) / ((a + b + 1.0f) * (a + b + 2.0f)));
  if(isnan(c1 * m))
    return false;
  const float mu_prime = c1 * m + c2 * mu;
  seed.y = c1 * (s_sq + m * m) + c2 * (sigma_sq + mu * mu) - mu_prime * mu_prime;
  seed.x = mu_prime;
  const float a_prime = (e - f) / (f - e / f);
  seed.z = a_prime;
  seed.w = a_prime * (1.0f - f) / f;
  return true;
}

} // namespace rmdb
