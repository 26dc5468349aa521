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
#include "packed_f32x2.cuh"
#include "candidate_range.cuh"

namespace rmdb
{

#define RMD_MAX_EPIPOLAR_EXTENT 100.0f  // RMD_MAX_EXTENT_EPIPOLAR_SEARCH, CMakeLists.txt:53
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

// ---------------------------------------------------------------------------------------------
// Every rounding of the per-seed and per-candidate arithmetic is PINNED.
//
// The reference builds with -use_fast_math, i.e. nvcc contracts a*b + c*d into ONE fused multiply-add
// and one plain product -- and which of the two products stays plain is ptxas' choice (it depends on
// which operand is ready first, so it differs between otherwise identical instantiations of the same
// source: measured here, a kernel that reads the pose from a dynamically indexed parameter block
// rounded ~12 % of the updated seeds differently, by 1-2 ulp, from the one that reads it from fixed
// constant-bank slots).  Results that are bit-identical to the reference's own build
// (tests/test_ref_cuda_parity.py) must not hang on such a choice: below, every sum of products is
// written with explicit __fmaf_rn / __fmul_rn in exactly the form the reference's kernels
// (src/epipolar_match.cu:60-123, src/seed_update.cu:40-121, src/triangulation.cu:30-68, rebuilt for
// sm_100a: oracle/_ref) and the round-1 build of this file compile to -- read off their SASS -- so
// that every kernel that includes this header (direct, staged, batched; 5x5 and 7x7) performs the same
// roundings whatever ptxas would have chosen.  The pattern throughout: in x*a + y*b (+ z*c) the y
// product is the plain one, x and z are fused onto it.
// Divisions are the fast-math form x * rcp.approx(y), as -use_fast_math emits them.

__device__ __forceinline__ float rcp_approx(const float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// a.x*b.x + a.y*b.y + a.z*b.z as the reference's build rounds it: the y product is the plain one
__device__ __forceinline__ float dot3_pinned(const float3 a, const float3 b)
{
  return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}

// normalize(cam.cam2world(px)) inside the update: pinhole_camera.cuh:40-46, helper_math.h:1309-1313
__device__ __forceinline__ float3 unit_bearing_pinned(const Camera &cam, const float u, const float v)
{
  const float px = __fmul_rn(__fadd_rn(u, -cam.cx), rcp_approx(cam.fx));
  const float py = __fmul_rn(__fadd_rn(v, -cam.cy), rcp_approx(cam.fy));
  const float inv = rsqrtf(__fadd_rn(__fmaf_rn(px, px, __fmul_rn(py, py)), 1.0f));
  return make_float3(__fmul_rn(px, inv), __fmul_rn(py, inv), inv);
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

// se3.cuh:164-168 + pinhole_camera.cuh:48-53: project(cam, T * p), roundings pinned
__device__ __forceinline__ float2 transform_project_pinned(const Camera &cam, const Pose &T, const float3 p)
{
  const float X = __fadd_rn(__fmaf_rn(p.z, T.m[2], __fmaf_rn(p.x, T.m[0], __fmul_rn(p.y, T.m[1]))), T.m[3]);
  const float Y = __fadd_rn(__fmaf_rn(p.z, T.m[6], __fmaf_rn(p.x, T.m[4], __fmul_rn(p.y, T.m[5]))), T.m[7]);
  const float Z = __fadd_rn(__fmaf_rn(p.z, T.m[10], __fmaf_rn(p.x, T.m[8], __fmul_rn(p.y, T.m[9]))), T.m[11]);
  const float inv_z = rcp_approx(Z);
  return make_float2(__fmaf_rn(__fmul_rn(X, cam.fx), inv_z, cam.cx), __fmaf_rn(__fmul_rn(Y, cam.fy), inv_z, cam.cy));
}

__device__ __forceinline__ float3 scaled_pinned(const float3 v, const float s)
{
  return make_float3(__fmul_rn(v.x, s), __fmul_rn(v.y, s), __fmul_rn(v.z, s));
}

__device__ __forceinline__ EpiSegment epipolar_segment(
    const FilterParams &P, const int x, const int y, const float mu, const float sigma_sq)
{
  const float sigma = sqrtf(sigma_sq);
  const float3 f = unit_bearing_pinned(P.cam, (float)x, (float)y);
  EpiSegment s;
  s.mean = transform_project_pinned(P.cam, P.T_curr_ref, scaled_pinned(f, mu));
  const float2 lo = transform_project_pinned(P.cam, P.T_curr_ref,
                                             scaled_pinned(f, fmaxf(__fmaf_rn(sigma, -3.0f, mu), 0.01f)));
  const float2 hi = transform_project_pinned(P.cam, P.T_curr_ref, scaled_pinned(f, __fmaf_rn(sigma, 3.0f, mu)));
  const float2 line = make_float2(__fadd_rn(hi.x, -lo.x), __fadd_rn(hi.y, -lo.y));
  const float len_sq = __fmaf_rn(line.x, line.x, __fmul_rn(line.y, line.y));
  const float inv_len = rsqrtf(len_sq);
  s.dir = make_float2(__fmul_rn(line.x, inv_len), __fmul_rn(line.y, inv_len));
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
  s.half_len = __fmul_rn(fminf(sqrtf(len_sq), RMD_MAX_EPIPOLAR_EXTENT), 0.5f);
  if(!(fabsf(len_sq) < CUDART_INF_F))
    s.half_len = CUDART_NAN_F;     // `l <= half_len` is never true: the search loop does not run
  return s;
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
  const float bx = __fadd_rn(px.x, (float)(-(PS / 2)));
  const float by = __fadd_rn(px.y, (float)(-(PS / 2)));
  TapFrame t;
  float al, be;
  if(quant > 0.0f)
  {
    const float inv_q = rcp_approx(quant);
    const float tx = floorf(__fmaf_rn(bx, quant, 0.5f));
    const float ty = floorf(__fmaf_rn(by, quant, 0.5f));
    const float fi = floorf(__fmul_rn(tx, inv_q));
    const float fj = floorf(__fmul_rn(ty, inv_q));
    t.i0 = (int)fi;
    t.j0 = (int)fj;
    al = __fmul_rn(__fmaf_rn(-fi, quant, tx), inv_q);
    be = __fmul_rn(__fmaf_rn(-fj, quant, ty), inv_q);
  }
  else
  {
    const float fi = floorf(bx), fj = floorf(by);
    t.i0 = (int)fi;
    t.j0 = (int)fj;
    al = __fadd_rn(bx, -fi);
    be = __fadd_rn(by, -fj);
  }
  t.wx1 = al; t.wx0 = __fadd_rn(1.0f, -al);
  t.wy1 = be; t.wy0 = __fadd_rn(1.0f, -be);
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
    for(int i = 0; i < PS; ++i) lower[i] = __fmaf_rn(v[i], t.wx0, __fmul_rn(v[i + 1], t.wx1));
    if(j > 0)
    {
#pragma unroll
      for(int i = 0; i < PS; ++i)
      {
        const float img = __fmaf_rn(upper[i], t.wy0, __fmul_rn(lower[i], t.wy1));
        const float tv = templ[(j - 1) * PS + i];
        sum_img = __fadd_rn(sum_img, img);
        sum_img_sq = __fmaf_rn(img, img, sum_img_sq);
        sum_img_templ = __fmaf_rn(img, tv, sum_img_templ);
      }
    }
#pragma unroll
    for(int i = 0; i < PS; ++i) upper[i] = lower[i];
  }
  const float area = (float)(PS * PS);
  const float numerator = __fmaf_rn(sum_img_templ, area, -__fmul_rn(sum_img, sum_templ));
  const float spread = __fmaf_rn(sum_img_sq, area, -__fmul_rn(sum_img, sum_img));
  return __fmul_rn(numerator, rsqrtf(__fmaf_rn(const_templ_denom, spread, FLT_MIN)));
}

// Two candidates at once: lane 0 of every packed register belongs to candidate A, lane 1 to candidate B.
// The texels of the two blocks are loaded into adjacent registers, the separable filter and the sums of
// squares run as f32x2 operations (one issue slot for both candidates), the template products stay scalar
// (the template is shared by the two lanes).  Each lane performs exactly ncc_score's operations in
// ncc_score's order, so the two results are bit for bit those of two ncc_score calls.
template<int PS, typename Taps>
__device__ __forceinline__ float2 ncc_score_pair(
    const Taps &taps_a, const TapFrame &ta, const Taps &taps_b, const TapFrame &tb, const float (&templ)[PS * PS],
    const float sum_templ, const float const_templ_denom)
{
  const f2 wx0 = pack(ta.wx0, tb.wx0), wx1 = pack(ta.wx1, tb.wx1);
  const f2 wy0 = pack(ta.wy0, tb.wy0), wy1 = pack(ta.wy1, tb.wy1);
  f2 sum_img = pack(0.0f, 0.0f), sum_img_sq = pack(0.0f, 0.0f);
  float sum_it_a = 0.0f, sum_it_b = 0.0f;
  f2 upper[PS];
#pragma unroll
  for(int j = 0; j <= PS; ++j)
  {
    f2 v[PS + 1];
#pragma unroll
    for(int i = 0; i <= PS; ++i) v[i] = pack(taps_a.at(j, i), taps_b.at(j, i));
    f2 lower[PS];
#pragma unroll
    for(int i = 0; i < PS; ++i) lower[i] = f2_fma(v[i], wx0, f2_mul(v[i + 1], wx1));
    if(j > 0)
    {
#pragma unroll
      for(int i = 0; i < PS; ++i)
      {
        const f2 img = f2_fma(upper[i], wy0, f2_mul(lower[i], wy1));
        const float2 im = unpack(img);
        const float tv = templ[(j - 1) * PS + i];
        sum_img = f2_add(sum_img, img);
        sum_img_sq = f2_fma(img, img, sum_img_sq);
        sum_it_a = __fmaf_rn(im.x, tv, sum_it_a);
        sum_it_b = __fmaf_rn(im.y, tv, sum_it_b);
      }
    }
#pragma unroll
    for(int i = 0; i < PS; ++i) upper[i] = lower[i];
  }
  const float area = (float)(PS * PS);
  const float2 si = unpack(sum_img), sq = unpack(sum_img_sq);
  const float num_a = __fmaf_rn(sum_it_a, area, -__fmul_rn(si.x, sum_templ));
  const float num_b = __fmaf_rn(sum_it_b, area, -__fmul_rn(si.y, sum_templ));
  const float spread_a = __fmaf_rn(sq.x, area, -__fmul_rn(si.x, si.x));
  const float spread_b = __fmaf_rn(sq.y, area, -__fmul_rn(si.y, si.y));
  return make_float2(__fmul_rn(num_a, rsqrtf(__fmaf_rn(const_templ_denom, spread_a, FLT_MIN))),
                     __fmul_rn(num_b, rsqrtf(__fmaf_rn(const_templ_denom, spread_b, FLT_MIN))));
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

// src/seed_update.cu:31-37
__device__ __forceinline__ float normal_pdf(const float x, const float mu, const float sigma_sq)
{
  const float d = __fadd_rn(x, -mu);
  const float arg = __fmul_rn(__fmul_rn(d, -d), rcp_approx(__fadd_rn(sigma_sq, sigma_sq)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(arg, 1.4426950216293334961f)));   // __expf
  return __fmul_rn(e, rsqrtf((float)(2.0 * CUDART_PI * (double)sigma_sq)));
}

// Vogiatzis-Hernandez posterior update of one seed from the matched pixel: triangulation
// (src/triangulation.cu:30-50), one-pixel uncertainty (:53-68), moment-matched Gaussian x Beta update
// (src/seed_update.cu:58-110).  Returns false (seed untouched) when the triangulated point is behind the
// camera (:77-80) or the update is NaN (:100-103).
__device__ __forceinline__ bool bayes_update(
    const FilterParams &P, const int x, const int y, const float2 match, float4 &seed)
{
  const float mu = seed.x, sigma_sq = seed.y, a = seed.z, b = seed.w;
  const Pose &T = P.T_ref_curr;
  const float3 t = make_float3(T.m[3], T.m[7], T.m[11]);
  const float3 f_ref = unit_bearing_pinned(P.cam, (float)x, (float)y);
  const float3 f_cur = unit_bearing_pinned(P.cam, match.x, match.y);
  // f2 = R_ref_curr * f_curr (se3.cuh:111-117): per row, the y product is the plain one
  const float3 f2 = make_float3(
      __fmaf_rn(f_cur.z, T.m[2], __fmaf_rn(f_cur.x, T.m[0], __fmul_rn(f_cur.y, T.m[1]))),
      __fmaf_rn(f_cur.z, T.m[6], __fmaf_rn(f_cur.x, T.m[4], __fmul_rn(f_cur.y, T.m[5]))),
      __fmaf_rn(f_cur.z, T.m[10], __fmaf_rn(f_cur.x, T.m[8], __fmul_rn(f_cur.y, T.m[9]))));
  // 2x2 system of triangulatenNonLin: A = [f1.f1, -f1.f2; f1.f2, -f2.f2], rhs = [t.f1, t.f2]
  const float a2 = dot3_pinned(f_ref, f2);
  const float a0 = dot3_pinned(f_ref, f_ref);
  const float a3 = __fmaf_rn(-f2.z, f2.z, __fmaf_rn(f2.y, -f2.y, -__fmul_rn(f2.x, f2.x)));   // dot(-f2, f2)
  const float inv_det = rcp_approx(__fmaf_rn(a0, a3, __fmul_rn(a2, a2)));                   // a0 a3 - a1 a2, a1 = -a2
  const float bx = dot3_pinned(f_ref, t);
  const float by = dot3_pinned(f2, t);
  const float by_a2 = __fmul_rn(by, a2), bx_a2 = __fmul_rn(bx, a2);
  const float l1 = __fmul_rn(__fmaf_rn(bx, a3, by_a2), inv_det);                            // (a3 bx - a1 by) / det
  const float l2 = __fmul_rn(__fmaf_rn(by, a0, -bx_a2), inv_det);                           // (-a2 bx + a0 by) / det
  // midpoint (l1 f1 + t + l2 f2) / 2
  const float3 pt = make_float3(
      __fmul_rn(__fmaf_rn(f_ref.x, l1, __fmaf_rn(f2.x, l2, t.x)), 0.5f),
      __fmul_rn(__fmaf_rn(f_ref.y, l1, __fmaf_rn(f2.y, l2, t.y)), 0.5f),
      __fmul_rn(__fmaf_rn(f_ref.z, l1, __fmaf_rn(f2.z, l2, t.z)), 0.5f));
  if(pt.z < 0.0f)
    return false;
  const float depth = sqrtf(__fmaf_rn(pt.z, pt.z, __fmaf_rn(pt.y, pt.y, __fmul_rn(pt.x, pt.x))));

  // triangulationUncertainty: tau = z+ - z with beta+ = beta + one pixel, gamma+ = pi - alpha - beta+ (double)
  const float t_norm = sqrtf(dot3_pinned(t, t));
  const float3 av = make_float3(__fmaf_rn(f_ref.x, depth, -t.x), __fmaf_rn(f_ref.y, depth, -t.y),
                                __fmaf_rn(f_ref.z, depth, -t.z));
  const float a_norm = sqrtf(dot3_pinned(av, av));
  const float alpha = acosf(__fmul_rn(bx, rcp_approx(t_norm)));
  const float beta = acosf(__fmul_rn(dot3_pinned(av, t), -rcp_approx(__fmul_rn(t_norm, a_norm))));
  const float beta_plus = __fadd_rn(beta, P.one_pix_angle);
  const float gamma_plus = (float)(CUDART_PI - (double)alpha - (double)beta_plus);
  const float tau = __fmaf_rn(__fmul_rn(t_norm, sinf(beta_plus)), rcp_approx(sinf(gamma_plus)), -depth);
  const float tau_sq = __fmul_rn(tau, tau);

  const float var_sum = __fadd_rn(sigma_sq, tau_sq);
  const float s_sq = __fmul_rn(__fmul_rn(sigma_sq, tau_sq), rcp_approx(var_sum));
  const float m = __fmul_rn(s_sq, __fmaf_rn(mu, rcp_approx(sigma_sq), __fmul_rn(depth, rcp_approx(tau_sq))));
  const float ab = __fadd_rn(a, b);
  const float inv_ab = rcp_approx(ab);
  float c1 = __fmul_rn(__fmul_rn(a, inv_ab), normal_pdf(depth, mu, var_sum));
  float c2 = __fmul_rn(__fmul_rn(b, inv_ab), rcp_approx(P.depth_range));
  const float inv_norm = rcp_approx(__fadd_rn(c1, c2));
  c1 = __fmul_rn(c1, inv_norm);
  c2 = __fmul_rn(c2, inv_norm);
  const float ab1 = __fadd_rn(ab, 1.0f), ab2 = __fadd_rn(ab, 2.0f), a1 = __fadd_rn(a, 1.0f), a_2 = __fadd_rn(a, 2.0f);
  const float inv_ab1 = rcp_approx(ab1), inv_d = rcp_approx(__fmul_rn(ab1, ab2));
  const float f = __fmaf_rn(c1, __fmul_rn(a1, inv_ab1), __fmul_rn(c2, __fmul_rn(a, inv_ab1)));
  const float e = __fmaf_rn(c1, __fmul_rn(__fmul_rn(a1, a_2), inv_d), __fmul_rn(c2, __fmul_rn(__fmul_rn(a, a1), inv_d)));
  const float c1_m = __fmul_rn(c1, m);
  if(isnan(c1_m))
    return false;
  const float mu_prime = __fmaf_rn(mu, c2, c1_m);
  const float second = __fmaf_rn(c1, __fmaf_rn(m, m, s_sq), __fmul_rn(c2, __fmaf_rn(mu, mu, sigma_sq)));
  seed.y = __fmaf_rn(-mu_prime, mu_prime, second);
  seed.x = mu_prime;
  const float inv_f = rcp_approx(f);
  const float a_prime = __fmul_rn(__fadd_rn(e, -f), rcp_approx(__fmaf_rn(-e, inv_f, f)));   // (e - f) / (f - e / f)
  seed.z = a_prime;
  seed.w = __fmul_rn(__fmul_rn(a_prime, __fadd_rn(1.0f, -f)), inv_f);
  return true;
}

} // namespace rmdb
