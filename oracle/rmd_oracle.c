/*
 * rmd_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C / IEEE fp32 restatement of the CUDA kernels on REMODE's depth
 * filter hot path.  Every function cites the reference lines it follows
 * (paths relative to /root/reference).  Expression order is kept as in the
 * reference so that fp32 rounding happens in the same places; build with
 * -ffp-contract=off so the compiler does not fuse multiply-adds behind our
 * back.  What cannot be restated from source is modelled and pinned:
 *   - texture-unit bilinear filtering of the current image (weights
 *     quantised to `tex_frac_bits` fractional bits, clamp addressing);
 *   - the reference is built -use_fast_math (approximate div/sqrt/exp/sin,
 *     flush-to-zero); this oracle is IEEE.  Tolerances in tests/ absorb it
 *     and the rebuilt reference CUDA (oracle/_ref) removes it as a confounder.
 *
 * Parity status: pinned against the reference's re-hosted known-answer tests
 * (tests/test_oracle_pins.py) and, on the GPU box, against the reference's
 * own kernels rebuilt for sm_100a (tests/test_ref_cuda_parity.py).
 */
#include "rmd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_EXTENT_EPIPOLAR_SEARCH 100 /* CMakeLists.txt:53 */

typedef struct { float x, y; } f2;
typedef struct { float x, y, z; } f3;

struct rmd_oracle_seeds {
  int w, h, patch;
  float fx, fy, cx, cy;
  int tex_frac_bits;
  /* scene + algorithm parameters, src/seed_matrix.cu:96-104 */
  float min_depth, max_depth, avg_depth, depth_range, sigma_sq_max;
  float eta_inlier, eta_outlier, epsilon;
  float T_world_ref[12];
  float T_curr_ref[12];
  float dist_from_ref;
  float *ref_img, *sum_templ, *const_templ_denom;
  float *mu, *sigma_sq, *a, *b;
  int *convergence;
  f2 *matches;
};

/* ---------------------------------------------------------------- helpers */

static float rsqrt_host(float x) { return 1.0f / sqrtf(x); } /* helper_math.h:62-65 */
static float fmin_host(float a, float b) { return a < b ? a : b; } /* helper_math.h:40-43 */
static float fmax_host(float a, float b) { return a > b ? a : b; } /* helper_math.h:45-48 */

static float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* helper_math.h:1248 */
static float dot2(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }             /* helper_math.h:1244 */
static f3 scale3(f3 v, float s) { f3 r = {v.x * s, v.y * s, v.z * s}; return r; }
static f3 normalize3(f3 v) { return scale3(v, rsqrt_host(dot3(v, v))); } /* helper_math.h:1309-1313 */
static f2 normalize2(f2 v) { /* helper_math.h:1304-1308 */
  float inv = rsqrt_host(dot2(v, v));
  f2 r = {v.x * inv, v.y * inv};
  return r;
}
static float norm3(f3 v) { return sqrtf(dot3(v, v)); } /* helper_vector_types.cuh:23-28 */
static float norm2(f2 v) { return sqrtf(dot2(v, v)); }

/* include/rmd/se3.cuh:111-117 */
static f3 se3_rotate(const float *T, f3 p) {
  f3 r = {T[0] * p.x + T[1] * p.y + T[2] * p.z,
          T[4] * p.x + T[5] * p.y + T[6] * p.z,
          T[8] * p.x + T[9] * p.y + T[10] * p.z};
  return r;
}
/* se3.cuh:120-125 and :164-168 */
static f3 se3_apply(const float *T, f3 p) {
  f3 r = se3_rotate(T, p);
  r.x = r.x + T[3];
  r.y = r.y + T[7];
  r.z = r.z + T[11];
  return r;
}
static f3 se3_translation(const float *T) { f3 t = {T[3], T[7], T[11]}; return t; }

/* se3.cuh:81-97 */
void rmd_oracle_se3_inv(const float *d, float *o) {
  float r[12];
  r[0] = d[0]; r[1] = d[4]; r[2] = d[8];
  r[4] = d[1]; r[5] = d[5]; r[6] = d[9];
  r[8] = d[2]; r[9] = d[6]; r[10] = d[10];
  r[3] = -d[0] * d[3] - d[4] * d[7] - d[8] * d[11];
  r[7] = -d[1] * d[3] - d[5] * d[7] - d[9] * d[11];
  r[11] = -d[2] * d[3] - d[6] * d[7] - d[10] * d[11];
  memcpy(o, r, sizeof r);
}
/* se3.cuh:146-162 */
void rmd_oracle_se3_mul(const float *l, const float *r, float *o) {
  float t[12];
  for (int row = 0; row < 3; ++row) {
    const float *L = l + 4 * row;
    t[4 * row + 0] = L[0] * r[0] + L[1] * r[4] + L[2] * r[8];
    t[4 * row + 1] = L[0] * r[1] + L[1] * r[5] + L[2] * r[9];
    t[4 * row + 2] = L[0] * r[2] + L[1] * r[6] + L[2] * r[10];
    t[4 * row + 3] = L[3] + L[0] * r[3] + L[1] * r[7] + L[2] * r[11];
  }
  memcpy(o, t, sizeof t);
}
/* se3.cuh:37-66 */
void rmd_oracle_se3_from_quat(float qw, float qx, float qy, float qz, float tx,
                              float ty, float tz, float *o) {
  const float x = 2 * qx, y = 2 * qy, z = 2 * qz;
  const float wx = x * qw, wy = y * qw, wz = z * qw;
  const float xx = x * qx, xy = y * qx, xz = z * qx;
  const float yy = y * qy, yz = z * qy, zz = z * qz;
  o[0] = 1 - (yy + zz); o[1] = xy - wz;       o[2] = xz + wy;       o[3] = tx;
  o[4] = xy + wz;       o[5] = 1 - (xx + zz); o[6] = yz - wx;       o[7] = ty;
  o[8] = xz - wy;       o[9] = yz + wx;       o[10] = 1 - (xx + yy); o[11] = tz;
}

/* include/rmd/pinhole_camera.cuh:40-53 */
static f3 cam2world(const rmd_oracle_seeds *s, f2 uv) {
  f3 r = {(uv.x - s->cx) / s->fx, (uv.y - s->cy) / s->fy, 1.0f};
  return r;
}
static f2 world2cam(const rmd_oracle_seeds *s, f3 p) {
  f2 r = {s->fx * p.x / p.z + s->cx, s->fy * p.y / p.z + s->cy};
  return r;
}
/* pinhole_camera.cuh:55-59 */
static float one_pix_angle(const rmd_oracle_seeds *s) {
  return atan2f(1.0f, 2.0f * s->fx) * 2.0f;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Texture fetch at a texel centre (coordinate i+0.5): exact texel, clamp
 * addressing.  include/rmd/texture_memory.cuh:45-52 (clamp, unnormalised). */
static float tex_centre(const float *img, int w, int h, int x, int y) {
  return img[(size_t)clampi(y, 0, h - 1) * w + clampi(x, 0, w - 1)];
}

/* Texture fetch with cudaFilterModeLinear at unnormalised coordinate (x, y)
 * (texture_memory.cuh:48: linear is the default bind mode and the one used
 * for curr_img, src/seed_matrix.cu:130).  CUDA programming-guide model:
 * xB = x - 0.5, i = floor(xB), alpha = frac(xB) held in 1.8 fixed point. */
static float tex_linear(const float *img, int w, int h, float x, float y,
                        int frac_bits) {
  const float xb = x - 0.5f, yb = y - 0.5f;
  int i, j;
  float al, be;
  if (frac_bits > 0) {
    const float q = (float)(1 << frac_bits);
    const float tx = floorf(xb * q + 0.5f), ty = floorf(yb * q + 0.5f);
    const float fi = floorf(tx / q), fj = floorf(ty / q);
    i = (int)fi;
    j = (int)fj;
    al = (tx - fi * q) / q;
    be = (ty - fj * q) / q;
  } else {
    const float fi = floorf(xb), fj = floorf(yb);
    i = (int)fi;
    j = (int)fj;
    al = xb - fi;
    be = yb - fj;
  }
  const float t00 = tex_centre(img, w, h, i, j);
  const float t10 = tex_centre(img, w, h, i + 1, j);
  const float t01 = tex_centre(img, w, h, i, j + 1);
  const float t11 = tex_centre(img, w, h, i + 1, j + 1);
  return (1.0f - al) * (1.0f - be) * t00 + al * (1.0f - be) * t10 +
         (1.0f - al) * be * t01 + al * be * t11;
}

/* -------------------------------------------------------------- lifecycle */

rmd_oracle_seeds *rmd_oracle_seeds_create(int width, int height, float fx,
                                          float fy, float cx, float cy,
                                          int patch) {
  if (width <= 0 || height <= 0 || patch < 1 || (patch & 1) == 0) return NULL;
  rmd_oracle_seeds *s = (rmd_oracle_seeds *)calloc(1, sizeof *s);
  if (!s) return NULL;
  const size_t n = (size_t)width * height;
  s->w = width; s->h = height; s->patch = patch;
  s->fx = fx; s->fy = fy; s->cx = cx; s->cy = cy;
  s->tex_frac_bits = 8;
  s->ref_img = (float *)calloc(n, sizeof(float));
  s->sum_templ = (float *)calloc(n, sizeof(float));
  s->const_templ_denom = (float *)calloc(n, sizeof(float));
  s->mu = (float *)calloc(n, sizeof(float));
  s->sigma_sq = (float *)calloc(n, sizeof(float));
  s->a = (float *)calloc(n, sizeof(float));
  s->b = (float *)calloc(n, sizeof(float));
  s->convergence = (int *)calloc(n, sizeof(int));
  s->matches = (f2 *)calloc(n, sizeof(f2));
  return s;
}

void rmd_oracle_seeds_destroy(rmd_oracle_seeds *s) {
  if (!s) return;
  free(s->ref_img); free(s->sum_templ); free(s->const_templ_denom);
  free(s->mu); free(s->sigma_sq); free(s->a); free(s->b);
  free(s->convergence); free(s->matches);
  free(s);
}

void rmd_oracle_seeds_set_tex_model(rmd_oracle_seeds *s, int frac_bits) {
  s->tex_frac_bits = frac_bits;
}

void rmd_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int rmd_oracle_get_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void *rmd_oracle_seeds_field(rmd_oracle_seeds *s, int field) {
  switch (field) {
    case RMD_O_F_MU: return s->mu;
    case RMD_O_F_SIGMA_SQ: return s->sigma_sq;
    case RMD_O_F_A: return s->a;
    case RMD_O_F_B: return s->b;
    case RMD_O_F_CONVERGENCE: return s->convergence;
    case RMD_O_F_SUM_TEMPL: return s->sum_templ;
    case RMD_O_F_CONST_TEMPL_DENOM: return s->const_templ_denom;
    case RMD_O_F_EPIPOLAR_MATCHES: return s->matches;
    case RMD_O_F_REF_IMG: return s->ref_img;
    default: return NULL;
  }
}

/* ------------------------------------------------------------- seed init */

/* src/seed_init.cu:28-61 (kernel) */
static void stage_init(rmd_oracle_seeds *s) {
  const int w = s->w, h = s->h, P = s->patch;
  const int off = -P / 2;                 /* mvs_device_data.cuh:42 */
  const int area = P * P;                 /* mvs_device_data.cuh:43 */
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      float sum_templ = 0.0f, sum_templ_sq = 0.0f;
      for (int py = 0; py < P; ++py)       /* seed_init.cu:39-51 */
        for (int px = 0; px < P; ++px) {
          const float t = tex_centre(s->ref_img, w, h, x + off + px, y + off + py);
          sum_templ += t;
          sum_templ_sq += t * t;
        }
      const size_t k = (size_t)y * w + x;
      s->sum_templ[k] = sum_templ;
      /* seed_init.cu:53-54: combine in double, store as float */
      s->const_templ_denom[k] =
          (float)((double)area * sum_templ_sq - (double)sum_templ * sum_templ);
      s->mu[k] = s->avg_depth;             /* seed_init.cu:57-60 */
      s->sigma_sq[k] = s->sigma_sq_max;
      s->a[k] = 10.0f;
      s->b[k] = 10.0f;
    }
  }
}

/* src/seed_matrix.cu:87-118 (host driver) */
int rmd_oracle_seeds_set_reference(rmd_oracle_seeds *s, const float *img,
                                   const float *T_curr_world, float min_depth,
                                   float max_depth) {
  memcpy(s->ref_img, img, (size_t)s->w * s->h * sizeof(float));
  s->min_depth = min_depth;
  s->max_depth = max_depth;
  s->avg_depth = (min_depth + max_depth) / 2.0f;
  s->depth_range = max_depth - min_depth;
  s->sigma_sq_max = s->depth_range * s->depth_range / 36.0f;
  s->eta_inlier = 0.7f;
  s->eta_outlier = 0.05f;
  s->epsilon = s->depth_range / 1000.0f;
  rmd_oracle_se3_inv(T_curr_world, s->T_world_ref);
  stage_init(s);
  return 1;
}

/* ------------------------------------------------------------ seed check */

/* src/seed_check.cu:29-67 */
void rmd_oracle_stage_check(rmd_oracle_seeds *s) {
  const int w = s->w, h = s->h, P = s->patch;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      const size_t k = (size_t)y * w + x;
      if (x > w - P - 1 || y > h - P - 1 || x < P || y < P) { /* :37-42 */
        s->convergence[k] = RMD_O_BORDER;
        continue;
      }
      const float sigma_sq = s->sigma_sq[k];
      const float a = s->a[k], b = s->b[k];
      if ((a / (a + b)) > s->eta_inlier && sigma_sq < s->epsilon) /* :54-55 */
        s->convergence[k] = RMD_O_CONVERGED;
      else if ((a - 1) / (a + b - 2) < s->eta_outlier)           /* :59 */
        s->convergence[k] = RMD_O_DIVERGED;
      else
        s->convergence[k] = RMD_O_UPDATE;
    }
  }
}

/* -------------------------------------------------------- epipolar match */

/* src/epipolar_match.cu:38-140 */
void rmd_oracle_stage_match(rmd_oracle_seeds *s, const float *curr,
                            const float *T_curr_ref) {
  const int w = s->w, h = s->h, P = s->patch;
  const int off = -P / 2;
  const float area = (float)(P * P);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      const size_t k = (size_t)y * w + x;
      const int state = s->convergence[k];
      if (state == RMD_O_BORDER || state == RMD_O_CONVERGED ||
          state == RMD_O_DIVERGED) /* :51-57 */
        continue;

      const float mu = s->mu[k];
      const float sigma = sqrtf(s->sigma_sq[k]); /* :60-61 */
      const f2 px_ref = {(float)x, (float)y};
      const f3 f_ref = normalize3(cam2world(s, px_ref)); /* :63-64 */
      const f2 px_mean = world2cam(s, se3_apply(T_curr_ref, scale3(f_ref, mu)));
      const f2 px_min = world2cam(
          s, se3_apply(T_curr_ref,
                       scale3(f_ref, fmax_host(mu - 3.0f * sigma, 0.01f)))); /* :68-69 */
      const f2 px_max = world2cam(
          s, se3_apply(T_curr_ref, scale3(f_ref, mu + (3.0f * sigma)))); /* :70-71 */

      const f2 epi_line = {px_max.x - px_min.x, px_max.y - px_min.y};
      const f2 epi_dir = normalize2(epi_line); /* :74 */
      float half_length =
          0.5f * fmin_host(norm2(epi_line), (float)MAX_EXTENT_EPIPOLAR_SEARCH);
      /* Defined deviation 2: a NaN / infinite segment (sigma_sq < 0: the variance of a
       * collapsed seed is a rounding residue of either sign) makes the reference walk 143 candidates at NaN
       * texture coordinates -- hardware-defined; measured on B200 it reports NO_MATCH.  Here, as in the
       * product: no candidates, NO_MATCH. */
      if (!isfinite(dot2(epi_line, epi_line))) half_length = NAN;

      const float sum_templ = s->sum_templ[k];
      const float const_templ_denom = s->const_templ_denom[k];

      float best_ncc = -1.0f;
      f2 best_px = {0.0f, 0.0f};
      for (float l = -half_length; l <= half_length; l += 0.7f) { /* :88 */
        f2 px;
        px.x = px_mean.x + l * epi_dir.x;
        px.y = px_mean.y + l * epi_dir.y;
        if (px.x >= (float)(w - P) || px.y >= (float)(h - P) ||
            px.x < (float)P || px.y < (float)P) /* :91-97 */
          continue;
        float sum_img = 0.0f, sum_img_sq = 0.0f, sum_img_templ = 0.0f;
        for (int py = 0; py < P; ++py)
          for (int pxi = 0; pxi < P; ++pxi) { /* :103-119 */
            const float templ =
                tex_centre(s->ref_img, w, h, x + off + pxi, y + off + py);
            const float cx_ = px.x + (float)(off + pxi) + 0.5f;
            const float cy_ = px.y + (float)(off + py) + 0.5f;
            const float img = tex_linear(curr, w, h, cx_, cy_, s->tex_frac_bits);
            sum_img += img;
            sum_img_sq += img * img;
            sum_img_templ += img * templ;
          }
        const float num = area * sum_img_templ - sum_img * sum_templ; /* :120 */
        const float den =
            (area * sum_img_sq - sum_img * sum_img) * const_templ_denom; /* :121 */
        const float ncc = num * rsqrt_host(den + FLT_MIN); /* :123 */
        if (ncc > best_ncc) { /* :125-129, strict: first maximum wins */
          best_px = px;
          best_ncc = ncc;
        }
      }
      if (best_ncc < 0.5f) { /* :131-139 */
        s->convergence[k] = RMD_O_NO_MATCH;
      } else {
        s->matches[k] = best_px;
        s->convergence[k] = RMD_O_UPDATE;
      }
    }
  }
}

/* ---------------------------------------- triangulation + Bayesian update */

/* src/triangulation.cu:30-50 */
static f3 triangulate_non_lin(f3 f1, f3 f_curr, const float *T_ref_curr) {
  const f3 t = se3_translation(T_ref_curr);
  const f3 f2v = se3_rotate(T_ref_curr, f_curr);
  const float bx = dot3(t, f1), by = dot3(t, f2v);
  float A[4];
  A[0] = dot3(f1, f1);
  A[2] = dot3(f1, f2v);
  A[1] = -A[2];
  {
    const f3 neg = {-f2v.x, -f2v.y, -f2v.z};
    A[3] = dot3(neg, f2v);
  }
  const float det = A[0] * A[3] - A[1] * A[2];
  const float lx = (A[3] * bx - A[1] * by) / det;
  const float ly = (-A[2] * bx + A[0] * by) / det;
  const f3 xm = {lx * f1.x, lx * f1.y, lx * f1.z};
  const f3 xn = {t.x + ly * f2v.x, t.y + ly * f2v.y, t.z + ly * f2v.z};
  const f3 r = {(xm.x + xn.x) / 2.0f, (xm.y + xn.y) / 2.0f, (xm.z + xn.z) / 2.0f};
  return r;
}

/* src/triangulation.cu:53-68 */
static float triangulation_uncertainty(float z, f3 f, f3 t, float one_pix) {
  const f3 a = {f.x * z - t.x, f.y * z - t.y, f.z * z - t.z};
  const float t_norm = norm3(t);
  const float a_norm = norm3(a);
  const float alpha = acosf(dot3(f, t) / t_norm);
  const float beta = acosf((-dot3(a, t)) / (t_norm * a_norm));
  const float beta_plus = beta + one_pix;
  const float gamma_plus = (float)(M_PI - alpha - beta_plus); /* :65, double */
  const float z_plus = t_norm * sinf(beta_plus) / sinf(gamma_plus);
  return z_plus - z;
}

/* src/seed_update.cu:31-37; 2.0f*M_PI*sigma_sq is a double product narrowed
 * to float at the rsqrtf call. */
static float normpdf(float x, float mu, float sigma_sq) {
  return (expf(-(x - mu) * (x - mu) / (2.0f * sigma_sq))) *
         rsqrt_host((float)(2.0f * M_PI * sigma_sq));
}

/* src/seed_update.cu:40-121 */
void rmd_oracle_stage_update(rmd_oracle_seeds *s, const float *T_ref_curr) {
  const int w = s->w, h = s->h;
  const float opa = one_pix_angle(s);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      const size_t k = (size_t)y * w + x;
      const int state = s->convergence[k];
      if (state == RMD_O_CONVERGED || state == RMD_O_DIVERGED) continue; /* :54-56 */
      if (state == RMD_O_UPDATE) {
        const float mu = s->mu[k], sigma_sq = s->sigma_sq[k];
        const float a = s->a[k], b = s->b[k];
        const f2 px_ref = {(float)x, (float)y};
        const f3 f_ref = normalize3(cam2world(s, px_ref));
        const f3 f_match = normalize3(cam2world(s, s->matches[k]));
        const f3 pt = triangulate_non_lin(f_ref, f_match, T_ref_curr);
        if (pt.z < 0.0f) continue; /* :77-80 */
        const float depth = norm3(pt);
        const float tau = triangulation_uncertainty(
            depth, f_ref, se3_translation(T_ref_curr), opa);
        const float tau_sq = tau * tau;
        const float s_sq = (tau_sq * sigma_sq) / (tau_sq + sigma_sq);
        const float m = s_sq * (mu / sigma_sq + depth / tau_sq);
        float c1 = (a / (a + b)) * normpdf(depth, mu, sigma_sq + tau_sq);
        float c2 = (b / (a + b)) * (1.0f / s->depth_range);
        const float norm_const = c1 + c2;
        c1 = c1 / norm_const;
        c2 = c2 / norm_const;
        const float f = c1 * ((a + 1.0f) / (a + b + 1.0f)) + c2 * (a / (a + b + 1.0f));
        const float e =
            c1 * (((a + 1.0f) * (a + 2.0f)) / ((a + b + 1.0f) * (a + b + 2.0f))) +
            c2 * (a * (a + 1.0f) / ((a + b + 1.0f) * (a + b + 2.0f)));
        if (isnan(c1 * m)) continue; /* :100-103 */
        const float mu_prime = c1 * m + c2 * mu;
        s->sigma_sq[k] = c1 * (s_sq + m * m) + c2 * (sigma_sq + mu * mu) -
                         mu_prime * mu_prime;
        s->mu[k] = mu_prime;
        const float a_prime = (e - f) / (f - e / f);
        s->a[k] = a_prime;
        s->b[k] = a_prime * (1.0f - f) / f;
      } else if (state == RMD_O_NO_MATCH) { /* :113-117 */
        s->b[k] = s->b[k] + 1.0f;
      }
    }
  }
}

/* src/seed_matrix.cu:120-158 (host driver) */
int rmd_oracle_seeds_update(rmd_oracle_seeds *s, const float *img,
                            const float *T_curr_world) {
  float T_ref_curr[12];
  rmd_oracle_se3_mul(T_curr_world, s->T_world_ref, s->T_curr_ref); /* :124 */
  s->dist_from_ref = norm3(se3_translation(s->T_curr_ref));          /* :125 */
  rmd_oracle_stage_check(s);                                         /* :139 */
  rmd_oracle_stage_match(s, img, s->T_curr_ref);                     /* :149 */
  rmd_oracle_se3_inv(s->T_curr_ref, T_ref_curr);
  rmd_oracle_stage_update(s, T_ref_curr);                            /* :155 */
  return 1;
}

/* src/seed_matrix.cu:195-198 */
size_t rmd_oracle_seeds_converged_count(const rmd_oracle_seeds *s) {
  return rmd_oracle_count_equal_i32(s->convergence, (size_t)s->w, (size_t)s->w,
                                    (size_t)s->h, RMD_O_CONVERGED);
}
float rmd_oracle_seeds_dist_from_ref(const rmd_oracle_seeds *s) {
  return s->dist_from_ref;
}
void rmd_oracle_seeds_T_curr_ref(const rmd_oracle_seeds *s, float *out12) {
  memcpy(out12, s->T_curr_ref, sizeof s->T_curr_ref);
}

/* ----------------------------------------------------------- TV-L1 denoise */

/*
 * src/depthmap_denoiser.cu:46-59 (weights), :62-118 (primal-dual step),
 * :124-141 (constants), :179-224 (driver), :226-229 (large sigma).
 *
 * The reference kernel updates p, u and u_head in place with only an
 * intra-block barrier, so values read across a 16x16 tile seam may belong to
 * iteration n or n+1 depending on block scheduling.
 * Inside a tile it is a Jacobi sweep: every dual p^{n+1} is computed from
 * (u^n, u_head^n, p^n), then every primal from p^{n+1}.  This oracle applies
 * that Jacobi semantics to the whole image, which is the deterministic limit
 * of the reference ("all blocks run in lock-step").
 */
int rmd_oracle_denoise(const float *mu, const float *sigma_sq, const float *a,
                       const float *b, int w, int h, float depth_range,
                       float lambda, int iterations, float *out) {
  const size_t n = (size_t)w * h;
  const float L = sqrtf(8.0f);         /* :130 */
  const float tau = 0.02f;             /* :131 */
  const float sigma = (1 / (L * L)) / tau; /* :132 */
  const float theta = 0.5f;            /* :133 */
  const float large_sigma_sq = depth_range * depth_range / 72.0f; /* :228 */
  float *g = (float *)malloc(n * sizeof(float));
  float *u = (float *)malloc(n * sizeof(float));
  float *uh = (float *)malloc(n * sizeof(float));
  f2 *p = (f2 *)calloc(n, sizeof(f2)); /* p_.zero(), :217 */
  if (!g || !u || !uh || !p) { free(g); free(u); free(uh); free(p); return 0; }

#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < n; ++k) { /* computeWeightsKernel :54-57 */
    const float E_pi = a[k] / (a[k] + b[k]);
    const float v = (E_pi * sigma_sq[k] + (1.0f - E_pi) * large_sigma_sq) / large_sigma_sq;
    g[k] = v > 1.0f ? v : 1.0f;
    u[k] = mu[k];  /* u_ = mu; u_head_ = u_  :215-216 */
    uh[k] = mu[k];
  }

  for (int it = 0; it < iterations; ++it) {
    /* dual ascent + projection, :70-83 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const size_t k = (size_t)y * w + x;
        const float cur_u = u[k];
        const int xe = x + 1 < w - 1 ? x + 1 : w - 1;
        const int ys = y + 1 < h - 1 ? y + 1 : h - 1;
        const float gx = uh[(size_t)y * w + xe] - cur_u;
        const float gy = uh[(size_t)ys * w + x] - cur_u;
        const float tx = g[k] * gx * sigma + p[k].x;
        const float ty = g[k] * gy * sigma + p[k].y;
        const float sq = sqrtf(tx * tx + ty * ty);
        const float d = 1.0f > sq ? 1.0f : sq;
        p[k].x = tx / d;
        p[k].y = ty / d;
      }
    /* divergence + primal shrink + over-relaxation, :86-115 */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const size_t k = (size_t)y * w + x;
        f2 cp = p[k];
        f2 wp = p[(size_t)y * w + (x - 1 > 0 ? x - 1 : 0)];
        f2 np = p[(size_t)(y - 1 > 0 ? y - 1 : 0) * w + x];
        if (x == 0) wp.x = 0.0f;
        else if (x >= w - 1) cp.x = 0.0f;
        if (y == 0) np.y = 0.0f;
        else if (y >= h - 1) cp.y = 0.0f;
        const float divergence = cp.x - wp.x + cp.y - np.y;
        const float old_u = u[k];
        const float noisy = mu[k];
        const float temp_u = old_u + tau * g[k] * divergence;
        float nu;
        if ((temp_u - noisy) > (tau * lambda)) nu = temp_u - tau * lambda;
        else if ((temp_u - noisy) < (-tau * lambda)) nu = temp_u + tau * lambda;
        else nu = noisy;
        u[k] = nu;
        uh[k] = nu + theta * (nu - old_u);
      }
  }
  memcpy(out, u, n * sizeof(float)); /* u_.getDevData :223 */
  free(g); free(u); free(uh); free(p);
  return 1;
}

/* -------------------------------------------------------------- reductions */

/* src/reduction_kernels.cu:59-105 run as in src/reduction.cu:88-104 with the
 * launch shape every caller uses (4x4 blocks of 16x16, src/seed_matrix.cu:72-79):
 * grid-stride per-thread partial sums, shared-memory tree per block, then one
 * block reducing the 4x4 partials the same way. */
#define RB 16
#define RG 4
static float tree256(float *sp) {
  for (int act = (RB * RB) >> 1; act; act >>= 1)
    for (int t = 0; t < act; ++t) sp[t] += sp[t + act];
  return sp[0];
}
float rmd_oracle_sum_f32_ref_order(const float *img, size_t stride, size_t w,
                                   size_t h) {
  float partial[RG * RG];
  float sp[RB * RB];
  for (int by = 0; by < RG; ++by)
    for (int bx = 0; bx < RG; ++bx) {
      for (int ty = 0; ty < RB; ++ty)
        for (int tx = 0; tx < RB; ++tx) {
          float sum = 0;
          for (size_t x = (size_t)bx * RB + tx; x < w; x += RB * RG)
            for (size_t y = (size_t)by * RB + ty; y < h; y += RB * RG)
              sum += img[y * stride + x];
          sp[ty * RB + tx] = sum;
        }
      partial[by * RG + bx] = tree256(sp);
    }
  /* second launch: n = grid.x, m = grid.y over the partials (stride = RG) */
  for (int ty = 0; ty < RB; ++ty)
    for (int tx = 0; tx < RB; ++tx) {
      float sum = 0;
      for (int x = tx; x < RG; x += RB)
        for (int y = ty; y < RG; y += RB) sum += partial[y * RG + x];
      sp[ty * RB + tx] = sum;
    }
  return tree256(sp);
}

double rmd_oracle_sum_f32_f64(const float *img, size_t stride, size_t w,
                              size_t h) {
  double s = 0.0;
  for (size_t y = 0; y < h; ++y)
    for (size_t x = 0; x < w; ++x) s += (double)img[y * stride + x];
  return s;
}

int rmd_oracle_sum_i32(const int *img, size_t stride, size_t w, size_t h) {
  int s = 0; /* integer addition is associative: order is irrelevant */
  for (size_t y = 0; y < h; ++y)
    for (size_t x = 0; x < w; ++x) s += img[y * stride + x];
  return s;
}

/* src/reduction_kernels.cu:109-159 */
size_t rmd_oracle_count_equal_i32(const int *img, size_t stride, size_t w,
                                  size_t h, int value) {
  size_t c = 0;
  for (size_t y = 0; y < h; ++y)
    for (size_t x = 0; x < w; ++x) c += (img[y * stride + x] == value);
  return c;
}
