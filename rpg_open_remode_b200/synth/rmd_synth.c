/*
 * rmd_synth.c -- deterministic synthetic sequences for the depth filter.
 *
 * The reference's test data (first_200_frames_traj_over_table, read by
 * test/dataset.cpp:81-186) is not redistributable with this repo and there
 * is no network, so tests and bench.py render a stand-in with the same
 * structure: a textured height field ("objects on a table") seen by a pinhole
 * camera with the data set's intrinsics (test/dataset_main.cpp:37, scaled to
 * the image size), a smooth hand-held-like trajectory of ~2.3 cm / frame
 * (paper Table I: 0.686 m/s at 30 fps) always looking at the scene centre,
 * exact poses, exact per-pixel ground-truth depth (distance along the viewing
 * ray, which is what the filter estimates: src/seed_update.cu:81), and 8-bit
 * gray images (the node feeds MONO8, src/depthmap.cpp:105 scales by 1/255).
 *
 * Host code, plain C + OpenMP.  No dependency on the CUDA library or oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_BUMPS 10

typedef struct {
  float cx, cy, r2inv, h;
} bump_t;

typedef struct {
  int width, height;
  float fx, fy, cx, cy;
  uint32_t seed;
  float tex_cell;    /* metres, finest texture lattice */
  bump_t bumps[N_BUMPS];
  float box_x0, box_x1, box_y0, box_y1, box_h, box_edge;
  float h_max;
  /* trajectory */
  double c0[3], amp[3], omega[3], phase[3], target[3];
} scene_t;

/* ----------------------------------------------------------------- random */
static uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
static float rnd01(uint32_t *state) {
  *state = hash_u32(*state + 0x9e3779b9U);
  return (float)(*state >> 8) * (1.0f / 16777216.0f);
}

/* ------------------------------------------------------------------ scene */
void *rmd_synth_create(int width, int height, float fx, float fy, float cx,
                       float cy, uint32_t seed) {
  scene_t *s = (scene_t *)calloc(1, sizeof *s);
  if (!s) return NULL;
  s->width = width; s->height = height;
  s->fx = fx; s->fy = fy; s->cx = cx; s->cy = cy;
  s->seed = seed;
  s->tex_cell = 5.0f * 1.5f / fabsf(fx); /* ~5 px at 1.5 m */
  uint32_t st = seed ^ 0xA5A5F00DU;
  float hsum = 0.0f;
  for (int i = 0; i < N_BUMPS; ++i) {
    bump_t *b = &s->bumps[i];
    b->cx = -0.9f + 1.8f * rnd01(&st);
    b->cy = -0.9f + 1.8f * rnd01(&st);
    const float r = 0.18f + 0.30f * rnd01(&st);
    b->r2inv = 1.0f / (r * r);
    b->h = 0.05f + 0.16f * rnd01(&st);
    hsum += b->h;
  }
  s->box_x0 = -0.55f + 0.3f * rnd01(&st);
  s->box_x1 = s->box_x0 + 0.35f + 0.2f * rnd01(&st);
  s->box_y0 = -0.45f + 0.3f * rnd01(&st);
  s->box_y1 = s->box_y0 + 0.30f + 0.2f * rnd01(&st);
  s->box_h = 0.12f + 0.10f * rnd01(&st);
  s->box_edge = 0.02f;
  s->h_max = hsum + s->box_h + 1e-3f;

  /* Trajectory: Lissajous wiggle around c0, always looking at `target`. */
  s->target[0] = 0.0; s->target[1] = 0.0; s->target[2] = 0.1;
  s->c0[0] = 0.15 - 0.3 * rnd01(&st);
  s->c0[1] = -1.05 - 0.2 * rnd01(&st);
  s->c0[2] = 1.45 + 0.2 * rnd01(&st);
  s->amp[0] = 0.50; s->amp[1] = 0.22; s->amp[2] = 0.12;
  s->omega[0] = 0.060; s->omega[1] = 0.047; s->omega[2] = 0.083;
  s->phase[0] = 0.0; s->phase[1] = 0.9; s->phase[2] = 0.3;
  return s;
}

void rmd_synth_destroy(void *p) { free(p); }

static inline float smooth01(float t) {
  t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
  return t * t * (3.0f - 2.0f * t);
}

static inline float height_at(const scene_t *s, float x, float y) {
  float h = 0.0f;
  for (int i = 0; i < N_BUMPS; ++i) {
    const bump_t *b = &s->bumps[i];
    const float dx = x - b->cx, dy = y - b->cy;
    const float q = 1.0f - (dx * dx + dy * dy) * b->r2inv;
    if (q > 0.0f) h += b->h * q * q * q;
  }
  /* a box with steep (2 cm) sides: depth discontinuities and occlusions */
  const float e = s->box_edge;
  const float wx = smooth01((x - s->box_x0) / e) * smooth01((s->box_x1 - x) / e);
  const float wy = smooth01((y - s->box_y0) / e) * smooth01((s->box_y1 - y) / e);
  h += s->box_h * wx * wy;
  return h;
}

static inline float lattice(uint32_t seed, int ix, int iy) {
  const uint32_t hsh =
      hash_u32((uint32_t)ix * 0x8da6b343U ^ (uint32_t)iy * 0xd8163841U ^ seed);
  return (float)(hsh >> 8) * (2.0f / 16777216.0f) - 1.0f;
}

static inline float vnoise(uint32_t seed, float x, float y) {
  const float fx = floorf(x), fy = floorf(y);
  const int ix = (int)fx, iy = (int)fy;
  float tx = x - fx, ty = y - fy;
  tx = tx * tx * tx * (tx * (tx * 6.0f - 15.0f) + 10.0f);
  ty = ty * ty * ty * (ty * (ty * 6.0f - 15.0f) + 10.0f);
  const float v00 = lattice(seed, ix, iy), v10 = lattice(seed, ix + 1, iy);
  const float v01 = lattice(seed, ix, iy + 1), v11 = lattice(seed, ix + 1, iy + 1);
  const float a = v00 + tx * (v10 - v00), b = v01 + tx * (v11 - v01);
  return a + ty * (b - a);
}

static inline float albedo_at(const scene_t *s, float x, float y) {
  const float c = 1.0f / s->tex_cell;
  float v = 0.5f;
  v += 0.20f * vnoise(s->seed + 1, x * c, y * c);
  v += 0.16f * vnoise(s->seed + 2, x * c * 0.37f + 11.3f, y * c * 0.37f - 4.1f);
  v += 0.12f * vnoise(s->seed + 3, x * c * 0.11f - 7.7f, y * c * 0.11f + 2.9f);
  v += 0.05f * sinf(9.0f * x + 1.3f) * sinf(7.0f * y - 0.4f);
  return v < 0.02f ? 0.02f : (v > 0.98f ? 0.98f : v);
}

/* ------------------------------------------------------------- trajectory */

/* T_world_cam (3x4 row major, camera -> world), float, for frame k.
 * Camera axes: x right, y down, z forward (right-handed).  With the data
 * set's negative fy the rendered image is simply upside-down, which the
 * filter does not care about; what matters is that poses, projection and
 * images are mutually consistent. */
void rmd_synth_pose(const void *p, int k, float *T_world_cam) {
  const scene_t *s = (const scene_t *)p;
  double c[3];
  for (int i = 0; i < 3; ++i)
    c[i] = s->c0[i] + s->amp[i] * (sin(s->omega[i] * k + s->phase[i]) - sin(s->phase[i]));
  double z[3] = {s->target[0] - c[0], s->target[1] - c[1], s->target[2] - c[2]};
  double n = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
  for (int i = 0; i < 3; ++i) z[i] /= n;
  /* x = z cross world-up(0,0,1) -> "right" for a downward-looking camera is
   * (z x up); y = z cross x  (so that x cross y = z). */
  double x[3] = {z[1] * 1.0 - z[2] * 0.0, z[2] * 0.0 - z[0] * 1.0, 0.0};
  n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int i = 0; i < 3; ++i) x[i] /= n;
  double y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2],
                 z[0] * x[1] - z[1] * x[0]};
  for (int r = 0; r < 3; ++r) {
    T_world_cam[4 * r + 0] = (float)x[r];
    T_world_cam[4 * r + 1] = (float)y[r];
    T_world_cam[4 * r + 2] = (float)z[r];
    T_world_cam[4 * r + 3] = (float)c[r];
  }
}

/* ----------------------------------------------------------------- render */

/* Renders one view.  Any output pointer may be NULL.
 * img_u8: gray 0..255; img_f32: the same value * (1/255.f) as
 * cv::Mat::convertTo(CV_32F, 1/255.f) produces (src/depthmap.cpp:105);
 * depth: distance from the camera centre to the surface point (metres). */
int rmd_synth_render(const void *p, const float *T_world_cam, uint8_t *img_u8,
                     float *img_f32, float *depth) {
  const scene_t *s = (const scene_t *)p;
  const int w = s->width, h = s->height;
  const float *T = T_world_cam;
  const float ox = T[3], oy = T[7], oz = T[11];
  const int n_march = 48;
#pragma omp parallel for schedule(dynamic, 8)
  for (int v = 0; v < h; ++v) {
    for (int u = 0; u < w; ++u) {
      /* viewing ray through the pixel centre convention of the reference:
       * integer (u, v) is the sample position (pinhole_camera.cuh:40-46). */
      const float xc = ((float)u - s->cx) / s->fx;
      const float yc = ((float)v - s->cy) / s->fy;
      const float dx = T[0] * xc + T[1] * yc + T[2];
      const float dy = T[4] * xc + T[5] * yc + T[6];
      const float dz = T[8] * xc + T[9] * yc + T[10];
      float t_hit;
      if (dz >= -1e-6f) {
        t_hit = 10.0f; /* a ray that does not descend: never happens on the generated trajectories */
      } else {
        /* h_max (the SUM of all bump heights + box) only bounds the height field; a camera below
         * that bound is still above the terrain, so its rays start at the camera itself.  (Round 1
         * sent every ray of such a camera to t = 10: keyframe seeds 1, 3, 4, 5 and the 1080p seed
         * rendered a view-dependent shell instead of the scene, and nothing could converge.) */
        const float t_top = oz > s->h_max ? (s->h_max - oz) / dz : 0.0f;
        const float t_bot = (0.0f - oz) / dz;
        const float dt = (t_bot - t_top) / (float)n_march;
        float t_lo = t_top, t_hi = t_bot;
        float t_prev = t_top;
        for (int i = 1; i <= n_march; ++i) {
          const float t = t_top + dt * (float)i;
          const float f = oz + t * dz - height_at(s, ox + t * dx, oy + t * dy);
          if (f <= 0.0f) { t_lo = t_prev; t_hi = t; break; }
          t_prev = t;
        }
        for (int i = 0; i < 18; ++i) {
          const float tm = 0.5f * (t_lo + t_hi);
          const float f = oz + tm * dz - height_at(s, ox + tm * dx, oy + tm * dy);
          if (f <= 0.0f) t_hi = tm; else t_lo = tm;
        }
        t_hit = 0.5f * (t_lo + t_hi);
      }
      const float px = ox + t_hit * dx, py = oy + t_hit * dy;
      const float alb = albedo_at(s, px, py);
      int q = (int)lrintf(alb * 255.0f);
      q = q < 0 ? 0 : (q > 255 ? 255 : q);
      const size_t k = (size_t)v * w + u;
      if (img_u8) img_u8[k] = (uint8_t)q;
      if (img_f32) img_f32[k] = (float)q * (1.0f / 255.0f);
      if (depth) depth[k] = t_hit * sqrtf(xc * xc + yc * yc + 1.0f);
    }
  }
  return 0;
}
