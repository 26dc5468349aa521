// tex_probe.cu -- measures how the texture unit quantises bilinear weights
// (the reference samples the current frame through cudaFilterModeLinear,
// include/rmd/texture_memory.cuh:48, src/epipolar_match.cu:111-114).
// Build + run on the GPU box:  nvcc -arch=sm_100a -o /tmp/tex_probe tools/tex_probe.cu
// Output feeds the texture model of oracle/rmd_oracle.c (tex_linear) and of
// the kernels (tap_frame), see DESIGN.md "bilinear parity".
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while(0)

__global__ void sample(cudaTextureObject_t tex, const float2 *xy, float *out, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n) out[i] = tex2D<float>(tex, xy[i].x, xy[i].y);
}

int main()
{
  const int W = 1024, H = 16;
  std::vector<float> img(W * H, 0.0f);
  float *d_img; size_t pitch;
  CK(cudaMallocPitch(&d_img, &pitch, W * sizeof(float), H));

  auto make_tex = [&]() {
    CK(cudaMemcpy2D(d_img, pitch, img.data(), W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice));
    cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d_img; rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = H; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td; memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t t; CK(cudaCreateTextureObject(&t, &rd, &td, NULL));
    return t;
  };
  auto run = [&](cudaTextureObject_t t, const std::vector<float2> &xy) {
    float2 *d_xy; float *d_out; const int n = (int)xy.size();
    CK(cudaMalloc(&d_xy, n * sizeof(float2))); CK(cudaMalloc(&d_out, n * sizeof(float)));
    CK(cudaMemcpy(d_xy, xy.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
    sample<<<(n + 127) / 128, 128>>>(t, d_xy, d_out, n);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(n);
    CK(cudaMemcpy(out.data(), d_out, n * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_xy); cudaFree(d_out);
    return out;
  };

  // ---- experiment 1: alpha staircase between texel bx and bx+1 (values 0 and 1)
  const int bases[3] = {3, 300, 1000};
  for(int b = 0; b < 3; ++b)
  {
    const int bx = bases[b];
    std::fill(img.begin(), img.end(), 0.0f);
    for(int y = 0; y < H; ++y) img[y * W + bx + 1] = 1.0f;
    cudaTextureObject_t t = make_tex();
    const int N = 4096;
    std::vector<float2> xy(N + 1);
    for(int k = 0; k <= N; ++k) xy[k] = make_float2((float)bx + 0.5f + (float)k / (float)N, 5.5f);
    std::vector<float> out = run(t, xy);
    // summarise: distinct output levels and the k at which each level starts
    printf("EXP1 base=%d  levels(k_start:value*256):", bx);
    float prev = -1.0f; int levels = 0;
    for(int k = 0; k <= N; ++k)
      if(out[k] != prev) { if(levels < 12 || levels > 250) printf(" %d:%.4f", k, out[k] * 256.0f); prev = out[k]; ++levels; }
    printf("\nEXP1 base=%d  n_levels=%d\n", bx, levels);
    // which rounding rule? compare to round-half-up, floor, round-half-even of frac*256
    int mism_round = 0, mism_floor = 0;
    for(int k = 0; k <= N; ++k)
    {
      const float xb = xy[k].x - 0.5f;
      const float fr = xb - floorf(xb);
      const float r = floorf(fr * 256.0f + 0.5f) / 256.0f, f = floorf(fr * 256.0f) / 256.0f;
      const float got = (floorf(xb) > (float)bx) ? 1.0f + out[k] : out[k];  // k==N lands on next texel
      if(got != r && !(floorf(xb) > (float)bx)) ++mism_round;
      if(got != f && !(floorf(xb) > (float)bx)) ++mism_floor;
    }
    printf("EXP1 base=%d  mismatches vs round-half-up=%d  vs floor=%d (of %d)\n", bx, mism_round, mism_floor, N + 1);
    CK(cudaDestroyTextureObject(t));
  }

  // ---- experiment 1y: the same staircase along y (texel rows by and by+1 hold 0 and 1)
  {
    const int by = 7;
    std::fill(img.begin(), img.end(), 0.0f);
    for(int x = 0; x < W; ++x) img[(by + 1) * W + x] = 1.0f;
    cudaTextureObject_t t = make_tex();
    const int N = 4096;
    std::vector<float2> xy(N + 1);
    for(int k = 0; k <= N; ++k) xy[k] = make_float2(10.5f, (float)by + 0.5f + (float)k / (float)N);
    std::vector<float> out = run(t, xy);
    int mism_round = 0, mism_floor = 0, levels = 0; float prev = -1.0f;
    for(int k = 0; k < N; ++k)
    {
      const float fr = (float)k / (float)N;
      mism_round += (out[k] != floorf(fr * 256.0f + 0.5f) / 256.0f);
      mism_floor += (out[k] != floorf(fr * 256.0f) / 256.0f);
      if(out[k] != prev) { prev = out[k]; ++levels; }
    }
    printf("EXP1y y-staircase: n_levels=%d mismatches vs round-half-up=%d vs floor=%d (of %d)\n", levels, mism_round, mism_floor, N);
    CK(cudaDestroyTextureObject(t));
  }

  // ---- experiment 2: value arithmetic. texels t0=0.3, t1=0.9 in x; check result vs formulas
  {
    std::fill(img.begin(), img.end(), 0.0f);
    for(int y = 0; y < H; ++y) { img[y * W + 10] = 0.3f; img[y * W + 11] = 0.9f; }
    cudaTextureObject_t t = make_tex();
    std::vector<float2> xy;
    for(int k = 0; k < 256; ++k) xy.push_back(make_float2(10.5f + (float)k / 256.0f, 5.5f));
    std::vector<float> out = run(t, xy);
    int eq_lerp = 0, eq_w = 0, eq_fma = 0; double maxd = 0;
    for(int k = 0; k < 256; ++k)
    {
      const float a = (float)k / 256.0f, t0 = 0.3f, t1 = 0.9f;
      const float lerp = t0 + a * (t1 - t0);
      const float wsum = (1.0f - a) * t0 + a * t1;
      const float f = fmaf(a, t1, (1.0f - a) * t0);
      eq_lerp += (out[k] == lerp); eq_w += (out[k] == wsum); eq_fma += (out[k] == f);
      maxd = fmax(maxd, fabs((double)out[k] - ((1.0 - a) * (double)t0 + a * (double)t1)));
    }
    printf("EXP2 1-D: equal to t0+a*(t1-t0): %d/256, (1-a)*t0+a*t1: %d/256, fma form: %d/256, max |err| vs exact=%.3g\n",
           eq_lerp, eq_w, eq_fma, maxd);
    CK(cudaDestroyTextureObject(t));
  }

  // ---- experiment 3: 2-D product weights, random texels
  {
    srand(7);
    for(size_t i = 0; i < img.size(); ++i) img[i] = (float)(rand() % 256) / 255.0f;
    cudaTextureObject_t t = make_tex();
    std::vector<float2> xy;
    for(int k = 0; k < 4096; ++k)
      xy.push_back(make_float2(20.0f + 900.0f * (float)rand() / (float)RAND_MAX, 2.0f + 10.0f * (float)rand() / (float)RAND_MAX));
    std::vector<float> out = run(t, xy);
    double max_q = 0, max_exact = 0; int eq4 = 0, eqsep = 0;
    for(size_t k = 0; k < xy.size(); ++k)
    {
      const float xb = xy[k].x - 0.5f, yb = xy[k].y - 0.5f;
      // quantised model
      const float tx = floorf(xb * 256.0f + 0.5f), ty = floorf(yb * 256.0f + 0.5f);
      const int i = (int)floorf(tx / 256.0f), j = (int)floorf(ty / 256.0f);
      const float a = (tx - i * 256.0f) / 256.0f, b = (ty - j * 256.0f) / 256.0f;
      const float t00 = img[j * W + i], t10 = img[j * W + i + 1], t01 = img[(j + 1) * W + i], t11 = img[(j + 1) * W + i + 1];
      const float four = (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
      const float sep = (1 - b) * ((1 - a) * t00 + a * t10) + b * ((1 - a) * t01 + a * t11);
      const double q = (1.0 - a) * (1.0 - b) * t00 + (double)a * (1.0 - b) * t10 + (1.0 - a) * (double)b * t01 + (double)a * b * t11;
      // exact-weight model
      const int ie = (int)floorf(xb), je = (int)floorf(yb);
      const double ae = xb - ie, be = yb - je;
      const double ex = (1 - ae) * (1 - be) * img[je * W + ie] + ae * (1 - be) * img[je * W + ie + 1] + (1 - ae) * be * img[(je + 1) * W + ie] + ae * be * img[(je + 1) * W + ie + 1];
      max_q = fmax(max_q, fabs(out[k] - q)); max_exact = fmax(max_exact, fabs(out[k] - ex));
      eq4 += (out[k] == four); eqsep += (out[k] == sep);
    }
    printf("EXP3 2-D: max|hw - quantised-weight model|=%.3g  max|hw - exact-weight model|=%.3g  bit-equal 4-term=%d sep=%d of %zu\n",
           max_q, max_exact, eq4, eqsep, xy.size());
    // error distribution of the quantised model and where the large errors sit
    int n_small = 0, n_mid = 0, n_big = 0;
    for(size_t k = 0; k < xy.size(); ++k)
    {
      const float xb = xy[k].x - 0.5f, yb = xy[k].y - 0.5f;
      const float tx = floorf(xb * 256.0f + 0.5f), ty = floorf(yb * 256.0f + 0.5f);
      const int i = (int)floorf(tx / 256.0f), j = (int)floorf(ty / 256.0f);
      const float a = (tx - i * 256.0f) / 256.0f, b = (ty - j * 256.0f) / 256.0f;
      const double q = (1.0 - a) * (1.0 - b) * img[j * W + i] + (double)a * (1.0 - b) * img[j * W + i + 1] +
                       (1.0 - a) * (double)b * img[(j + 1) * W + i] + (double)a * b * img[(j + 1) * W + i + 1];
      const double e = fabs(out[k] - q);
      if(e < 1e-6) ++n_small; else if(e < 1e-4) ++n_mid; else
      {
        if(n_big < 12)
          printf("EXP3 big err %.3g at x=%.6f y=%.6f  frac_x*256=%.4f frac_y*256=%.4f  hw=%.6f model=%.6f\n", e,
                 xy[k].x, xy[k].y, (xb - floorf(xb)) * 256.0f, (yb - floorf(yb)) * 256.0f, out[k], q);
        ++n_big;
      }
    }
    printf("EXP3 quantised-model error histogram: <1e-6: %d  <1e-4: %d  >=1e-4: %d\n", n_small, n_mid, n_big);
    // ---- experiment 4: NaN / inf coordinates
    std::vector<float2> bad;
    bad.push_back(make_float2(nanf(""), nanf("")));
    bad.push_back(make_float2(nanf(""), 5.5f));
    bad.push_back(make_float2(INFINITY, 5.5f));
    bad.push_back(make_float2(-INFINITY, 5.5f));
    std::vector<float> o2 = run(t, bad);
    printf("EXP4 NaN/inf coords: tex(NaN,NaN)=%g [img(0,0)=%g]  tex(NaN,5.5)=%g [img(0,5)=%g]  tex(+inf,5.5)=%g [img(W-1,5)=%g]  tex(-inf,5.5)=%g\n",
           o2[0], img[0], o2[1], img[5 * W], o2[2], img[5 * W + W - 1], o2[3]);
    CK(cudaDestroyTextureObject(t));
  }
  return 0;
}
