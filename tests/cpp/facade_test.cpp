// facade_test.cpp -- the reference's gtests re-hosted against the
// header-compatible C++ classes in include/rmd/ (which forward to the C-ABI).
// No gtest / OpenCV in this image, so plain checks; frames come from the
// synthetic generator (the reference's data set is external).
//
//   seedMatrixInit      test/seed_matrix_test.cpp:29-151
//   seedMatrixCheck     test/seed_matrix_test.cpp:154-243
//   epipolarMatchTest   test/epipolar_test.cpp:138-225
//   reduction sum/count test/reduction_test.cpp:24-122
//   device image        test/device_image_test.cpp:27-115
//   Depthmap sequence   src/depthmap.cpp:63-122 call order (OpenCV-free)
//
// Build (tests/test_cpp_facade.py does this):
//   g++ -std=c++14 -DRMD_BUILD_TESTS=1 -Iinclude -I/usr/local/cuda/include tests/cpp/facade_test.cpp \
//       -Lrpg_open_remode_b200 -lrmd_b200 -Lrpg_open_remode_b200/synth -lrmd_synth -L/usr/local/cuda/lib64 -lcudart
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <rmd/depthmap_denoiser.cuh>
#include <rmd/reduction.cuh>
#include <rmd/se3.cuh>
#include <rmd/seed_matrix.cuh>

extern "C"
{
void *rmd_synth_create(int width, int height, float fx, float fy, float cx, float cy, uint32_t seed);
void rmd_synth_destroy(void *p);
void rmd_synth_pose(const void *p, int k, float *T_world_cam);
int rmd_synth_render(const void *p, const float *T_world_cam, uint8_t *img_u8, float *img_f32, float *depth);
}

static int g_failures = 0;
#define CHECK(cond)                                                                  \
  do {                                                                               \
    if(!(cond)) { std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_failures; } \
  } while(0)

static bool float_eq_4ulp(float a, float b)  // ASSERT_FLOAT_EQ
{
  int32_t ia, ib;
  std::memcpy(&ia, &a, 4); std::memcpy(&ib, &b, 4);
  if(ia < 0) ia = INT32_MIN - ia;
  if(ib < 0) ib = INT32_MIN - ib;
  return std::llabs((long long)ia - (long long)ib) <= 4;
}

struct Frame
{
  std::vector<float> img, depth;
  rmd::SE3<float> T_world_cam;
};

static Frame render(void *scene, int w, int h, int k)
{
  Frame f;
  f.img.resize((size_t)w * h);
  f.depth.resize((size_t)w * h);
  float T[12];
  rmd_synth_pose(scene, k, T);
  rmd_synth_render(scene, T, NULL, f.img.data(), f.depth.data());
  float r[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  float t[3] = {T[3], T[7], T[11]};
  f.T_world_cam = rmd::SE3<float>(r, t);
  return f;
}

int main()
{
  const int W = 160, H = 120, P = RMD_CORR_PATCH_SIDE;
  const rmd::PinholeCamera cam(481.2f * W / 640.0f, -480.0f * H / 480.0f, (W - 1) / 2.0f, (H - 1) / 2.0f);
  void *scene = rmd_synth_create(W, H, cam.fx, cam.fy, cam.cx, cam.cy, 0x5EED0001u);
  const Frame f1 = render(scene, W, H, 1), f20 = render(scene, W, H, 20);
  const size_t n = (size_t)W * H;
  const float min_d = 0.4f, max_d = 1.8f;

  try
  {
    // ---- seedMatrixInit
    {
      rmd::SeedMatrix seeds(W, H, cam);
      CHECK(seeds.setReferenceImage(const_cast<float*>(f1.img.data()), f1.T_world_cam.inv(), min_d, max_d));
      std::vector<float> mu(n), s2(n), a(n), b(n), st(n), sd(n);
      seeds.downloadDepthmap(mu.data()); seeds.downloadSigmaSq(s2.data());
      seeds.downloadA(a.data()); seeds.downloadB(b.data());
      seeds.downloadSumTempl(st.data()); seeds.downloadConstTemplDenom(sd.data());
      const float avg = (min_d + max_d) / 2.0f, sig = (max_d - min_d) * (max_d - min_d) / 36.0f;
      bool ok = true;
      for(size_t i = 0; i < n; ++i)
        ok = ok && float_eq_4ulp(avg, mu[i]) && float_eq_4ulp(sig, s2[i]) && a[i] == 10.0f && b[i] == 10.0f;
      CHECK(ok);
      double worst_sum = 0, worst_den = 0;
      for(int y = P; y < H - P / 2; ++y)
        for(int x = P; x < W - P / 2; ++x)
        {
          double s = 0, ss = 0;
          for(int py = 0; py < P; ++py)
            for(int px = 0; px < P; ++px)
            {
              const double t = f1.img[(size_t)(y - P / 2 + py) * W + (x - P / 2 + px)];
              s += t; ss += t * t;
            }
          worst_sum = std::fmax(worst_sum, std::fabs((float)s - st[(size_t)y * W + x]));
          worst_den = std::fmax(worst_den, std::fabs((float)((double)(P * P) * ss - s * s) - sd[(size_t)y * W + x]));
        }
      CHECK(worst_sum <= 2e-5);
      CHECK(worst_den <= 1e-3);
      CHECK(seeds.getConvergedCount() == 0);
    }
    // ---- seedMatrixCheck: the reference image as current image with the pose of frame 20
    {
      rmd::SeedMatrix seeds(W, H, cam);
      seeds.setReferenceImage(const_cast<float*>(f1.img.data()), f1.T_world_cam.inv(), min_d, max_d);
      CHECK(seeds.update(const_cast<float*>(f1.img.data()), f20.T_world_cam.inv()));
      std::vector<int> conv(n);
      seeds.downloadConvergence(conv.data());
      bool ring_ok = true, inner_ok = true;
      for(int r = 0; r < H; ++r)
        for(int c = 0; c < W; ++c)
        {
          const int v = conv[(size_t)r * W + c];
          if(r > H - P - 1 || r < P || c > W - P - 1 || c < P)
            ring_ok = ring_ok && (v == rmd::ConvergenceStates::BORDER);
          else
            inner_ok = inner_ok && (v == rmd::ConvergenceStates::UPDATE || v == rmd::ConvergenceStates::DIVERGED ||
                                    v == rmd::ConvergenceStates::CONVERGED || v == rmd::ConvergenceStates::NOT_VISIBLE ||
                                    v == rmd::ConvergenceStates::NO_MATCH);
        }
      CHECK(ring_ok);
      CHECK(inner_ok);
      CHECK(seeds.getDistFromRef() > 0.0f);
      // getConvergence() + ImageReducer::countEqual == getConvergedCount()
      rmd::ImageReducer<int> reducer(dim3(16, 16), dim3(4, 4));
      CHECK(reducer.countEqual(seeds.getConvergence(), rmd::ConvergenceStates::CONVERGED) == seeds.getConvergedCount());
    }
    // ---- epipolarMatchTest: identity motion, UPDATE pixels match themselves
    {
      rmd::SeedMatrix seeds(W, H, cam);
      seeds.setReferenceImage(const_cast<float*>(f1.img.data()), f1.T_world_cam.inv(), min_d, max_d);
      seeds.update(const_cast<float*>(f1.img.data()), f1.T_world_cam.inv());
      std::vector<float2> m(n);
      std::vector<int> conv(n);
      seeds.downloadEpipolarMatches(m.data());
      seeds.downloadConvergence(conv.data());
      size_t upd = 0; bool ok = true;
      for(int r = 0; r < H; ++r)
        for(int c = 0; c < W; ++c)
          if(conv[(size_t)r * W + c] == rmd::ConvergenceStates::UPDATE)
          {
            ++upd;
            ok = ok && std::fabs(m[(size_t)r * W + c].x - (float)c) <= 0.01f &&
                 std::fabs(m[(size_t)r * W + c].y - (float)r) <= 0.01f;
          }
      CHECK(ok);
      CHECK(upd > (size_t)(W - 2 * P) * (H - 2 * P) / 2);
    }
    // ---- reductions
    {
      const size_t w = 752, h = 480;
      std::vector<float> img(w * h);
      std::vector<int> ints(w * h);
      uint32_t st = 12345u; double dsum = 0; size_t cnt = 0;
      for(size_t i = 0; i < w * h; ++i)
      {
        st = st * 1664525u + 1013904223u;
        img[i] = (float)(st >> 8) * (1.0f / 16777216.0f);
        dsum += img[i];
        ints[i] = (int)((st >> 12) & 255u);
        cnt += (ints[i] == 2);
      }
      rmd::DeviceImage<float> d_img(w, h);
      d_img.setDevData(img.data());
      rmd::ImageReducer<float> fr(dim3(16, 16), dim3(4, 4));
      CHECK(float_eq_4ulp((float)dsum, fr.sum(d_img)));
      rmd::DeviceImage<int> d_int(w, h);
      d_int.setDevData(ints.data());
      rmd::ImageReducer<int> ir(dim3(16, 16), dim3(4, 4));
      CHECK(ir.countEqual(d_int, 2) == cnt);
      // device image round trip + copy + zero
      std::vector<float> back(w * h);
      d_img.getDevData(back.data());
      CHECK(back == img);
      rmd::DeviceImage<float> d_copy(w, h);
      d_copy = d_img;
      d_copy.getDevData(back.data());
      CHECK(back == img);
      d_copy.zero();
      d_copy.getDevData(back.data());
      bool allzero = true;
      for(float v : back) allzero = allzero && v == 0.0f;
      CHECK(allzero);
      CHECK(d_img.stride * sizeof(float) == d_img.pitch && d_img.width == w && d_img.height == h);
    }
    // ---- rmd::Depthmap call order (src/depthmap.cpp:63-122) without OpenCV
    {
      const Frame f0 = render(scene, W, H, 0);
      float dmin = 1e9f, dmax = 0.f;
      for(float v : f0.depth) { dmin = std::fmin(dmin, v); dmax = std::fmax(dmax, v); }
      rmd::SeedMatrix seeds(W, H, cam);
      std::unique_ptr<rmd::DepthmapDenoiser> denoiser(new rmd::DepthmapDenoiser(W, H));
      std::vector<float> out(n, -1.0f);
      // denoise before setLargeSigmaSq: refused, output untouched (depthmap_denoiser.cu:189-193)
      seeds.setReferenceImage(const_cast<float*>(f0.img.data()), f0.T_world_cam.inv(), dmin, dmax);
      denoiser->denoise(seeds.getMu(), seeds.getSigmaSq(), seeds.getA(), seeds.getB(), out.data(), 0.5f, 5);
      CHECK(out[0] == -1.0f);
      denoiser->setLargeSigmaSq(dmax - dmin);
      for(int k = 1; k <= 30; ++k)
      {
        const Frame f = render(scene, W, H, k);
        seeds.update(const_cast<float*>(f.img.data()), f.T_world_cam.inv());
      }
      std::vector<float> mu(n);
      std::vector<int> conv(n);
      seeds.downloadDepthmap(mu.data());
      seeds.downloadConvergence(conv.data());
      denoiser->denoise(seeds.getMu(), seeds.getSigmaSq(), seeds.getA(), seeds.getB(), out.data(), 0.5f, 200);
      size_t nconv = 0; std::vector<float> errs;
      for(size_t i = 0; i < n; ++i)
        if(conv[i] == rmd::ConvergenceStates::CONVERGED) { ++nconv; errs.push_back(std::fabs(mu[i] - f0.depth[i])); }
      CHECK(nconv == seeds.getConvergedCount());
      CHECK(nconv > (size_t)(W - 2 * P) * (H - 2 * P) / 2);
      std::nth_element(errs.begin(), errs.begin() + errs.size() / 2, errs.end());
      CHECK(errs[errs.size() / 2] < 0.02f * (dmax - dmin));
      bool finite = true;
      for(float v : out) finite = finite && std::isfinite(v);
      CHECK(finite);
      const float pct = (float)seeds.getConvergedCount() / (float)(W * H) * 100.0f;  // depthmap.cpp:150-154
      std::printf("converged %.1f %% after 30 frames, median error %.4f m\n", pct, errs[errs.size() / 2]);
    }
    // ---- error convention: CUDA failures surface as rmd::CudaException
    {
      bool thrown = false;
      try
      {
        rmd::SeedMatrix seeds(W, H, cam);
        std::vector<float> img(n, 0.5f);
        seeds.update(img.data(), f1.T_world_cam.inv());  // no reference image yet
      }
      catch(const rmd::CudaException &e)
      {
        thrown = true;
        CHECK(std::string(e.what()).find("CudaException") != std::string::npos);
      }
      CHECK(thrown);
    }
  }
  catch(const rmd::CudaException &e)
  {
    std::printf("unexpected CudaException: %s\n", e.what());
    ++g_failures;
  }
  rmd_synth_destroy(scene);
  std::printf(g_failures ? "FAILED (%d)\n" : "ALL FACADE TESTS PASSED\n", g_failures);
  return g_failures ? 1 : 0;
}
