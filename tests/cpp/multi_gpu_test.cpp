// multi_gpu_test.cpp -- the C-ABI multi-GPU path from plain C++ (no Python, no torch): one host thread, one
// rmd_seeds_t per visible GPU (independent keyframes), rmd_multi_create (ncclCommInitAll) and the final
// gather of depth + convergence maps to rank 0, compared with each keyframe's own downloads.
// Built and run by tests/test_cpp_multi_gpu.py; uses every visible GPU (1 is enough: a 1-rank communicator).
#include <cuda_runtime.h>
#include <rmd_b200.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
void *rmd_synth_create(int w, int h, float fx, float fy, float cx, float cy, unsigned int seed);
void rmd_synth_destroy(void *p);
void rmd_synth_pose(const void *p, int k, float *T_world_cam);
int rmd_synth_render(const void *p, const float *T_world_cam, unsigned char *u8, float *f32, float *depth);
}

#define CHECK(expr) do { const int rc_ = (expr); if(rc_ != 0) { \
  std::fprintf(stderr, "FAILED %s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #expr, rc_, rmd_last_error_string()); \
  return 1; } } while(0)
#define REQUIRE(cond) do { if(!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while(0)

static void invert(const float *T, float *out)   // 3x4 [R|t] -> [R^T | -R^T t]
{
  for(int r = 0; r < 3; ++r)
  {
    for(int c = 0; c < 3; ++c) out[4 * r + c] = T[4 * c + r];
    out[4 * r + 3] = -(T[r] * T[3] + T[4 + r] * T[7] + T[8 + r] * T[11]);
  }
}

int main()
{
  int n_dev = 0;
  CHECK(rmd_device_count(&n_dev));
  REQUIRE(n_dev >= 1);
  const int n = n_dev > 8 ? 8 : n_dev;
  const int W = 320, H = 240, N_FRAMES = 12;
  const float fx = 481.2f * W / 640.0f, fy = -480.0f * H / 480.0f, cx = (W - 1) / 2.0f, cy = (H - 1) / 2.0f;
  std::vector<int> devices(n);
  std::vector<rmd_seeds_t*> seeds(n, (rmd_seeds_t*)NULL);
  std::vector<float> img((size_t)W * H), depth((size_t)W * H);
  for(int r = 0; r < n; ++r)
  {
    devices[r] = r;
    CHECK(rmd_seeds_create(W, H, fx, fy, cx, cy, 5, r, &seeds[r]));
    void *scene = rmd_synth_create(W, H, fx, fy, cx, cy, 0x5EED0002u + 16u * (unsigned)r);   // keyframe r
    REQUIRE(scene != NULL);
    for(int k = 0; k < N_FRAMES; ++k)
    {
      float T_world_cam[12], T_cam_world[12];
      rmd_synth_pose(scene, k, T_world_cam);
      invert(T_world_cam, T_cam_world);
      rmd_synth_render(scene, T_world_cam, NULL, img.data(), k == 0 ? depth.data() : NULL);
      if(k == 0)
      {
        float dmin = 1e30f, dmax = -1e30f;
        for(float d : depth) { dmin = d < dmin ? d : dmin; dmax = d > dmax ? d : dmax; }
        CHECK(rmd_seeds_set_reference(seeds[r], img.data(), T_cam_world, dmin, dmax));
      }
      else
        CHECK(rmd_seeds_update(seeds[r], img.data(), T_cam_world));   // asynchronous: the gather waits on the device
    }
    rmd_synth_destroy(scene);
  }
  rmd_multi_t *multi = NULL;
  CHECK(rmd_multi_create(devices.data(), n, W, H, &multi));
  int n_ranks = 0, n_local = 0, first = -1;
  CHECK(rmd_multi_size(multi, &n_ranks, &n_local, &first));
  REQUIRE(n_ranks == n && n_local == n && first == 0);
  const size_t px = (size_t)W * H;
  std::vector<float> all_depth(px * n), one_depth(px);
  std::vector<int32_t> all_conv(px * n), one_conv(px);
  for(int round = 0; round < 2; ++round)   // twice: buffers and communicator are reusable
  {
    std::fill(all_depth.begin(), all_depth.end(), -1.0f);
    CHECK(rmd_multi_gather_maps(multi, seeds.data(), NULL, NULL, 0, all_depth.data(), all_conv.data()));
    for(int r = 0; r < n; ++r)
    {
      CHECK(rmd_seeds_download(seeds[r], RMD_FIELD_MU, one_depth.data()));
      CHECK(rmd_seeds_download(seeds[r], RMD_FIELD_CONVERGENCE, one_conv.data()));
      REQUIRE(std::memcmp(one_depth.data(), all_depth.data() + px * r, px * sizeof(float)) == 0);
      REQUIRE(std::memcmp(one_conv.data(), all_conv.data() + px * r, px * sizeof(int32_t)) == 0);
      size_t interior_update = 0;
      for(size_t i = 0; i < px; ++i) interior_update += (one_conv[i] != RMD_BORDER);
      REQUIRE(interior_update == (size_t)(W - 10) * (H - 10));
    }
  }
  // a caller-provided device depth image (what the node sends after denoising) instead of mu
  {
    std::vector<float*> dev(n, (float*)NULL);
    std::vector<const float*> cdev(n);
    std::vector<size_t> pitch(n);
    for(int r = 0; r < n; ++r)
    {
      cudaSetDevice(r);
      void *p = NULL; size_t pb = 0;
      CHECK(rmd_image_alloc(W, H, sizeof(float), &p, &pb));
      std::vector<float> ramp(px);
      for(size_t i = 0; i < px; ++i) ramp[i] = (float)(r * 1000 + (int)(i % 997));
      CHECK(rmd_image_upload(p, pb, ramp.data(), W, H, sizeof(float)));
      dev[r] = (float*)p; cdev[r] = dev[r]; pitch[r] = pb;
    }
    CHECK(rmd_multi_gather_maps(multi, seeds.data(), cdev.data(), pitch.data(), 0, all_depth.data(), all_conv.data()));
    for(int r = 0; r < n; ++r)
      for(size_t i = 0; i < px; i += 101)
        REQUIRE(all_depth[px * r + i] == (float)(r * 1000 + (int)(i % 997)));
    for(int r = 0; r < n; ++r) { cudaSetDevice(r); rmd_image_free(dev[r]); }
  }
  CHECK(rmd_multi_destroy(multi));
  for(int r = 0; r < n; ++r) CHECK(rmd_seeds_destroy(seeds[r]));
  std::printf("MULTI GPU TEST PASSED on %d GPU(s)\n", n);
  return 0;
}
