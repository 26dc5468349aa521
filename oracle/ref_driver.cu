// ref_driver.cu -- TEST INFRASTRUCTURE: C entry points around the reference's
// own, unmodified rmd::SeedMatrix / rmd::DepthmapDenoiser / rmd::ImageReducer
// (compiled from /root/reference/src/*.cu by oracle/Makefile into
// oracle/_ref/librmd_ref*.so).  It gives tests/ and bench.py --impl reference
// the "reference CUDA build on the same box" without OpenCV/ROS/gtest.
// Nothing here is product code and no reference source is copied: the driver
// only calls the reference's public API (include/rmd/seed_matrix.cuh:45-109,
// depthmap_denoiser.cuh:27-54, reduction.cuh:27-62).
#include <cstdio>
#include <cstring>
#include <string>

#include <rmd/seed_matrix.cuh>
#include <rmd/depthmap_denoiser.cuh>
#include <rmd/reduction.cuh>
#include <rmd/se3.cuh>

namespace
{
std::string g_last_error;

rmd::SE3<float> se3_from(const float *v)
{
  rmd::SE3<float> T;
  for(int i = 0; i < 12; ++i) T.data[i] = v[i];
  return T;
}

struct RefSeeds
{
  RefSeeds(int w, int h, const rmd::PinholeCamera &cam)
    : width(w), height(h), seeds(w, h, cam) {}
  int width, height;
  rmd::SeedMatrix seeds;
};

template<typename F>
int guarded(F f)
{
  try
  {
    f();
    return 0;
  }
  catch(const rmd::CudaException &e)
  {
    g_last_error = e.what_ + std::string(": ") + cudaGetErrorString(e.err_);
    return static_cast<int>(e.err_) ? static_cast<int>(e.err_) : -1;
  }
  catch(const std::exception &e)
  {
    g_last_error = e.what();
    return -1;
  }
}
}

extern "C"
{

const char * ref_last_error() { return g_last_error.c_str(); }
int ref_patch_side() { return RMD_CORR_PATCH_SIDE; }

int ref_device_count()
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int ref_sync()
{
  return static_cast<int>(cudaDeviceSynchronize());
}

void * ref_seeds_create(int w, int h, float fx, float fy, float cx, float cy)
{
  RefSeeds *s = NULL;
  const int rc = guarded([&] { s = new RefSeeds(w, h, rmd::PinholeCamera(fx, fy, cx, cy)); });
  return rc == 0 ? s : NULL;
}

void ref_seeds_destroy(void *p)
{
  guarded([&] { delete static_cast<RefSeeds*>(p); });
}

int ref_seeds_set_reference(void *p, const float *img, const float *T_curr_world,
                            float min_depth, float max_depth)
{
  RefSeeds *s = static_cast<RefSeeds*>(p);
  return guarded([&] {
    s->seeds.setReferenceImage(const_cast<float*>(img), se3_from(T_curr_world), min_depth, max_depth);
  });
}

int ref_seeds_update(void *p, const float *img, const float *T_curr_world)
{
  RefSeeds *s = static_cast<RefSeeds*>(p);
  return guarded([&] { s->seeds.update(const_cast<float*>(img), se3_from(T_curr_world)); });
}

// field ids: 0 mu, 1 sigma_sq, 2 a, 3 b, 4 convergence(int), 5 sum_templ,
// 6 const_templ_denom, 7 epipolar matches (float2)
int ref_seeds_download(void *p, int field, void *dst)
{
  RefSeeds *s = static_cast<RefSeeds*>(p);
  return guarded([&] {
    switch(field)
    {
    case 0: s->seeds.downloadDepthmap(static_cast<float*>(dst)); break;
    case 1: s->seeds.downloadSigmaSq(static_cast<float*>(dst)); break;
    case 2: s->seeds.downloadA(static_cast<float*>(dst)); break;
    case 3: s->seeds.downloadB(static_cast<float*>(dst)); break;
    case 4: s->seeds.downloadConvergence(static_cast<int*>(dst)); break;
    case 5: s->seeds.downloadSumTempl(static_cast<float*>(dst)); break;
    case 6: s->seeds.downloadConstTemplDenom(static_cast<float*>(dst)); break;
    case 7: s->seeds.downloadEpipolarMatches(static_cast<float2*>(dst)); break;
    default: throw rmd::CudaException("ref_seeds_download: bad field", cudaErrorInvalidValue);
    }
  });
}

// Test hook: overwrite mu / sigma_sq / a / b through the public accessors
// (DeviceImage::setDevData is public, device_image.cuh:93-106).
int ref_seeds_upload(void *p, int field, const float *src)
{
  RefSeeds *s = static_cast<RefSeeds*>(p);
  return guarded([&] {
    const rmd::DeviceImage<float> *img = NULL;
    switch(field)
    {
    case 0: img = &s->seeds.getMu(); break;
    case 1: img = &s->seeds.getSigmaSq(); break;
    case 2: img = &s->seeds.getA(); break;
    case 3: img = &s->seeds.getB(); break;
    default: throw rmd::CudaException("ref_seeds_upload: bad field", cudaErrorInvalidValue);
    }
    const_cast<rmd::DeviceImage<float>*>(img)->setDevData(src);
  });
}

long long ref_seeds_converged_count(void *p)
{
  RefSeeds *s = static_cast<RefSeeds*>(p);
  long long n = -1;
  guarded([&] { n = static_cast<long long>(s->seeds.getConvergedCount()); });
  return n;
}

float ref_seeds_dist_from_ref(void *p)
{
  return static_cast<RefSeeds*>(p)->seeds.getDistFromRef();
}

void * ref_denoiser_create(int w, int h)
{
  rmd::DepthmapDenoiser *d = NULL;
  const int rc = guarded([&] { d = new rmd::DepthmapDenoiser(w, h); });
  return rc == 0 ? d : NULL;
}

void ref_denoiser_destroy(void *p)
{
  guarded([&] { delete static_cast<rmd::DepthmapDenoiser*>(p); });
}

int ref_denoiser_run(void *dp, void *sp, float depth_range, float lambda, int iterations,
                     float *host_out)
{
  rmd::DepthmapDenoiser *d = static_cast<rmd::DepthmapDenoiser*>(dp);
  RefSeeds *s = static_cast<RefSeeds*>(sp);
  return guarded([&] {
    d->setLargeSigmaSq(depth_range);
    d->denoise(s->seeds.getMu(), s->seeds.getSigmaSq(), s->seeds.getA(), s->seeds.getB(),
               host_out, lambda, iterations);
  });
}

// Reductions on a host image uploaded into a DeviceImage, launch shape as in
// test/reduction_test.cpp:51-57 (16x16 threads, 4x4 blocks).
int ref_reduce_sum_f32(const float *host_img, int w, int h, float *out)
{
  return guarded([&] {
    rmd::DeviceImage<float> img(w, h);
    img.setDevData(host_img);
    rmd::ImageReducer<float> reducer(dim3(16, 16), dim3(4, 4));
    *out = reducer.sum(img);
  });
}

int ref_reduce_sum_i32(const int *host_img, int w, int h, int *out)
{
  return guarded([&] {
    rmd::DeviceImage<int> img(w, h);
    img.setDevData(host_img);
    rmd::ImageReducer<int> reducer(dim3(16, 16), dim3(4, 4));
    *out = reducer.sum(img);
  });
}

int ref_reduce_count_eq_i32(const int *host_img, int w, int h, int value, long long *out)
{
  return guarded([&] {
    rmd::DeviceImage<int> img(w, h);
    img.setDevData(host_img);
    rmd::ImageReducer<int> reducer(dim3(16, 16), dim3(4, 4));
    *out = static_cast<long long>(reducer.countEqual(img, value));
  });
}

} // extern "C"
