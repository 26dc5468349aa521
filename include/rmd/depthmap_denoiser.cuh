// rmd/depthmap_denoiser.cuh -- rmd::DepthmapDenoiser with the reference's
// interface (include/rmd/depthmap_denoiser.cuh:27-54), forwarding to
// rmd_denoiser_* (include/rmd_b200.h).
#ifndef RMD_DEPTHMAP_DENOISER_CUH
#define RMD_DEPTHMAP_DENOISER_CUH

#include <iostream>
#include <rmd/device_image.cuh>

namespace rmd
{

class DepthmapDenoiser
{
public:
  DepthmapDenoiser(size_t width, size_t height) : handle_(NULL)
  {
    detail::throw_on_error(rmd_denoiser_create(static_cast<int>(width), static_cast<int>(height), -1, &handle_),
                           "DepthmapDenoiser: unable to create");
  }
  ~DepthmapDenoiser() { rmd_denoiser_destroy(handle_); }

  void denoise(const rmd::DeviceImage<float> &mu, const rmd::DeviceImage<float> &sigma_sq,
               const rmd::DeviceImage<float> &a, const rmd::DeviceImage<float> &b, float *host_denoised,
               float lambda, int iterations)
  {
    const int rc = rmd_denoiser_run(handle_, mu.data, mu.pitch, sigma_sq.data, sigma_sq.pitch, a.data, a.pitch,
                                    b.data, b.pitch, host_denoised, lambda, iterations);
    if(rc == RMD_ERR_NOT_INITIALISED)
    {
      // the reference's behaviour, src/depthmap_denoiser.cu:189-193
      std::cerr << "ERROR: setLargeSigmaSq must be called before this method" << std::endl;
      return;
    }
    detail::throw_on_error(rc, "DepthmapDenoiser: unable to denoise");
  }

  void setLargeSigmaSq(float depth_range) { rmd_denoiser_set_large_sigma_sq(handle_, depth_range); }

  rmd_denoiser_t *handle() const { return handle_; }

private:
  DepthmapDenoiser(const DepthmapDenoiser &);
  DepthmapDenoiser &operator=(const DepthmapDenoiser &);
  rmd_denoiser_t *handle_;
};

} // rmd namespace

#endif // RMD_DEPTHMAP_DENOISER_CUH
