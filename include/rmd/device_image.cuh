// rmd/device_image.cuh -- rmd::DeviceImage<T> with the reference's interface
// (include/rmd/device_image.cuh:34-180): RAII pitched 2-D device buffer with
// public fields width, height, pitch, stride (elements), data, and
// setDevData / getDevData / zero / operator= / getCudaChannelFormatDesc.
// All memory operations go through the C-ABI (rmd_image_*, include/rmd_b200.h).
// Differences: `dev_ptr` (a device-resident copy of this header that only the
// reference's own kernels dereference) is kept for layout compatibility but is
// always null; destructors do not throw; images are non-copyable (the
// reference's implicit copy would double-free).
#ifndef DEVICE_IMAGE_CUH
#define DEVICE_IMAGE_CUH

#include <cassert>
#include <cuda_runtime.h>
#include <rmd/cuda_exception.cuh>

namespace rmd
{

struct Size
{
  int width;
  int height;
};

template<typename ElementType>
struct DeviceImage
{
  DeviceImage(size_t width_, size_t height_)
    : width(width_), height(height_), pitch(0), stride(0), data(NULL), dev_ptr(NULL), owns_(true)
  {
    void *ptr = NULL;
    detail::throw_on_error(rmd_image_alloc(width, height, sizeof(ElementType), &ptr, &pitch),
                           "Image: unable to allocate pitched memory.");
    data = static_cast<ElementType*>(ptr);
    stride = pitch / sizeof(ElementType);
  }

  // Non-owning view of memory held by a seed matrix (getMu() etc.).
  struct ViewTag {};
  DeviceImage(ViewTag, size_t width_, size_t height_, ElementType *ptr, size_t pitch_)
    : width(width_), height(height_), pitch(pitch_), stride(pitch_ / sizeof(ElementType)), data(ptr),
      dev_ptr(NULL), owns_(false) {}

  void rebind(ElementType *ptr, size_t pitch_)  // views only
  {
    assert(!owns_);
    data = ptr; pitch = pitch_; stride = pitch_ / sizeof(ElementType);
  }

  ~DeviceImage()
  {
    if(owns_ && data)
      rmd_image_free(data);
  }

  /// Upload aligned_data_row_major (densely packed, width*sizeof(T) per row)
  void setDevData(const ElementType *aligned_data_row_major)
  {
    detail::throw_on_error(
        rmd_image_upload(data, pitch, aligned_data_row_major, width, height, sizeof(ElementType)),
        "Image: unable to copy data from host to device.");
  }

  /// Download into a preallocated, densely packed host array
  void getDevData(ElementType *aligned_data_row_major) const
  {
    detail::throw_on_error(
        rmd_image_download(data, pitch, aligned_data_row_major, width, height, sizeof(ElementType)),
        "Image: unable to copy data from device to host.");
  }

  cudaChannelFormatDesc getCudaChannelFormatDesc() const { return cudaCreateChannelDesc<ElementType>(); }

  void zero()
  {
    detail::throw_on_error(rmd_image_zero(data, pitch, width, height, sizeof(ElementType)),
                           "Image: unable to zero.");
  }

  DeviceImage<ElementType> &operator=(const DeviceImage<ElementType> &other_image)
  {
    if(this != &other_image)
    {
      assert(width == other_image.width && height == other_image.height);
      detail::throw_on_error(rmd_image_copy(data, pitch, other_image.data, other_image.pitch, width, height,
                                            sizeof(ElementType)),
                             "Image, operator '=': unable to copy data from another image.");
    }
    return *this;
  }

  // fields (public in the reference, used by callers: src/reduction.cu:123-126)
  size_t width;
  size_t height;
  size_t pitch;
  size_t stride;
  ElementType *data;
  DeviceImage<ElementType> *dev_ptr;

private:
  DeviceImage(const DeviceImage &);  // non-copyable
  bool owns_;
};

} // namespace rmd

#endif // DEVICE_IMAGE_CUH
