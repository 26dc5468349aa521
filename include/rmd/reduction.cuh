// rmd/reduction.cuh -- rmd::ImageReducer<T> with the reference's interface
// (include/rmd/reduction.cuh:27-62; T in {int, float}, src/reduction.cu:186-187).
// The launch-shape arguments are accepted for source compatibility; the
// single-launch warp-shuffle kernels size their own grid (rmd_reduce_*).
#ifndef RMD_REDUCTION_CUH
#define RMD_REDUCTION_CUH

#include <rmd/device_image.cuh>

namespace rmd
{

template<typename T>
class ImageReducer
{
public:
  ImageReducer(dim3 num_threads_per_block, dim3 num_blocks_per_grid)
    : block_dim_(num_threads_per_block), grid_dim_(num_blocks_per_grid) {}
  ~ImageReducer() {}

  // Sum image by reduction; stride in elements
  T sum(const T *in_img_data, size_t in_img_stride, size_t in_img_width, size_t in_img_height);
  T sum(const DeviceImage<T> &in_img)
  {
    return sum(in_img.data, in_img.stride, in_img.width, in_img.height);
  }

  // Count elements equal to 'value'
  size_t countEqual(const int *in_img_data, size_t in_img_stride, size_t in_img_width,
                    size_t in_img_height, int value)
  {
    size_t count = 0;
    detail::throw_on_error(
        rmd_reduce_count_eq_i32(in_img_data, in_img_stride, in_img_width, in_img_height, value, &count),
        "countEqual: unable to reduce");
    return count;
  }
  size_t countEqual(const DeviceImage<int> &in_img, int value)
  {
    return countEqual(in_img.data, in_img.stride, in_img.width, in_img.height, value);
  }

private:
  dim3 block_dim_;
  dim3 grid_dim_;
};

template<>
inline int ImageReducer<int>::sum(const int *d, size_t stride, size_t w, size_t h)
{
  int32_t out = 0;
  detail::throw_on_error(rmd_reduce_sum_i32(d, stride, w, h, &out), "sum: unable to reduce");
  return out;
}

template<>
inline float ImageReducer<float>::sum(const float *d, size_t stride, size_t w, size_t h)
{
  float out = 0.0f;
  detail::throw_on_error(rmd_reduce_sum_f32(d, stride, w, h, &out), "sum: unable to reduce");
  return out;
}

} // namespace rmd

#endif // RMD_REDUCTION_CUH
