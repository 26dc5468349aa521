// ingest.cu -- see ingest.cuh.
#include "ingest.cuh"

namespace rmdb
{

namespace
{

__global__ void undistort_u8_kernel(const uint8_t *__restrict__ src, int src_pitch,
                                    const short2 *__restrict__ map_xy, const uint16_t *__restrict__ map_frac,
                                    float *__restrict__ dst_f32, int dst_f32_stride,
                                    uint8_t *__restrict__ dst_u8, int dst_u8_pitch, int width, int height)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if(x >= width || y >= height)
    return;
  const size_t o = (size_t)y * width + x;
  const short2 m = map_xy[o];
  const unsigned int f = map_frac[o];
  const int ax = (int)(f & 31u), ay = (int)(f >> 5);
  const int sx = m.x, sy = m.y;
  // BORDER_CONSTANT, value 0 (cv::remap's default): taps outside the image read 0
  const bool x0 = (unsigned)sx < (unsigned)width, x1 = (unsigned)(sx + 1) < (unsigned)width;
  const bool y0 = (unsigned)sy < (unsigned)height, y1 = (unsigned)(sy + 1) < (unsigned)height;
  const uint8_t *r0 = src + (size_t)(y0 ? sy : 0) * src_pitch, *r1 = src + (size_t)(y1 ? sy + 1 : 0) * src_pitch;
  const int p00 = (x0 && y0) ? (int)__ldg(r0 + sx) : 0, p01 = (x1 && y0) ? (int)__ldg(r0 + sx + 1) : 0;
  const int p10 = (x0 && y1) ? (int)__ldg(r1 + sx) : 0, p11 = (x1 && y1) ? (int)__ldg(r1 + sx + 1) : 0;
  // weights (32-ax)(32-ay)/1024 etc. in 1.15 fixed point are exact integers: w * 32
  const int acc = ((32 - ax) * (32 - ay) * p00 + ax * (32 - ay) * p01 + (32 - ax) * ay * p10 + ax * ay * p11) * 32;
  const int v = (acc + (1 << 14)) >> 15;
  if(dst_u8)
    dst_u8[(size_t)y * dst_u8_pitch + x] = (uint8_t)v;
  if(dst_f32)
    dst_f32[(size_t)y * dst_f32_stride + x] = (float)v * (1.0f / 255.0f);   // convertTo(CV_32F, 1.0f/255.0f)
}

} // namespace

cudaError_t launch_undistort_u8(const uint8_t *src, int src_pitch, const short2 *map_xy, const uint16_t *map_frac,
                                float *dst_f32, int dst_f32_stride, uint8_t *dst_u8, int dst_u8_pitch,
                                int width, int height, cudaStream_t stream)
{
  const dim3 block(32, 8);
  const dim3 grid((width + block.x - 1) / block.x, (height + block.y - 1) / block.y);
  undistort_u8_kernel<<<grid, block, 0, stream>>>(src, src_pitch, map_xy, map_frac, dst_f32, dst_f32_stride,
                                                   dst_u8, dst_u8_pitch, width, height);
  return cudaGetLastError();
}

} // namespace rmdb
