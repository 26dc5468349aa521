// ingest.cu -- see ingest.cuh.
#include "ingest.cuh"

namespace rmdb
{

void compute_undistort_maps(int width, int height, float fx_f, float fy_f, float cx_f, float cy_f,
                            float k1_f, float k2_f, float p1_f, float p2_f, int16_t *xy, uint16_t *frac)
{
  // cv::initUndistortRectifyMap with R = I and newCameraMatrix = cameraMatrix
  // (src/depthmap.cpp:52-59), everything in double as there.
  const double S[9] = {fx_f, 0.0, cx_f, 0.0, fy_f, cy_f, 0.0, 0.0, 1.0};   // cv_K_, src/depthmap.cpp:35
  const double fx = fx_f, fy = fy_f, u0 = cx_f, v0 = cy_f;
  const double k1 = k1_f, k2 = k2_f, p1 = p1_f, p2 = p2_f;
  // iR = (K * I).inv(DECOMP_LU): cv::invert treats 3x3 by the adjugate times the
  // reciprocal determinant, so e.g. iR(2,2) = (fx*fy) * (1/(fx*fy)) is not exactly 1
  double ir[9];
  {
    const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                       S[2] * (S[3] * S[7] - S[4] * S[6]);
    const double d = 1.0 / det;
    ir[0] = (S[4] * S[8] - S[5] * S[7]) * d;
    ir[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    ir[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    ir[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    ir[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    ir[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    ir[6] = (S[3] * S[7] - S[4] * S[6]) * d;
    ir[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    ir[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  }
  for(int i = 0; i < height; ++i)
  {
    // the homogeneous coordinate of a row is ACCUMULATED column by column,
    // which is part of the result's rounding and therefore kept
    double hx = i * ir[1] + ir[2], hy = i * ir[4] + ir[5], hw = i * ir[7] + ir[8];
    for(int j = 0; j < width; ++j, hx += ir[0], hy += ir[3], hw += ir[6])
    {
      const double w = 1.0 / hw, x = hx * w, y = hy * w;
      const double x2 = x * x, y2 = y * y;
      const double r2 = x2 + y2, two_xy = 2.0 * x * y;
      // k3..k6 = 0: (1 + ((k3*r2 + k2)*r2 + k1)*r2) / (1 + ((k6*r2 + k5)*r2 + k4)*r2)
      const double kr = (1.0 + ((0.0 * r2 + k2) * r2 + k1) * r2) / (1.0 + ((0.0 * r2 + 0.0) * r2 + 0.0) * r2);
      const double xd = x * kr + p1 * two_xy + p2 * (r2 + 2.0 * x2);
      const double yd = y * kr + p1 * (r2 + 2.0 * y2) + p2 * two_xy;
      const double u = fx * xd + u0, v = fy * yd + v0;
      // saturate_cast<int>(u * INTER_TAB_SIZE)
      // = cvRound = cvtsd2si: round half to even; NaN / out of range give INT_MIN
      const double su = __builtin_rint(u * 32.0), sv = __builtin_rint(v * 32.0);
      const int iu = (su >= -2147483648.0 && su <= 2147483647.0) ? (int)su : (int)0x80000000;
      const int iv = (sv >= -2147483648.0 && sv <= 2147483647.0) ? (int)sv : (int)0x80000000;
      const size_t o = (size_t)i * width + j;
      xy[2 * o] = (int16_t)(iu >> 5);
      xy[2 * o + 1] = (int16_t)(iv >> 5);
      frac[o] = (uint16_t)((iv & 31) * 32 + (iu & 31));
    }
  }
}

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
