// undistort_maps.h -- the fixed-point undistortion maps of the frame ingest
// (ingest.cuh), plain host C++ so that the CPU test suite can pin THIS code
// (not only the oracle) against OpenCV: tests/test_ingest_undistort.py.
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace rmdb
{

// Maps of cv::initUndistortRectifyMap(K, (k1 k2 p1 p2), I, K, size, CV_16SC2) as
// rmd::Depthmap::initUndistortionMap builds them (src/depthmap.cpp:45-61):
// xy[2i], xy[2i+1] = integer source pixel, frac[i] = (fy << 5) | fx with 5-bit
// fractions.  Double precision, once per camera.
inline void compute_undistort_maps(int width, int height, float fx_f, float fy_f, float cx_f, float cy_f,
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

} // namespace rmdb
