/* rmd_oracle_ingest.c -- CPU restatement of the reference's frame ingest,
 * rmd::Depthmap::initUndistortionMap + inputImage (src/depthmap.cpp:45-61,95-106).
 *
 * TEST INFRASTRUCTURE ONLY (see rmd_oracle.h): nothing in the product links or
 * loads this file.
 *
 * The reference does these steps with OpenCV on the host:
 *   cv::initUndistortRectifyMap(cv_K_, cv_D_, I, cv_K_, size, CV_16SC2, map1, map2)   depthmap.cpp:52-59
 *   cv::remap(img_8uc1, img_undistorted_8uc1_, map1, map2, CV_INTER_LINEAR)          depthmap.cpp:99
 *   img_undistorted_8uc1_.convertTo(img_undistorted_32fc1_, CV_32F, 1.0f/255.0f)     depthmap.cpp:105
 * OpenCV is a dependency that is absent from /root/reference and unpinned there
 * (find_package(OpenCV REQUIRED), CMakeLists.txt:55).  The algorithm below
 * restates its documented behaviour (calib3d initUndistortRectifyMap, imgproc
 * remap with fixed-point maps, core convertTo) and is PINNED bit for bit against
 * OpenCV 4.13.0 by the golden vectors of tests/golden/make_golden_undistort.py
 * (maps, remapped images and float conversion for three cameras, including the
 * reference's own launch/px4_2.launch parameters).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define RMDO_INTER_BITS 5
#define RMDO_INTER_TAB_SIZE (1 << RMDO_INTER_BITS)
#define RMDO_REMAP_COEF_BITS 15

/* cv::invert of a 3x3 double matrix (any method): adjugate * (1 / det). */
static void invert3(const double m[3][3], double out[3][3]) {
  const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) -
                     m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                     m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  const double d = 1. / det;
  out[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * d;
  out[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * d;
  out[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * d;
  out[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * d;
  out[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * d;
  out[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * d;
  out[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * d;
  out[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * d;
  out[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * d;
}

/* cvRound on x86-64 (cvtsd2si): nearest, ties to even; NaN and out-of-range
 * values give INT_MIN. */
static int cv_round(double v) {
  const double r = nearbyint(v);
  if (!(r >= -2147483648.0 && r <= 2147483647.0)) return (int)0x80000000;
  return (int)r;
}

/* Maps of cv::initUndistortRectifyMap(K, (k1 k2 p1 p2), R = I, newK = K, CV_16SC2):
 * map1[2*i], map1[2*i+1] = integer source pixel, map2[i] = (fy << 5) | fx. */
void rmd_oracle_undistort_maps(int width, int height, float fx, float fy, float cx,
                               float cy, float k1f, float k2f, float p1f, float p2f,
                               int16_t *map1, uint16_t *map2) {
  const double K[3][3] = {{fx, 0, cx}, {0, fy, cy}, {0, 0, 1}};
  double iR[3][3];
  invert3(K, iR); /* (newK * R).inv(DECOMP_LU) */
  const double k1 = k1f, k2 = k2f, p1 = p1f, p2 = p2f, k3 = 0, k4 = 0, k5 = 0, k6 = 0;
  const double u0 = cx, v0 = cy, fxd = fx, fyd = fy;
  for (int i = 0; i < height; ++i) {
    double _x = i * iR[0][1] + iR[0][2], _y = i * iR[1][1] + iR[1][2],
           _w = i * iR[2][1] + iR[2][2];
    for (int j = 0; j < width; ++j) {
      const double w = 1. / _w, x = _x * w, y = _y * w;
      const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
      const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) /
                        (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
      const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
      const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
      const double u = fxd * xd + u0, v = fyd * yd + v0;
      const int iu = cv_round(u * RMDO_INTER_TAB_SIZE);
      const int iv = cv_round(v * RMDO_INTER_TAB_SIZE);
      const size_t o = (size_t)i * width + j;
      map1[2 * o] = (int16_t)(iu >> RMDO_INTER_BITS);
      map1[2 * o + 1] = (int16_t)(iv >> RMDO_INTER_BITS);
      map2[o] = (uint16_t)((iv & (RMDO_INTER_TAB_SIZE - 1)) * RMDO_INTER_TAB_SIZE +
                           (iu & (RMDO_INTER_TAB_SIZE - 1)));
      _x += iR[0][0];
      _y += iR[1][0];
      _w += iR[2][0];
    }
  }
}

/* cv::remap(8UC1, INTER_LINEAR, BORDER_CONSTANT 0) through fixed-point maps: the
 * four bilinear weights are held in 1.15 fixed point (for 5-bit fractions they
 * are exact integers), the sum is rounded to nearest. */
void rmd_oracle_remap_u8(const uint8_t *src, int width, int height, const int16_t *map1,
                         const uint16_t *map2, uint8_t *dst) {
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const size_t o = (size_t)y * width + x;
      const int sx = map1[2 * o], sy = map1[2 * o + 1];
      const int ax = map2[o] & (RMDO_INTER_TAB_SIZE - 1), ay = map2[o] >> RMDO_INTER_BITS;
      int tap[2][2];
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          const int px = sx + dx, py = sy + dy;
          tap[dy][dx] = (px >= 0 && px < width && py >= 0 && py < height)
                            ? src[(size_t)py * width + px]
                            : 0;
        }
      const int scale = 1 << (RMDO_REMAP_COEF_BITS - 2 * RMDO_INTER_BITS); /* 32 */
      const int w00 = (RMDO_INTER_TAB_SIZE - ax) * (RMDO_INTER_TAB_SIZE - ay) * scale;
      const int w01 = ax * (RMDO_INTER_TAB_SIZE - ay) * scale;
      const int w10 = (RMDO_INTER_TAB_SIZE - ax) * ay * scale;
      const int w11 = ax * ay * scale;
      const int sum = w00 * tap[0][0] + w01 * tap[0][1] + w10 * tap[1][0] + w11 * tap[1][1];
      dst[o] = (uint8_t)((sum + (1 << (RMDO_REMAP_COEF_BITS - 1))) >> RMDO_REMAP_COEF_BITS);
    }
}

/* Mat::convertTo(CV_32F, 1.0f/255.0f) of an 8-bit image: float multiply. */
void rmd_oracle_u8_to_float(const uint8_t *src, size_t n, float *dst) {
  const float alpha = 1.0f / 255.0f;
  for (size_t i = 0; i < n; ++i) dst[i] = (float)src[i] * alpha;
}
