/* rmd_oracle_pointcloud.c -- CPU restatement of rmd::Publisher::publishPointCloud
 * (src/publisher.cpp:54-86), the step right after the depth filter: the reference's
 * own CPU double loop, without ROS / PCL / OpenCV containers.
 *
 * TEST INFRASTRUCTURE ONLY (see rmd_oracle.h).  Unlike the depth filter, this step IS
 * CPU code in the reference, so the restatement is a transcription: same loop order,
 * same float expressions (gcc -O3 on x86-64, CMakeLists.txt:33: SSE2 arithmetic, no
 * fused multiply-add; this file is built with -ffp-contract=off to say the same).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define RMDO_CONVERGED 1 /* include/rmd/seed_matrix.cuh:37-45 */

/* depth, conv, ref_u8: dense row-major w x h maps (what Depthmap::getDepthmap,
 * getConvergenceMap, getReferenceImage return); T_world_ref: 3x4 row-major.
 * out: 4 floats per point (x, y, z, intensity) or NULL to count only.
 * Returns the number of points. */
size_t rmd_oracle_point_cloud(const float *depth, const int *conv, const uint8_t *ref_u8,
                              int width, int height, float fx, float fy, float cx, float cy,
                              const float *T_world_ref, float *out) {
  size_t n = 0;
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) {
      /* const float3 f = normalize(make_float3((x-cx)/fx, (y-cy)/fy, 1.0f));   publisher.cpp:73 */
      const float vx = (x - cx) / fx, vy = (y - cy) / fy, vz = 1.0f;
      const float dot = vx * vx + vy * vy + vz * vz;  /* helper_math.h dot(float3, float3) */
      const float inv_len = 1.0f / sqrtf(dot);        /* host rsqrtf, helper_math.h:62-65 */
      const float f_x = vx * inv_len, f_y = vy * inv_len, f_z = vz * inv_len;
      /* const float3 xyz = T_world_ref * (f * depth.at<float>(y, x));           publisher.cpp:74 */
      const float d = depth[(size_t)y * width + x];
      const float px = f_x * d, py = f_y * d, pz = f_z * d;
      const float *T = T_world_ref;
      const float rx = T[0] * px + T[1] * py + T[2] * pz; /* se3.cuh:111-116 */
      const float ry = T[4] * px + T[5] * py + T[6] * pz;
      const float rz = T[8] * px + T[9] * py + T[10] * pz;
      if (conv[(size_t)y * width + x] == RMDO_CONVERGED) { /* publisher.cpp:75 */
        if (out) {
          out[4 * n + 0] = rx + T[3]; /* se3.cuh:119-124 */
          out[4 * n + 1] = ry + T[7];
          out[4 * n + 2] = rz + T[11];
          out[4 * n + 3] = (float)ref_u8[(size_t)y * width + x]; /* publisher.cpp:81-82 */
        }
        ++n;
      }
    }
  }
  return n;
}
