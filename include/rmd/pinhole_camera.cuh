// rmd/pinhole_camera.cuh -- rmd::PinholeCamera with the reference's interface
// (include/rmd/pinhole_camera.cuh:27-63): public fx, fy, cx, cy; cam2world,
// world2cam (pixel coordinates are the integer sample positions, no +0.5),
// getOnePixAngle (uses fx only).
#ifndef RMD_PINHOLE_CAMERA_CUH_
#define RMD_PINHOLE_CAMERA_CUH_

#include <cuda_runtime.h>
#include <cmath>

namespace rmd
{

struct PinholeCamera
{
  PinholeCamera() : fx(0.0f), fy(0.0f), cx(0.0f), cy(0.0f) {}
  PinholeCamera(float fx_, float fy_, float cx_, float cy_) : fx(fx_), fy(fy_), cx(cx_), cy(cy_) {}

  // bearing (not normalised) of pixel uv
  float3 cam2world(const float2 &uv) const
  {
    return make_float3((uv.x - cx) / fx, (uv.y - cy) / fy, 1.0f);
  }

  // projection of a point given in the camera frame
  float2 world2cam(const float3 &xyz) const
  {
    return make_float2(fx * xyz.x / xyz.z + cx, fy * xyz.y / xyz.z + cy);
  }

  // angle subtended by one pixel
  float getOnePixAngle() const { return atan2f(1.0f, 2.0f * fx) * 2.0f; }

  float fx, fy;
  float cx, cy;
};

} // namespace rmd

#endif // RMD_PINHOLE_CAMERA_CUH_
