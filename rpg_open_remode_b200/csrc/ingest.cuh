// ingest.cuh -- frame ingest in front of the depth filter: lens undistortion
// of the 8-bit camera image and conversion to float, on the GPU.
//
// Replaces rmd::Depthmap::initUndistortionMap / inputImage
// (src/depthmap.cpp:45-61, 95-106), which call OpenCV on the host:
//   cv::initUndistortRectifyMap(K, (k1 k2 r1 r2), I, K, size, CV_16SC2, map1, map2)
//   cv::remap(img_8uc1, undistorted_8uc1, map1, map2, CV_INTER_LINEAR)
//   undistorted_8uc1.convertTo(img_32fc1, CV_32F, 1.0f / 255.0f)
// OpenCV is a dependency of the reference that is not in its tree (and not
// pinned: find_package(OpenCV REQUIRED), CMakeLists.txt:55); the algorithm is
// restated from its documented behaviour and pinned bit for bit against
// OpenCV 4.13 (tests/golden/make_golden_undistort.py).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "undistort_maps.h"   // compute_undistort_maps(): host, once per camera

namespace rmdb
{

// One pass over the frame: bilinear remap of the 8-bit source through the maps
// (constant 0 outside the image, 15-bit fixed-point weights, round to nearest
// like cv::remap), optional 8-bit result (the reference keeps it for the
// point-cloud intensities), float result * (1/255) for the depth filter.
cudaError_t launch_undistort_u8(const uint8_t *src, int src_pitch, const short2 *map_xy, const uint16_t *map_frac,
                                float *dst_f32, int dst_f32_stride, uint8_t *dst_u8, int dst_u8_pitch,
                                int width, int height, cudaStream_t stream);

} // namespace rmdb
