// point_cloud.cuh -- point-cloud extraction behind the depth filter
// (SURVEY.md 8f row 3): rmd::Publisher::publishPointCloud, src/publisher.cpp:54-86,
// whose CPU double loop back-projects every CONVERGED pixel with T_world_ref and
// tags it with the reference image's 8-bit intensity, after downloading the
// whole depth and convergence maps.  Here: ordered stream compaction on the
// device (same row-major point order), only the points leave the GPU.
#pragma once

#include <cuda_runtime.h>

#include "rmd_common.cuh"

namespace rmdb
{

struct PointCloudParams
{
  int width, height;
  const int *conv; int conv_stride;          // ConvergenceState per pixel
  const float *depth; int depth_stride;      // floats per row ...
  int depth_comps;                           // ... and per pixel: 4 = mu inside the float4 seed records, 1 = planar image
  const float *ref; int ref_stride;          // reference image in [0,1] (intensity = rint(255 * ref))
  Camera cam;
  Pose T_world_ref;
  float4 *out;                               // (x, y, z, intensity) per point, row-major pixel order
  unsigned int capacity;                     // points that fit in out
  unsigned int *block_counts;                // [n_blocks] converged pixels per block, then exclusive offsets
  unsigned int *total;                       // [0] number of points, [1] ticket of the counting pass
  int n_blocks;
};

constexpr int POINT_CLOUD_BLOCK = 256;          // threads
constexpr int POINT_CLOUD_PIXELS = 4 * POINT_CLOUD_BLOCK;   // consecutive pixels (row-major) per block

// Two launches: count (+ scan of the block totals by the last block), write.
cudaError_t launch_point_cloud(const PointCloudParams &P, cudaStream_t stream);

} // namespace rmdb
