// staged_maps.cuh -- TMA descriptor set of the staged depth filter.
#pragma once

#include <cuda.h>

#include "rmd_common.cuh"

namespace rmdb
{

// Geometry shared by the host encoder and the kernel.
namespace staged
{
constexpr int TILE_W = 32;             // pixels per CTA row (one warp)
constexpr int TILE_H = 8;              // rows per CTA (one per warp)
constexpr int NTHREADS = TILE_W * TILE_H;
constexpr int CHUNK = 8;               // candidates per work item
constexpr int MAX_CHUNKS = 18;         // ceil(143 / CHUNK): 100 px / 0.7 px + 1 = 143 candidates
constexpr int STRIP_FLOATS = 10240;    // 40 KB current-image strip per CTA
constexpr int STRIP_BOX_ROWS = 8;      // rows per TMA box of the strip
constexpr int REF_BOX_W = 40;          // >= TILE_W + 7 - 1, multiple of 4 floats (16 B)
constexpr int NUM_WIDTHS = 4;
__host__ __device__ constexpr int strip_width(int i)
{
  return i == 0 ? 48 : (i == 1 ? 80 : (i == 2 ? 112 : 160));
}
__host__ __device__ constexpr int ref_box_h(int patch) { return TILE_H + patch - 1; }
}

// The descriptors as the kernel receives them (one __grid_constant__ param).
struct alignas(64) StagedTensorMaps
{
  CUtensorMap ref;                         // box REF_BOX_W x ref_box_h(patch)
  CUtensorMap curr[staged::NUM_WIDTHS];    // boxes strip_width(i) x STRIP_BOX_ROWS
};

// Tiled 2-D tensor maps (cuTensorMapEncodeTiled) over the pitched reference
// and current images.  Box shapes are fixed when a map is encoded; the box
// origin is a run-time coordinate of the TMA instruction, so one map per box
// width serves every CTA.  The current image changes address every frame
// (upload ring / caller memory), hence encode() per frame: it is a pure host
// computation of 128-byte descriptors, no driver round trip.
struct StagedMaps
{
  StagedTensorMaps maps;
  int patch;
  const void *ref_ptr, *curr_ptr;
  int ref_stride, curr_stride, width, height;

  StagedMaps();
  // Returns 0 or an error code (and sets the thread's last error string).
  int encode(const FilterParams &P, int patch_side);
};

} // namespace rmdb
