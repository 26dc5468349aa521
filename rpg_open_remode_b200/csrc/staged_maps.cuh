// staged_maps.cuh -- host-side TMA descriptor set of the staged depth filter.
#pragma once

#include <cuda.h>

#include "rmd_common.cuh"

namespace rmdb
{

// Tiled 2-D tensor maps (cuTensorMapEncodeTiled) over the pitched reference
// and current images.  Box shapes are fixed when a map is encoded; the box
// origin is a run-time coordinate of the TMA instruction, so one map per box
// width serves every CTA.  The current image changes address every frame
// (upload ring / caller memory), hence encode() per frame: it is a pure host
// computation of a 128-byte descriptor, no driver round trip.
struct StagedMaps
{
  static const int kNumWidths = 3;
  CUtensorMap ref_map;                 // box: (REF_BOX_W x REF_BOX_H)
  CUtensorMap curr_map[kNumWidths];    // boxes: (kWidths[i] x STRIP_ROWS_PER_BOX)
  int patch;
  const void *ref_ptr, *curr_ptr;
  int ref_stride, curr_stride, width, height;

  StagedMaps();
  // Returns 0 or an error code (and sets the thread's last error string).
  int encode(const FilterParams &P, int patch_side);
};

} // namespace rmdb
