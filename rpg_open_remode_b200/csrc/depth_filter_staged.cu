// depth_filter_staged.cu -- TMA-staged variant of the fused depth filter.
// (placeholder until the staged kernel lands: reports "not supported" so the
// caller can never silently fall back)
#include "depth_filter.cuh"
#include "staged_maps.cuh"

namespace rmdb
{

StagedMaps::StagedMaps() : patch(0), ref_ptr(NULL), curr_ptr(NULL), ref_stride(0), curr_stride(0), width(0), height(0) {}

int StagedMaps::encode(const FilterParams &, int)
{
  return fail(RMD_ERR_UNSUPPORTED, "staged depth-filter kernel is not built into this library");
}

cudaError_t launch_depth_filter_staged(const FilterParams &, const StagedMaps &, int, cudaStream_t)
{
  return cudaErrorNotSupported;
}

} // namespace rmdb
