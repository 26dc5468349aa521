// rmd/helper_vector_types.cuh -- norm() of a CUDA vector type, the one helper
// the reference's header of this name provides (helper_vector_types.cuh:23-28).
#ifndef RMD_HELPER_VECTOR_TYPES_CUH
#define RMD_HELPER_VECTOR_TYPES_CUH

#include <cuda_runtime.h>
#include <cmath>

#if defined(__has_include)
#if __has_include(<cuda_toolkit/helper_math.h>)
#include <cuda_toolkit/helper_math.h>  // the NVIDIA sample header callers already vendor
#define RMD_HAVE_HELPER_MATH 1
#endif
#endif

namespace rmd
{
namespace detail
{
inline float dot(const float2 &a, const float2 &b) { return a.x * b.x + a.y * b.y; }
inline float dot(const float3 &a, const float3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
}
}

template<typename VectorType>
inline float norm(const VectorType &v)
{
  return sqrtf(rmd::detail::dot(v, v));
}

#endif // RMD_HELPER_VECTOR_TYPES_CUH
