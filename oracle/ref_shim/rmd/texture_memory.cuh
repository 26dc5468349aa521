// Shadow of the reference's <rmd/texture_memory.cuh> -- TEST INFRASTRUCTURE.
//
// The reference binds legacy texture *references* (texture<T,2>), an API that
// CUDA 12 removed, so its .cu files do not compile with this toolkit as they
// are.  All of them include this header with angle brackets, therefore putting
// this directory before /root/reference/include on the include path swaps the
// legacy binding layer for texture *objects* without touching a reference
// source file.  Same names, same address mode (clamp), same un-normalised
// coordinates, same default filter mode (linear) as
// /root/reference/include/rmd/texture_memory.cuh:27-66, and the same texture
// hardware does the filtering -- the numerics are the reference's.
//
// Only used by oracle/Makefile to build oracle/_ref/librmd_ref*.so.
#ifndef RMD_TEXTURE_MEMORY_SHADOW_CUH
#define RMD_TEXTURE_MEMORY_SHADOW_CUH

#include <cuda_runtime.h>
#include <map>
#include <rmd/device_image.cuh>

namespace rmd
{

template<typename ElementType>
struct TexSlot
{
  cudaTextureObject_t obj;
};

// One slot per legacy texture reference name; `static` gives every
// translation unit (seed_matrix.cu, depthmap_denoiser.cu) its own set, just as
// every TU had its own texture references.
static __device__ TexSlot<float>  ref_img_tex;
static __device__ TexSlot<float>  curr_img_tex;
static __device__ TexSlot<float>  mu_tex;
static __device__ TexSlot<float>  sigma_tex;
static __device__ TexSlot<float>  a_tex;
static __device__ TexSlot<float>  b_tex;
static __device__ TexSlot<int>    convergence_tex;
static __device__ TexSlot<float2> epipolar_matches_tex;
static __device__ TexSlot<float>  g_tex;
static __device__ TexSlot<float>  sum_templ_tex;
static __device__ TexSlot<float>  const_templ_denom_tex;

// tex2D(name, x, y) as the kernels spell it.
template<typename ElementType>
__device__ __forceinline__
ElementType tex2D(const TexSlot<ElementType> &slot, float x, float y)
{
  return ::tex2D<ElementType>(slot.obj, x, y);
}

namespace shadow
{
struct Binding
{
  const void *data;
  size_t width, height, pitch;
  int filter;
  cudaTextureObject_t obj;
};
inline std::map<const void*, Binding> & bindings()
{
  static std::map<const void*, Binding> table;
  return table;
}
}

template<typename ElementType>
inline void bindTexture(
    TexSlot<ElementType> &tex,
    const DeviceImage<ElementType> &mem,
    cudaTextureFilterMode filter_mode=cudaFilterModeLinear)
{
  std::map<const void*, shadow::Binding> &table = shadow::bindings();
  std::map<const void*, shadow::Binding>::iterator it = table.find(&tex);
  if(it != table.end()
     && it->second.data == mem.data && it->second.width == mem.width
     && it->second.height == mem.height && it->second.pitch == mem.pitch
     && it->second.filter == static_cast<int>(filter_mode))
  {
    return; // re-binding the same image: as cheap as the legacy call was
  }

  cudaResourceDesc res_desc;
  memset(&res_desc, 0, sizeof(res_desc));
  res_desc.resType = cudaResourceTypePitch2D;
  res_desc.res.pitch2D.devPtr = mem.data;
  res_desc.res.pitch2D.desc = mem.getCudaChannelFormatDesc();
  res_desc.res.pitch2D.width = mem.width;
  res_desc.res.pitch2D.height = mem.height;
  res_desc.res.pitch2D.pitchInBytes = mem.pitch;

  cudaTextureDesc tex_desc;
  memset(&tex_desc, 0, sizeof(tex_desc));
  tex_desc.addressMode[0] = cudaAddressModeClamp;
  tex_desc.addressMode[1] = cudaAddressModeClamp;
  tex_desc.filterMode = filter_mode;
  tex_desc.readMode = cudaReadModeElementType;
  tex_desc.normalizedCoords = 0;

  TexSlot<ElementType> host_slot;
  cudaError err = cudaCreateTextureObject(&host_slot.obj, &res_desc, &tex_desc, NULL);
  if(err != cudaSuccess)
    throw CudaException("Unable to create texture object: ", err);
  // Blocking copy on the legacy default stream: earlier kernels are done
  // before the slot changes and before the previous object is destroyed.
  err = cudaMemcpyToSymbol(tex, &host_slot, sizeof(host_slot));
  if(err != cudaSuccess)
    throw CudaException("Unable to bind texture: ", err);
  if(it != table.end())
    cudaDestroyTextureObject(it->second.obj);

  shadow::Binding b;
  b.data = mem.data; b.width = mem.width; b.height = mem.height; b.pitch = mem.pitch;
  b.filter = static_cast<int>(filter_mode); b.obj = host_slot.obj;
  table[&tex] = b;
}

} // rmd namespace

#endif
