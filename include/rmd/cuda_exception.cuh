// rmd/cuda_exception.cuh -- drop-in for the reference header of the same name
// (include/rmd/cuda_exception.cuh:27-45): same type, same public members
// (what_, err_), thrown by the facade classes whenever the C-ABI returns a
// non-zero code.  Unlike the reference, what() returns storage owned by the
// exception (the reference returns c_str() of a destroyed temporary).
#ifndef RMD_CUDA_EXCEPTION_CUH_
#define RMD_CUDA_EXCEPTION_CUH_

#include <cuda_runtime.h>

#include <exception>
#include <sstream>
#include <string>

namespace rmd
{

struct CudaException : public std::exception
{
  CudaException(const std::string &what, cudaError err)
    : what_(what), err_(err)
  {
    std::ostringstream text;
    text << "CudaException: " << what_ << "\n";
    if(err_ != cudaSuccess)
      text << "cudaError code: " << cudaGetErrorString(err_) << " (" << static_cast<int>(err_) << ")\n";
    text_ = text.str();
  }
  virtual ~CudaException() throw() {}
  virtual const char *what() const throw() { return text_.c_str(); }

  std::string what_;
  cudaError err_;

private:
  std::string text_;
};

namespace detail
{
// Translate a C-ABI return code (include/rmd_b200.h) into the reference's
// error convention.  Own negative codes map to cudaErrorUnknown.
inline void throw_on_error(int code, const char *where);
}

} // namespace rmd

#include <rmd_b200.h>

namespace rmd
{
namespace detail
{
inline void throw_on_error(int code, const char *where)
{
  if(code == 0)
    return;
  const std::string msg = std::string(where) + ": " + rmd_last_error_string();
  throw CudaException(msg, code > 0 ? static_cast<cudaError>(code) : cudaErrorUnknown);
}
}
}

#endif // RMD_CUDA_EXCEPTION_CUH_
