// rmd/matrix.cuh -- fixed-size row-major matrix with the interface of the
// reference's include/rmd/matrix.cuh:28-103 (element (r,c) and flat [i]
// access, stream output, product, 2x2 inverse).  Host-side type: the kernels
// of this implementation take poses as plain float[12].
#ifndef RMD_MATRIX_CUH_
#define RMD_MATRIX_CUH_

#include <cstddef>
#include <iomanip>
#include <ostream>

#ifdef __CUDACC__
#define RMD_HD __host__ __device__
#else
#define RMD_HD
#endif

namespace rmd
{

template<typename Type, unsigned R, unsigned C>
struct Matrix
{
  Type data[R * C];

  RMD_HD Type operator()(int row, int col) const { return data[row * C + col]; }
  RMD_HD Type &operator()(int row, int col) { return data[row * C + col]; }
  RMD_HD Type operator[](int ind) const { return data[ind]; }
  RMD_HD Type &operator[](int ind) { return data[ind]; }

  friend std::ostream &operator<<(std::ostream &out, const Matrix &m)
  {
    for(unsigned r = 0; r < R; ++r)
    {
      for(unsigned c = 0; c < C; ++c)
        out << std::setprecision(9) << m(r, c) << " ";
      out << std::endl;
    }
    return out;
  }
};

template<typename Type, unsigned R, unsigned K, unsigned C>
RMD_HD inline Matrix<Type, R, C> operator*(const Matrix<Type, R, K> &lhs, const Matrix<Type, K, C> &rhs)
{
  Matrix<Type, R, C> out;
  for(unsigned r = 0; r < R; ++r)
    for(unsigned c = 0; c < C; ++c)
    {
      Type acc = 0;
      for(unsigned k = 0; k < K; ++k)
        acc += lhs(r, k) * rhs(k, c);
      out(r, c) = acc;
    }
  return out;
}

template<typename Type>
RMD_HD inline Matrix<Type, 2, 2> inv(const Matrix<Type, 2, 2> &in)
{
  const float det = in[0] * in[3] - in[1] * in[2];
  Matrix<Type, 2, 2> out;
  out[0] = in[3] / det;
  out[1] = -in[1] / det;
  out[2] = -in[2] / det;
  out[3] = in[0] / det;
  return out;
}

} // namespace rmd

#endif // RMD_MATRIX_CUH_
