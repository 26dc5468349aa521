// rmd/se3.cuh -- rmd::SE3<Type> with the reference's interface
// (include/rmd/se3.cuh:27-168): 3x4 row-major [R|t] in `data`, constructors
// from a unit quaternion + translation and from C arrays, inv(), (r,c)
// access, rotate / translate / getTranslation, SE3*SE3 and SE3*float3.
// The floating-point evaluation order of inv() and operator* is the
// reference's, so poses computed by callers (T_world_curr.inv(), ...) are
// bit-identical to what they were.
#ifndef RMD_SE3_CUH_
#define RMD_SE3_CUH_

#include <rmd/helper_vector_types.cuh>
#include <rmd/matrix.cuh>

namespace rmd
{

template<typename Type>
struct SE3
{
  Matrix<Type, 3, 4> data;

  SE3() {}  // uninitialised, like the reference

  // unit quaternion (w, x, y, z) and translation
  SE3(Type qw, Type qx, Type qy, Type qz, Type tx, Type ty, Type tz)
  {
    const Type x = 2 * qx, y = 2 * qy, z = 2 * qz;
    const Type wx = x * qw, wy = y * qw, wz = z * qw;
    const Type xx = x * qx, xy = y * qx, xz = z * qx;
    const Type yy = y * qy, yz = z * qy, zz = z * qz;
    const Type rows[3][4] = {{1 - (yy + zz), xy - wz, xz + wy, tx},
                             {xy + wz, 1 - (xx + zz), yz - wx, ty},
                             {xz - wy, yz + wx, 1 - (xx + yy), tz}};
    for(int r = 0; r < 3; ++r)
      for(int c = 0; c < 4; ++c)
        data(r, c) = rows[r][c];
  }

  // r: rotation matrix, row major (9 values); t: translation (3 values)
  SE3(Type *r, Type *t)
  {
    for(int row = 0; row < 3; ++row)
    {
      for(int col = 0; col < 3; ++col)
        data(row, col) = r[3 * row + col];
      data(row, 3) = t[row];
    }
  }

  SE3<Type> inv() const
  {
    SE3<Type> out;
    for(int row = 0; row < 3; ++row)
    {
      for(int col = 0; col < 3; ++col)
        out.data(row, col) = data(col, row);
      out.data(row, 3) = -data(0, row) * data(0, 3) - data(1, row) * data(1, 3) - data(2, row) * data(2, 3);
    }
    return out;
  }

  Type operator()(int r, int c) const { return data(r, c); }
  Type &operator()(int r, int c) { return data(r, c); }

  float3 rotate(const float3 &p) const
  {
    return make_float3(data(0, 0) * p.x + data(0, 1) * p.y + data(0, 2) * p.z,
                       data(1, 0) * p.x + data(1, 1) * p.y + data(1, 2) * p.z,
                       data(2, 0) * p.x + data(2, 1) * p.y + data(2, 2) * p.z);
  }

  float3 translate(const float3 &p) const
  {
    return make_float3(p.x + data(0, 3), p.y + data(1, 3), p.z + data(2, 3));
  }

  float3 getTranslation() const { return make_float3(data(0, 3), data(1, 3), data(2, 3)); }

  friend std::ostream &operator<<(std::ostream &out, const SE3 &m)
  {
    out << m.data;
    return out;
  }
};

template<typename Type>
inline SE3<Type> operator*(const SE3<Type> &lhs, const SE3<Type> &rhs)
{
  SE3<Type> out;
  for(int row = 0; row < 3; ++row)
  {
    for(int col = 0; col < 3; ++col)
      out.data(row, col) = lhs.data(row, 0) * rhs.data(0, col) + lhs.data(row, 1) * rhs.data(1, col) +
                           lhs.data(row, 2) * rhs.data(2, col);
    out.data(row, 3) = lhs.data(row, 3) + lhs.data(row, 0) * rhs.data(0, 3) +
                       lhs.data(row, 1) * rhs.data(1, 3) + lhs.data(row, 2) * rhs.data(2, 3);
  }
  return out;
}

inline float3 operator*(const SE3<float> &se3, const float3 &p)
{
  return se3.translate(se3.rotate(p));
}

} // namespace rmd

#endif // RMD_SE3_CUH_
