// Packed fp32 pairs: Blackwell's fma/add/sub/mul .f32x2 instructions (SASS FFMA2, FADD2, FMUL2) operate on two
// floats held in one 64-bit register.  The two lanes of a packed operation are independent IEEE
// round-to-nearest (flush-to-zero, like the scalar code under -use_fast_math) operations, so a result never
// depends on what it was packed with; the gain is one issue slot for two operations.
#pragma once

namespace rmdb
{

typedef unsigned long long f2;

__device__ __forceinline__ f2 pack(const float lo, const float hi)
{
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 pack(const float2 v) { return pack(v.x, v.y); }
__device__ __forceinline__ float2 unpack(const f2 v)
{
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2 f2_add(const f2 a, const f2 b)
{
  f2 r;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_sub(const f2 a, const f2 b)
{
  f2 r;
  asm("sub.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_mul(const f2 a, const f2 b)
{
  f2 r;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 f2_fma(const f2 a, const f2 b, const f2 c)
{
  f2 r;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

} // namespace rmdb
