// pyr_math.cuh -- the per-pixel arithmetic of MakePyramid shared by the pyramid kernels.
// Operation order and FMA placement are those of oracle/oracle.c (pinned bit-for-bit against OpenCV 4.13):
// cv::GaussianBlur 5x5 (separable, rows then columns) and cv::pyrDown ([1 4 6 4 1]/16 per axis, /256 once).
// Every formula is symmetric under reversing its five taps, which is what lets the streaming kernels
// extend an image by reflection instead of special-casing BORDER_REFLECT_101.
#pragma once
#include "sfe_common.cuh"

namespace {

struct Taps { float k0, k1, k2; };

// cv::getGaussianKernel(5, sigma, CV_32F) bit patterns (oracle.c gauss_taps)
__host__ __device__ inline Taps taps_for(int which) {
  Taps t;
  if (which == 0) { t.k0 = __builtin_bit_cast(float, 0x3ebd3532u); t.k1 = __builtin_bit_cast(float, 0x3e7a53d4u); t.k2 = __builtin_bit_cast(float, 0x3d90edf6u); }       // 1.1
  else if (which == 1) { t.k0 = __builtin_bit_cast(float, 0x3eff8c30u); t.k1 = __builtin_bit_cast(float, 0x3e69ff17u); t.k2 = __builtin_bit_cast(float, 0x3cb3a5ccu); }  // 0.8
  else { t.k0 = __builtin_bit_cast(float, 0x3f29efffu); t.k1 = __builtin_bit_cast(float, 0x3e297f46u); t.k2 = __builtin_bit_cast(float, 0x3b282ed8u); }                  // 0.6
  return t;
}

__device__ __forceinline__ float blur_row(float m2, float m1, float c, float p1, float p2, Taps t) {
  float r = (m1 + p1) * t.k1;
  r = fmaf(t.k0, c, r);
  return fmaf(t.k2, m2 + p2, r);
}
__device__ __forceinline__ float blur_col(float m2, float m1, float c, float p1, float p2, Taps t) {
  float r = c * t.k0;
  r = fmaf(t.k1, m1 + p1, r);
  return fmaf(t.k2, m2 + p2, r);
}

// packed forms (two neighbouring columns per instruction), operation for operation the scalar ones
__device__ __forceinline__ float2 blur_col2(float2 m2, float2 m1, float2 c, float2 p1, float2 p2, Taps t) {
  float2 r = mul2(c, both(t.k0));
  r = fma2(both(t.k1), add2(m1, p1), r);
  return fma2(both(t.k2), add2(m2, p2), r);
}
__device__ __forceinline__ float2 pd_v2(float2 r0, float2 r1, float2 r2, float2 r3, float2 r4) {
  return mul2(add2(mul2(add2(add2(r1, r3), r2), both(4.f)), add2(add2(r0, r4), add2(r2, r2))), both(1.f / 256.f));
}

__device__ __forceinline__ float pd_h(float m2, float m1, float c, float p1, float p2) {
  return ((m2 + p2) + (m1 + p1) * 4.f) + c * 6.f;
}
__device__ __forceinline__ float pd_v(float r0, float r1, float r2, float r3, float r4) {
  return (((r1 + r3) + r2) * 4.f + ((r0 + r4) + (r2 + r2))) * (1.f / 256.f);
}

}  // namespace
