// pyr_math.cuh -- the per-pixel arithmetic of MakePyramid shared by the pyramid kernels.
// Operation order and FMA placement are those of oracle/oracle.c (pinned bit-for-bit against OpenCV 4.13):
// cv::GaussianBlur 5x5 (separable, rows then columns) and cv::pyrDown ([1 4 6 4 1]/16 per axis, /256 once).
//
// OpenCV evaluates each filter with one operation order in its vector bodies and another in its scalar
// paths (row tails, border tables), and which column takes which path is a function of the level's width
// alone (oracle.c orc_gauss5 / orc_pyrdown state the rules, probed against cv2 column by column).  The
// *_tail predicates below are those rules; the *_sc functions are the scalar-path forms.  Results are
// bit-identical to cv2 in every column, not only where the vector body ran.
//
// The vector-body forms and the blur's scalar form are symmetric under reversing their five taps, which is
// what lets the streaming kernels extend an image by reflection instead of special-casing
// BORDER_REFLECT_101; the pyrDown scalar forms are NOT symmetric (((6c + 4(m1+p1)) + m2) + p2), so the
// streaming kernels only take widths for which the vertical pyrDown pass has no scalar columns (w % 8 == 0)
// and apply the horizontal scalar form at real column positions only.
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

// ---- OpenCV's scalar-path forms (no FMA; -fmad=false keeps them unfused) and the columns that take them
__device__ __forceinline__ float blur_sc(float m2, float m1, float c, float p1, float p2, Taps t) {  // rows and columns
  return (t.k0 * c + (m1 + p1) * t.k1) + (m2 + p2) * t.k2;
}
__device__ __forceinline__ float pd_h_sc(float m2, float m1, float c, float p1, float p2) {
  return ((c * 6.f + (m1 + p1) * 4.f) + m2) + p2;
}
__device__ __forceinline__ float pd_v_sc(float r0, float r1, float r2, float r3, float r4) {
  return (((r2 * 6.f + (r1 + r3) * 4.f) + r0) + r4) * (1.f / 256.f);
}
// GaussianBlur of a w-wide level: the row filter is scalar in the last column of an odd width, the column filter
// in columns >= (w/8)*8
__host__ __device__ inline bool blur_row_tail(int x, int w) { return (w & 1) && x == w - 1; }
__host__ __device__ inline bool blur_col_tail(int x, int w) { return x >= (w & ~7); }
// pyrDown of a w-wide level into dw = (w+1)/2 columns: the horizontal pass runs its vector body on output columns
// 1..pd_hbody(w), the border table / scalar loop elsewhere; the vertical pass is scalar in columns >= (dw/4)*4
__host__ __device__ inline int pd_hbody(int w) {
  const int dw = (w + 1) / 2;
  int width0 = w < 3 ? 0 : (w - 3) / 2 + 1;
  width0 = width0 < dw ? width0 : dw;
  return width0 >= 1 ? ((width0 - 1) / 4) * 4 : 0;
}
__host__ __device__ inline bool pd_h_tail(int x1, int hbody) { return x1 < 1 || x1 > hbody; }
__host__ __device__ inline bool pd_v_tail(int x1, int dw) { return x1 >= (dw & ~3); }

}  // namespace
