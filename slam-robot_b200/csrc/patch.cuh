// patch.cuh -- warp-cooperative 13x13 patch machinery shared by the trackers.
//
// One warp owns one feature.  Patch pixel i (row-major, 0..168) lives in lane i%32, slot i/32
// (6 slots; slot 5 is populated in lanes 0..8 only).  Reductions over the patch are per-lane
// sequential over the slots followed by a 16,8,4,2,1 xor butterfly -- the order declared in
// oracle/oracle.h, so sums are bit-identical to the CPU oracle.
//
// Sampling restates cv::getRectSubPix (32f) including its border rules (oracle.c
// orc_rect_subpix): interior pixels are one multiply and three FMAs; overflow columns use a
// 2-tap vertical form, overflow rows a 2-tap horizontal form, corners replicate (with OpenCV's
// top-right irregularity).  All cases are expressed as one 4-tap FMA chain with per-pixel
// weights, which is exactly equivalent (a zero weight adds an exact zero).
#pragma once

#include "sfe_common.cuh"

// One axis of a getRectSubPix call: coordinate of patch column c is i0 + c (valid for c >= r).
struct AxisGeom {
  int i0;       // integer source coordinate of patch index 0 (= ip - r)
  int r;        // leading patch entries left at zero (GetPatch clipping, hessian.h:63-75)
  float a, a1;  // fractional weight and 1-a
};

// hessian.h:65-68 (x, round_up = true) / :71-74 (y, round_up = false) followed by the
// getRectSubPix prologue (center -= (n-1)/2; ip = floor; a = frac).  clip=false is the plain
// 13-wide call of klt.h:72-76 / brute.h:43-47.
__device__ __forceinline__ AxisGeom axis_geom(float p, bool clip, bool round_up) {
  AxisGeom g;
  int n = SFE_PATCH;
  g.r = 0;
  if (clip && p < 6.5f) {  // float->double is exact and 6.5 is representable: same compare
    int d = round_up ? (int)((6.5 - (double)p) + 0.9999) : (int)(6.5 - (double)p);
    p = (float)((double)p + 0.5 * d);
    g.r = d;
    n = SFE_PATCH - d;
  }
  float c = p - (float)(n - 1) * 0.5f;
  float fl = floorf(c);
  g.a = c - fl;
  g.a1 = 1.f - g.a;
  g.i0 = (int)fl - g.r;
  return g;
}

struct PatchGeom {
  AxisGeom x, y;
};

// Per-lane patch coordinates (slot k -> row pr[k], col pc[k]); invalid slots get row 99.
struct LanePix {
  int pr[SFE_SLOTS], pc[SFE_SLOTS];
};

__device__ __forceinline__ LanePix lane_pix(int lane) {
  LanePix lp;
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    int i = lane + 32 * k;
    lp.pr[k] = i < SFE_PLEN ? i / SFE_PATCH : 99;
    lp.pc[k] = i % SFE_PATCH;
  }
  return lp;
}

// Generic sample of one patch pixel through a fetch functor f(Y, X) that returns the image value
// at CLAMPED coordinates.  (X,Y) are the unclamped integer coordinates of the top-left tap.
template <class Fetch>
__device__ __forceinline__ float sample_general(const Fetch& f, int X, int Y, int w, int h, float a,
                                                float a1, float b, float b1) {
  bool xin = (X >= 0) && (X + 1 <= w - 1);
  bool yin = (Y >= 0) && (Y + 1 <= h - 1);
  int Xq = X;
  if (!xin && !yin && Y < 0 && X >= w - 1 && w >= 2) Xq = w - 2;  // OpenCV top-right quirk
  float s00 = f(Y, Xq), s01 = f(Y, X + 1), s10 = f(Y + 1, Xq), s11 = f(Y + 1, X + 1);
  float w0, w1, w2, w3;
  if (xin && yin) { w0 = a1 * b1; w1 = a * b1; w2 = a1 * b; w3 = a * b; }
  else if (yin)   { w0 = b1; w1 = 0.f; w2 = b;  w3 = 0.f; }
  else if (xin)   { w0 = a1; w1 = a;   w2 = 0.f; w3 = 0.f; }
  else            { w0 = b1; w1 = 0.f; w2 = b;  w3 = 0.f; }
  return fmaf(s11, w3, fmaf(s10, w2, fmaf(s01, w1, s00 * w0)));
}

struct GlobalFetch {
  ImgView im;
  __device__ __forceinline__ float operator()(int Y, int X) const {
    return __ldg(im.p + (long long)clampi(Y, 0, im.h - 1) * im.pitch + clampi(X, 0, im.w - 1));
  }
};

// GetPatch straight from global memory (used once per level for the template patch).
// v[k] = patch pixel of this lane's slot k (0 where clipped or beyond 169).
__device__ __forceinline__ void sample_patch_global(const ImgView& im, const PatchGeom& g,
                                                    const LanePix& lp, float (&v)[SFE_SLOTS]) {
  GlobalFetch f{im};
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    bool valid = lp.pr[k] < SFE_PATCH && lp.pr[k] >= g.y.r && lp.pc[k] >= g.x.r;
    float s = 0.f;
    if (valid) s = sample_general(f, g.x.i0 + lp.pc[k], g.y.i0 + lp.pr[k], im.w, im.h, g.x.a, g.x.a1, g.y.a, g.y.a1);
    v[k] = s;
  }
}

// mean and sumsq over all 169 entries (hessian.h:85-91), declared order.
__device__ __forceinline__ void patch_stats(const float (&v)[SFE_SLOTS], float& mean, float& sumsq) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    s = s + v[k];
    q = fmaf(v[k], v[k], q);
  }
  mean = warp_sum(s) / (float)SFE_PLEN;
  sumsq = warp_sum(q) / (float)SFE_PLEN;
}

// Packed warp reductions.  Each lane brings N partial sums; after log2(N) exchange stages in
// which every lane hands half of its values to its xor-partner and keeps the other half, then
// plain butterflies, lane l holds the warp total of value (l >> (5 - log2 N)).  Every value is
// summed by exactly the 16,8,4,2,1 pairwise tree of warp_sum() (IEEE addition is commutative, so
// which lane performs a node does not matter): results are bit-identical to N calls of warp_sum()
// at 1/3 of the instructions and 1/N of the dependent shuffle stages.
__device__ __forceinline__ float packed_reduce16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = (h16 ? v[j + 8] : v[j]) + __shfl_xor_sync(SFE_FULL, h16 ? v[j] : v[j + 8], 16);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = (h8 ? a[j + 4] : a[j]) + __shfl_xor_sync(SFE_FULL, h8 ? a[j] : a[j + 4], 8);
#pragma unroll
  for (int j = 0; j < 2; ++j) c[j] = (h4 ? b[j + 2] : b[j]) + __shfl_xor_sync(SFE_FULL, h4 ? b[j] : b[j + 2], 4);
  float d = (h2 ? c[1] : c[0]) + __shfl_xor_sync(SFE_FULL, h2 ? c[0] : c[1], 2);
  return d + __shfl_xor_sync(SFE_FULL, d, 1);  // value (lane >> 1)
}
__device__ __forceinline__ float packed_reduce8(const float (&v)[8], int lane) {
  float b[4], c[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = (h16 ? v[j + 4] : v[j]) + __shfl_xor_sync(SFE_FULL, h16 ? v[j] : v[j + 4], 16);
#pragma unroll
  for (int j = 0; j < 2; ++j) c[j] = (h8 ? b[j + 2] : b[j]) + __shfl_xor_sync(SFE_FULL, h8 ? b[j] : b[j + 2], 8);
  float d = (h4 ? c[1] : c[0]) + __shfl_xor_sync(SFE_FULL, h4 ? c[0] : c[1], 4);
  d = d + __shfl_xor_sync(SFE_FULL, d, 2);
  return d + __shfl_xor_sync(SFE_FULL, d, 1);  // value (lane >> 2)
}

// The Newton update of hessian.h:209-227 / klt.h:360-379 from the six float derivatives.
__device__ __forceinline__ void newton_step(float gdx, float gdy, float dxx, float dxy, float dyx, float dyy,
                                            float& dx, float& dy) {
  double H00 = dxx, H01 = dxy, H10 = dyx, H11 = dyy, g0 = gdx, g1 = gdy;
  double det = __dsub_rn(__dmul_rn(H00, H11), __dmul_rn(H10, H01));
  double invdet = __ddiv_rn(1.0, det);
  double i00 = __dmul_rn(H11, invdet), i10 = __dmul_rn(-H10, invdet);
  double i01 = __dmul_rn(-H01, invdet), i11 = __dmul_rn(H00, invdet);
  double j0 = __dadd_rn(__dmul_rn(i00, g0), __dmul_rn(i01, g1));
  double j1 = __dadd_rn(__dmul_rn(i10, g0), __dmul_rn(i11, g1));
  dx = (float)(-j0);
  dy = (float)(-j1);
  if (dx * dx + dy * dy > 1.f) {
    dx = dx / sqrtf(dx * dx + dy * dy);
    dy = dy / sqrtf(dx * dx + dy * dy);  // uses the UPDATED dx (hessian.h:226 quirk)
  }
}

__device__ __forceinline__ float clamp1(float v) { return fmaxf(-1.f, fminf(1.f, v)); }
