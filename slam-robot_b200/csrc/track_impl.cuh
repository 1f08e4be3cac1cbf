// track_impl.cuh -- the Newton patch trackers of the reference on the GPU: HessianTracker
// (hessian.h, MODE_HESSIAN -- the live one) and KLTTracker (klt.h, MODE_KLT), each driven
// forward and backward as matcher.cpp:173-206 does.
//
// One warp per feature; the whole chain (template patches, coarse-to-fine Newton iterations,
// backward track, consistency gate) runs inside one launch with no host round trips.
//
// The hot loop is BruteHessian (hessian.h:147-172 / klt.h:181-204): six 13x13 patches at
// sub-pixel offsets of at most 0.02 px around the current point.  A literal port extracts six
// patches from global memory per Newton step.  Here the six patches share ONE 16x16 footprint
// (origin floor(x)-7: the shifts never leave it), staged once per step in shared memory with
// replicate clamping applied at load time; in the common case all six shifts also share their
// integer taps, so every patch pixel loads its 4 taps once and evaluates six weight sets.
#pragma once
#include "patch.cuh"

namespace {

// What differs between the two trackers (everything else is shared):
//                      MODE_HESSIAN (hessian.h)              MODE_KLT (klt.h)
//  GetPatch            left/top clipping (:63-75)            plain 13x13 (:72-76)
//  score               alpha/beta-normalised, masked (:129)  masked SSD (:139-149)
//  finite differences  central, h = 0.02 (:154-169)          forward, h = 0.01 (:188-203)
//  bounds margin       0.01 (:196)                           0.1 (:272)
//  threshold           thr on every level (:253)             50*thr coarse, thr/10 test (:392,:413)
//  levels              min(levels, depth) (:176,:249)        the whole stack (:409)
enum { MODE_HESSIAN = 0, MODE_KLT = 1 };

constexpr int TRK_WARPS = 4;
constexpr int TS = 16;  // tile row stride (floats)
constexpr int TILE = 16 * TS;

struct TileFetch {
  const float* t;
  int ox, oy;
  __device__ __forceinline__ float operator()(int Y, int X) const { return t[(Y - oy) * TS + (X - ox)]; }
};

// Stage the 16x16 footprint around (x,y) with coordinates clamped to the image (replicate).
__device__ __forceinline__ void load_tile(float* tile, const ImgView& im, int ox, int oy, int lane) {
  const int col = clampi(ox + (lane & 15), 0, im.w - 1);
  const int rsub = lane >> 4;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int row = clampi(oy + 2 * j + rsub, 0, im.h - 1);
    tile[(2 * j + rsub) * TS + (lane & 15)] = __ldg(im.p + (long long)row * im.pitch + col);
  }
}

// per-lane partial of hessian.h:129-141 for one candidate patch
__device__ __forceinline__ float score_partial_hessian(const float (&T)[SFE_SLOTS], const float (&v)[SFE_SLOTS],
                                                       const float (&mk)[SFE_SLOTS], float alpha, float beta) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    float diff = fmaf(-v[k], alpha, T[k]) - beta;
    diff = diff * diff;
    float t = fmaf(diff, mk[k], s);
    s = (T[k] == 0.f || v[k] == 0.f) ? s : t;
  }
  return s;
}

// per-lane partial of klt.h:139-149
__device__ __forceinline__ float score_partial_klt(const float (&T)[SFE_SLOTS], const float (&v)[SFE_SLOTS],
                                                   const float (&mk)[SFE_SLOTS]) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    float diff = T[k] - v[k];
    float t = fmaf(diff * diff, mk[k], s);
    s = (T[k] == 0.f || v[k] == 0.f) ? s : t;
  }
  return s;
}

// BruteHessian at (x,y): the six derivatives d[6] = dx,dy,dxx,dxy,dyx,dyy (rounded to float as
// the reference stores them through float*); returns sad0.
template <int MODE>
__device__ __forceinline__ float brute_hessian(float* tile, const ImgView& im, const float (&T)[SFE_SLOTS],
                                               float tmean, float tsumsq, const float (&mk)[SFE_SLOTS],
                                               const LanePix& lp, float x, float y, int lane, float (&d)[6]) {
  constexpr bool CLIP = MODE == MODE_HESSIAN;
  // shifted coordinates are formed in double and rounded to float (cv::Point2f(pt.x - h, pt.y))
  float xs[3], ys[3];
  xs[0] = x;
  ys[0] = y;
  if (MODE == MODE_HESSIAN) {  // x, x-h, x+h with h = 0.02
    xs[1] = (float)((double)x - 0.02); xs[2] = (float)((double)x + 0.02);
    ys[1] = (float)((double)y - 0.02); ys[2] = (float)((double)y + 0.02);
  } else {                     // x, x+h, x+2h with h = 0.01 (2*h == 0.02 exactly as doubles)
    xs[1] = (float)((double)x + 0.01); xs[2] = (float)((double)x + 0.02);
    ys[1] = (float)((double)y + 0.01); ys[2] = (float)((double)y + 0.02);
  }
  AxisGeom gx[3], gy[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    gx[j] = axis_geom(xs[j], CLIP, true);
    gy[j] = axis_geom(ys[j], CLIP, false);
  }
  // shift s uses x-variant SX[s], y-variant SY[s]
  //   hessian: (0,0) (-h,0) (0,-h) (+h,0) (0,+h) (+h,+h)     klt: (0,0) (h,0) (0,h) (2h,0) (0,2h) (h,h)
  constexpr int SX[6] = {0, 1, 0, 2, 0, MODE == MODE_HESSIAN ? 2 : 1};
  constexpr int SY[6] = {0, 0, 1, 0, 2, MODE == MODE_HESSIAN ? 2 : 1};

  const int ox = (int)floorf(x) - 7, oy = (int)floorf(y) - 7;
  __syncwarp();
  load_tile(tile, im, ox, oy, lane);
  __syncwarp();

  float v[6][SFE_SLOTS];
  const bool same = gx[1].i0 == gx[0].i0 && gx[2].i0 == gx[0].i0 && gy[1].i0 == gy[0].i0 && gy[2].i0 == gy[0].i0;
  const bool noclip = (gx[0].r | gx[1].r | gx[2].r | gy[0].r | gy[1].r | gy[2].r) == 0;
  const bool interior = gx[0].i0 >= 0 && gx[0].i0 + SFE_PATCH <= im.w - 1 && gy[0].i0 >= 0 &&
                        gy[0].i0 + SFE_PATCH <= im.h - 1;
  if (same && noclip && interior) {
    // fast path: every patch pixel reads its 4 taps once; six weight sets
    float w[6][4];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const AxisGeom &ax = gx[SX[s]], &ay = gy[SY[s]];
      w[s][0] = ax.a1 * ay.a1; w[s][1] = ax.a * ay.a1; w[s][2] = ax.a1 * ay.a; w[s][3] = ax.a * ay.a;
    }
    const int base = (gy[0].i0 - oy) * TS + (gx[0].i0 - ox);
#pragma unroll
    for (int k = 0; k < SFE_SLOTS; ++k) {
      if (lp.pr[k] < SFE_PATCH) {
        const float* t = tile + base + lp.pr[k] * TS + lp.pc[k];
        float s00 = t[0], s01 = t[1], s10 = t[TS], s11 = t[TS + 1];
#pragma unroll
        for (int s = 0; s < 6; ++s) v[s][k] = fmaf(s11, w[s][3], fmaf(s10, w[s][2], fmaf(s01, w[s][1], s00 * w[s][0])));
      } else {
#pragma unroll
        for (int s = 0; s < 6; ++s) v[s][k] = 0.f;
      }
    }
  } else {
    TileFetch f{tile, ox, oy};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const AxisGeom ax = gx[SX[s]], ay = gy[SY[s]];
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) {
        bool valid = lp.pr[k] < SFE_PATCH && lp.pr[k] >= ay.r && lp.pc[k] >= ax.r;
        float r = 0.f;
        if (valid) r = sample_general(f, ax.i0 + lp.pc[k], ay.i0 + lp.pr[k], im.w, im.h, ax.a, ax.a1, ay.a, ay.a1);
        v[s][k] = r;
      }
    }
  }

  double sc[6];
#pragma unroll
  for (int s = 0; s < 6; ++s) {
    if (MODE == MODE_HESSIAN) {
      float m, q;
      patch_stats(v[s], m, q);
      float alpha = sqrtf(tsumsq / q);
      float beta = tmean - alpha * m;
      sc[s] = (double)warp_sum(score_partial_hessian(T, v[s], mk, alpha, beta));
    } else {
      sc[s] = (double)warp_sum(score_partial_klt(T, v[s], mk));
    }
  }
  if (MODE == MODE_HESSIAN) {  // hessian.h:154-169
    const double h = 0.02;
    const double sad0 = sc[0], sadn1x = sc[1], sadn1y = sc[2], sadp1x = sc[3], sadp1y = sc[4], sadxy = sc[5];
    const double A = __ddiv_rn(__dsub_rn(sadp1x, sad0), h), B = __ddiv_rn(__dsub_rn(sad0, sadn1x), h);
    const double C = __ddiv_rn(__dsub_rn(sadp1y, sad0), h), D = __ddiv_rn(__dsub_rn(sad0, sadn1y), h);
    d[0] = (float)__ddiv_rn(__dmul_rn(0.5, __dsub_rn(sadp1x, sadn1x)), h);
    d[1] = (float)__ddiv_rn(__dmul_rn(0.5, __dsub_rn(sadp1y, sadn1y)), h);
    d[2] = (float)__ddiv_rn(__dsub_rn(A, B), h);
    d[5] = (float)__ddiv_rn(__dsub_rn(C, D), h);
    d[3] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn(__dsub_rn(sadxy, sadp1y), h), A), h);
    d[4] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn(__dsub_rn(sadxy, sadp1x), h), C), h);
  } else {  // klt.h:188-203
    const double h = 0.01;
    const double sad0 = sc[0], sadx = sc[1], sady = sc[2], sadxx = sc[3], sadyy = sc[4], sadxy = sc[5];
    const double A = __ddiv_rn(__dsub_rn(sadx, sad0), h), C = __ddiv_rn(__dsub_rn(sady, sad0), h);
    d[0] = (float)A;
    d[1] = (float)C;
    d[2] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn(__dsub_rn(sadxx, sadx), h), A), h);
    d[5] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn(__dsub_rn(sadyy, sady), h), C), h);
    d[3] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn(__dsub_rn(sadxy, sady), h), A), h);
    d[4] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn(__dsub_rn(sadxy, sadx), h), C), h);
  }
  return (float)sc[0];
}

// Track (hessian.h:185-241 / klt.h:258-401)
template <int MODE>
__device__ __forceinline__ int track_level(float* tile, const ImgView& im, const float (&T)[SFE_SLOTS], float tmean,
                                           float tsumsq, const float (&mk)[SFE_SLOTS], const LanePix& lp,
                                           float threshold, int maxit, float& x, float& y, int lane, int& steps) {
  const float margin = MODE == MODE_HESSIAN ? 0.01f : 0.1f;
  for (int it = 0; it < maxit; ++it) {
    if (x < margin || y < margin || (x + margin) > (float)im.w || (y + margin) > (float)im.h) return SFE_OUT_OF_BOUNDS;
    float d[6];
    brute_hessian<MODE>(tile, im, T, tmean, tsumsq, mk, lp, x, y, lane, d);
    ++steps;
    float dx, dy;
    newton_step(d[0], d[1], d[2], d[3], d[4], d[5], dx, dy);
    x += clamp1(dx);
    y += clamp1(dy);
    if (MODE == MODE_HESSIAN) {
      if (fabsf(dx) < threshold && fabsf(dy) < threshold) break;
    } else {  // klt.h:392: float |d| against the double threshold/10.
      const double t10 = __ddiv_rn((double)threshold, 10.0);
      if ((double)fabsf(dx) < t10 && (double)fabsf(dy) < t10) break;
    }
  }
  return SFE_OK;
}

template <int MODE>
__device__ __forceinline__ void template_patch(const ImgView& tim, float tx, float ty, const LanePix& lp,
                                               float (&T)[SFE_SLOTS], float& tmean, float& tsumsq) {
  PatchGeom g;
  g.x = axis_geom(tx, MODE == MODE_HESSIAN, true);
  g.y = axis_geom(ty, MODE == MODE_HESSIAN, false);
  sample_patch_global(tim, g, lp, T);
  patch_stats(T, tmean, tsumsq);
}

// GetPatches (hessian.h:175-183 / klt.h:249-256) on the template pyramid + TrackFeature
// (hessian.h:243-264 / klt.h:403-424) on the search pyramid.  (x,y) is updated only on success.
template <int MODE>
__device__ __forceinline__ int track_feature(float* tile, const PyrView& tp, int tframe, float tx, float ty,
                                             const PyrView& sp, int sframe, int levels, float thr, int maxit,
                                             const float (&mk)[SFE_SLOTS], const LanePix& lp, float& x, float& y,
                                             int lane, int& steps) {
  int lv = min(tp.depth, sp.depth);
  if (MODE == MODE_HESSIAN) lv = min(lv, levels);
  float px = x * (float)(1. / (1 << (lv - 1))), py = y * (float)(1. / (1 << (lv - 1)));
  for (int i = lv - 1; i >= 0; --i) {
    const float sc = (float)(1. / (1 << i));  // pt *= 0.5 i times (exact)
    float T[SFE_SLOTS], tmean, tsumsq;
    template_patch<MODE>(img_of(tp, 0, i, tframe), tx * sc, ty * sc, lp, T, tmean, tsumsq);
    const float th = (MODE == MODE_KLT && i > 0) ? thr * 50.f : thr;  // klt.h:413
    int st = track_level<MODE>(tile, img_of(sp, 0, i, sframe), T, tmean, tsumsq, mk, lp, th, maxit, px, py, lane, steps);
    if (st != SFE_OK) return st;
    if (i > 0) { px *= 2.f; py *= 2.f; }
  }
  x = px;
  y = py;
  return SFE_OK;
}

__device__ __forceinline__ void load_mask(const float* __restrict__ mask, int lane, float (&mk)[SFE_SLOTS]) {
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) mk[k] = (lane + 32 * k < SFE_PLEN) ? __ldg(mask + lane + 32 * k) : 0.f;
}

template <int MODE>
__global__ void __launch_bounds__(32 * TRK_WARPS) track_fb_kernel(PyrView from, PyrView to, TrackArgs a,
                                                                  const float* __restrict__ mask) {
  __shared__ float tiles[TRK_WARPS][TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * TRK_WARPS + warp;
  if (i >= a.n) return;
  float* tile = tiles[warp];
  const LanePix lp = lane_pix(lane);
  float mk[SFE_SLOTS];
  load_mask(mask, lane, mk);

  const int pair = i / a.n_per_pair;
  const int ff = a.from_first + pair, tf = a.to_first + pair;
  const float fx = a.from_xy[2 * i], fy = a.from_xy[2 * i + 1];
  float tx = a.to_xy[2 * i], ty = a.to_xy[2 * i + 1];
  const int lv = a.levels ? a.levels[i] : a.default_levels;
  int steps = 0;

  int s1 = track_feature<MODE>(tile, from, ff, fx, fy, to, tf, lv, a.thr, a.maxit, mk, lp, tx, ty, lane, steps);  // :175-176
  float bx = fx, by = fy;                                                                                        // :181
  int s2 = track_feature<MODE>(tile, to, tf, tx, ty, from, ff, lv, a.thr, a.maxit, mk, lp, bx, by, lane, steps);  // :180-182
  bool ok = !(s1 || s2);                                                                                         // :192
  if (ok) {
    float ddx = fx - bx, ddy = fy - by;
    double nrm = sqrt(__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)));
    if (nrm > (double)a.fb_max) ok = false;                                                                      // :201
  }
  if (lane == 0) {
    a.to_xy[2 * i] = tx;
    a.to_xy[2 * i + 1] = ty;
    if (a.back_xy) { a.back_xy[2 * i] = bx; a.back_xy[2 * i + 1] = by; }
    if (a.status_fwd) a.status_fwd[i] = s1;
    if (a.status_bwd) a.status_bwd[i] = s2;
    if (a.accepted) a.accepted[i] = ok ? 1 : 0;
    if (a.steps) a.steps[i] = steps;
  }
}

template <int MODE>
int launch_track_fb(const PyrView& from, const PyrView& to, const TrackArgs& a, const float* mask, cudaStream_t s) {
  if (a.n <= 0) return 0;
  int blocks = (a.n + TRK_WARPS - 1) / TRK_WARPS;
  track_fb_kernel<MODE><<<blocks, 32 * TRK_WARPS, 0, s>>>(from, to, a, mask);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace
