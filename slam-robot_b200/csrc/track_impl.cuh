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
// integer taps, so every lane loads the 4 taps of its patch pixels once (24 registers) and a
// compact runtime loop over the six shifts evaluates weights, patch statistics and score.
//
// Code size matters as much as instruction count here: every warp sits at a different point of
// a long dependent chain, so the kernel body must stay inside the instruction cache (the first
// version, fully unrolled over shifts and inlined twice for forward/backward, was 157 KB of SASS
// and spent >90% of its issue slots waiting for instructions; profiles/README.md).  Hence: one
// direction loop, one level loop, one shift loop, lane-parallel geometry (lane j evaluates axis
// variant j) and lane-parallel finite differences (lane j evaluates quotient j, so the thirteen
// IEEE double divisions of BruteHessian cost three division sequences per warp).
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
#ifndef TRK_MINB
#define TRK_MINB 4  // resident CTAs per SM the register allocator must allow (4 -> <=128 registers)
#endif
constexpr int TS = 16;  // tile row stride (floats)

// per-warp shared scratch
struct WarpScratch {
  float tile[16 * TS];
  float zero[TS + 2];       // taps of unused patch entries point here (must directly follow tile)
  float a[2][3], a1[2][3];  // [axis][variant] fractional weights
  int i0[2][3], r[2][3];    // [axis][variant] integer origin / clipped entries
  float score[8];
};

struct TileFetch {
  const float* t;
  int ox, oy;
  __device__ __forceinline__ float operator()(int Y, int X) const { return t[(Y - oy) * TS + (X - ox)]; }
};

// shift s of BruteHessian uses x-variant (SXP >> 2s) & 3 and y-variant (SYP >> 2s) & 3
//   hessian variants: 0 -> p, 1 -> p-h, 2 -> p+h;  shifts (0,0) (-h,0) (0,-h) (+h,0) (0,+h) (+h,+h)
//   klt     variants: 0 -> p, 1 -> p+h, 2 -> p+2h; shifts (0,0) (h,0) (0,h) (2h,0) (0,2h) (h,h)
template <int MODE>
struct Shifts {
  static constexpr unsigned SXP = MODE == MODE_HESSIAN ? (0u | 1u << 2 | 0u << 4 | 2u << 6 | 0u << 8 | 2u << 10)
                                                       : (0u | 1u << 2 | 0u << 4 | 2u << 6 | 0u << 8 | 1u << 10);
  static constexpr unsigned SYP = MODE == MODE_HESSIAN ? (0u | 0u << 2 | 1u << 4 | 0u << 6 | 2u << 8 | 2u << 10)
                                                       : (0u | 0u << 2 | 1u << 4 | 0u << 6 | 2u << 8 | 1u << 10);
};

// Finite differences of the six scores, lane-parallel.  Stage A: lane j < NA forms
// q_j = [0.5 *] (score[P_j] - score[M_j]) / h; stage B: lane j < 4 forms (q[PB_j] - q[MB_j]) / h.
// All arithmetic is IEEE double, operation for operation what hessian.h:163-169 / klt.h:197-203
// evaluate.  Returns d[6] = dx,dy,dxx,dxy,dyx,dyy in every lane.
// `sc` holds score s in lanes 4s..4s+3 (the layout packed_reduce8 leaves).
template <int MODE>
__device__ __forceinline__ void finite_differences(float sc, int lane, float (&d)[6]) {
  // nibble tables indexed by lane
  constexpr unsigned PA = MODE == MODE_HESSIAN ? 0x55040343u : 0x00554321u;  // P_j (j = 0 is the low nibble)
  constexpr unsigned MA = MODE == MODE_HESSIAN ? 0x34201021u : 0x00122100u;  // M_j
  constexpr unsigned PB = MODE == MODE_HESSIAN ? 0x7642u : 0x5432u;          // dxx,dyy,dxy,dyx minuend lane
  constexpr unsigned MB = MODE == MODE_HESSIAN ? 0x4253u : 0x1010u;          // subtrahend lane
  const double h = MODE == MODE_HESSIAN ? 0.02 : 0.01;
  const int j = lane & 7;
  const double p = (double)__shfl_sync(SFE_FULL, sc, 4 * ((PA >> (4 * j)) & 7));
  const double m = (double)__shfl_sync(SFE_FULL, sc, 4 * ((MA >> (4 * j)) & 7));
  double num = __dsub_rn(p, m);
  if (MODE == MODE_HESSIAN && j < 2) num = __dmul_rn(0.5, num);
  const double q = __ddiv_rn(num, h);
  const int jb = lane & 3;
  const double qp = __shfl_sync(SFE_FULL, q, (PB >> (4 * jb)) & 7), qm = __shfl_sync(SFE_FULL, q, (MB >> (4 * jb)) & 7);
  const double r = __ddiv_rn(__dsub_rn(qp, qm), h);
  const float qf = (float)q, rf = (float)r;
  d[0] = __shfl_sync(SFE_FULL, qf, 0);
  d[1] = __shfl_sync(SFE_FULL, qf, 1);
  d[2] = __shfl_sync(SFE_FULL, rf, 0);
  d[5] = __shfl_sync(SFE_FULL, rf, 1);
  d[3] = __shfl_sync(SFE_FULL, rf, 2);
  d[4] = __shfl_sync(SFE_FULL, rf, 3);
}

// One patch evaluation at (x,y) of image `im`, two kinds sharing all of the sampling code:
//   is_tmpl   GetPatch (hessian.h:54-93 / klt.h:59-96): T[] <- the patch, tmean/tsumsq <- its statistics
//   otherwise BruteHessian (hessian.h:147-172 / klt.h:181-204) against the template T: the six
//             derivatives d[6] = dx,dy,dxx,dxy,dyx,dyy (rounded to float as the reference stores them
//             through float*); returns sad0.
template <int MODE>
__device__ __forceinline__ float evaluate(WarpScratch& S, const ImgView im, bool is_tmpl, float (&T)[SFE_SLOTS],
                                          float& tmean, float& tsumsq, const float (&mk)[SFE_SLOTS], float x, float y,
                                          int lane, float (&d)[6]) {
  constexpr bool CLIP = MODE == MODE_HESSIAN;
  __syncwarp();
  // ---- geometry, lane-parallel: lane j < 3 evaluates x-variant j and y-variant j.  Shifted
  // coordinates are formed in double and rounded to float (cv::Point2f(pt.x - h, pt.y)).
  if (lane < 3) {
    const double off = MODE == MODE_HESSIAN ? (lane == 0 ? 0.0 : (lane == 1 ? -0.02 : 0.02))
                                            : (lane == 0 ? 0.0 : (lane == 1 ? 0.01 : 0.02));
    const float xs = lane == 0 ? x : (float)((double)x + off), ys = lane == 0 ? y : (float)((double)y + off);
    const AxisGeom gx = axis_geom(xs, CLIP, true), gy = axis_geom(ys, CLIP, false);
    S.a[0][lane] = gx.a; S.a1[0][lane] = gx.a1; S.i0[0][lane] = gx.i0; S.r[0][lane] = gx.r;
    S.a[1][lane] = gy.a; S.a1[1][lane] = gy.a1; S.i0[1][lane] = gy.i0; S.r[1][lane] = gy.r;
  }
  // ---- stage the 16x16 footprint (replicate-clamped)
  const int ox = (int)floorf(x) - 7, oy = (int)floorf(y) - 7;
  {
    const int col = clampi(ox + (lane & 15), 0, im.w - 1);
    const int rsub = lane >> 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int row = clampi(oy + 2 * j + rsub, 0, im.h - 1);
      S.tile[(2 * j + rsub) * TS + (lane & 15)] = __ldg(im.p + (long long)row * im.pitch + col);
    }
  }
  __syncwarp();
  const int ix = S.i0[0][0], iy = S.i0[1][0];
  const bool same = S.i0[0][1] == ix && S.i0[0][2] == ix && S.i0[1][1] == iy && S.i0[1][2] == iy;
  const bool noclip = (S.r[0][0] | S.r[0][1] | S.r[0][2] | S.r[1][0] | S.r[1][1] | S.r[1][2]) == 0;
  const bool interior = ix >= 0 && ix + SFE_PATCH <= im.w - 1 && iy >= 0 && iy + SFE_PATCH <= im.h - 1;
  const bool fast = same && noclip && interior;

  float sc;  // score s in lanes 4s..4s+3
  if (fast) {
    // ---- fast path, fully unrolled: every patch pixel reads its 4 taps once; six weight sets
    float v[6][SFE_SLOTS];
    {
      float t00[SFE_SLOTS], t01[SFE_SLOTS], t10[SFE_SLOTS], t11[SFE_SLOTS];
      const float* t0 = S.tile + (iy - oy) * TS + (ix - ox);
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) {
        const int i = lane + 32 * k;
        const int pr = i / SFE_PATCH, pc = i - pr * SFE_PATCH;
        const bool valid = k < SFE_SLOTS - 1 || i < SFE_PLEN;
        const float* t = t0 + (valid ? pr * TS + pc : 0);
        t00[k] = valid ? t[0] : 0.f;
        t01[k] = valid ? t[1] : 0.f;
        t10[k] = valid ? t[TS] : 0.f;
        t11[k] = valid ? t[TS + 1] : 0.f;
      }
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const int jx = (Shifts<MODE>::SXP >> (2 * s)) & 3, jy = (Shifts<MODE>::SYP >> (2 * s)) & 3;
        const float ax = S.a[0][jx], ax1 = S.a1[0][jx], ay = S.a[1][jy], ay1 = S.a1[1][jy];
        const float w0 = ax1 * ay1, w1 = ax * ay1, w2 = ax1 * ay, w3 = ax * ay;
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) v[s][k] = fmaf(t11[k], w3, fmaf(t10[k], w2, fmaf(t01[k], w1, t00[k] * w0)));
      }
    }
    if (MODE == MODE_KLT && is_tmpl) {  // klt.h scoring does not use the template statistics
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) T[k] = v[0][k];
      return 0.f;
    }
    float part[8];
    part[6] = part[7] = 0.f;
    if (MODE == MODE_HESSIAN) {
      // patch statistics of the six candidates (hessian.h:85-91): 12 sums in one packed reduction
      float st[16];
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        float sm = 0.f, sq = 0.f;
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {
          sm = sm + v[s][k];
          sq = fmaf(v[s][k], v[s][k], sq);
        }
        st[s] = sm;
        st[8 + s] = sq;
      }
      st[6] = st[7] = st[14] = st[15] = 0.f;
      const float red = packed_reduce16(st, lane);  // lanes 2s,2s+1: sum_s; lanes 16+2s,17+2s: sumsq_s
      const float other = __shfl_xor_sync(SFE_FULL, red, 16);
      // lane-parallel alpha/beta (hessian.h:131-132): lanes 2s (s < 6) hold the values of shift s
      // (lanes >= 12 hold padding: give them benign operands so the IEEE divide/sqrt stay on their fast paths)
      const bool live = lane < 12;
      const float mean = red / (float)SFE_PLEN, sumsq = live ? other / (float)SFE_PLEN : 1.f;
      if (is_tmpl) {  // shift 0 is the patch itself: lane 0 holds its statistics
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) T[k] = v[0][k];
        tmean = __shfl_sync(SFE_FULL, mean, 0);
        tsumsq = __shfl_sync(SFE_FULL, sumsq, 0);
        return 0.f;
      }
      const float alpha_l = sqrtf((live ? tsumsq : 1.f) / sumsq);
      const float beta_l = tmean - alpha_l * mean;
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const float alpha = __shfl_sync(SFE_FULL, alpha_l, 2 * s), beta = __shfl_sync(SFE_FULL, beta_l, 2 * s);
        float p = 0.f;
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {  // hessian.h:133-139
          float diff = fmaf(-v[s][k], alpha, T[k]) - beta;
          diff = diff * diff;
          const float t = fmaf(diff, mk[k], p);
          p = (T[k] == 0.f || v[s][k] == 0.f) ? p : t;
        }
        part[s] = p;
      }
    } else {
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        float p = 0.f;
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {  // klt.h:141-147
          const float diff = T[k] - v[s][k];
          const float t = fmaf(diff * diff, mk[k], p);
          p = (T[k] == 0.f || v[s][k] == 0.f) ? p : t;
        }
        part[s] = p;
      }
    }
    sc = packed_reduce8(part, lane);
  } else {
    // ---- border path (image borders, GetPatch clipping, or a shift crossing an integer boundary).
    // cv::getRectSubPix's border rules are per pixel: "full" 4-tap, "vertical" 2-tap (overflow
    // column, and corners), "horizontal" 2-tap (overflow row).  With the unused taps zeroed and
    // A1' = xin ? 1-a : 1, B1' = vert ? 1-b : 1 the weights (A1'B1', aB1', A1'b, ab) reproduce the
    // rule exactly (x*1 is exact and a zero tap adds an exact zero), so one FMA chain serves all
    // pixels.  Per patch pixel: one tile offset (entries that are clipped or beyond 169 point into the
    // zero region behind the tile) and two bit masks; a compact runtime loop over the shifts.
    const bool shared_geom = same && S.r[0][1] == S.r[0][0] && S.r[0][2] == S.r[0][0] && S.r[1][1] == S.r[1][0] &&
                             S.r[1][2] == S.r[1][0];
    int toff[SFE_SLOTS];
    unsigned mx[SFE_SLOTS], mv[SFE_SLOTS];  // all-ones where the pixel uses x weights / vertical weights
    const int nshift = is_tmpl ? 1 : 6;
#pragma unroll 1
    for (int s = 0; s < nshift; ++s) {
      const int jx = (Shifts<MODE>::SXP >> (2 * s)) & 3, jy = (Shifts<MODE>::SYP >> (2 * s)) & 3;
      if (s == 0 || !shared_geom) {
        const int x0 = S.i0[0][jx], rx = S.r[0][jx], y0 = S.i0[1][jy], ry = S.r[1][jy];
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {
          const int i = lane + 32 * k;
          const int pr = i / SFE_PATCH, pc = i - pr * SFE_PATCH;
          const bool valid = (k < SFE_SLOTS - 1 || i < SFE_PLEN) && pr >= ry && pc >= rx;
          const int X = x0 + pc, Y = y0 + pr;
          const bool xin = X >= 0 && X + 1 <= im.w - 1, yin = Y >= 0 && Y + 1 <= im.h - 1;
          int Xq = X;
          if (!xin && !yin && Y < 0 && X >= im.w - 1 && im.w >= 2) Xq = im.w - 2;  // OpenCV top-right quirk
          // the tile spans [ox, ox+15] x [oy, oy+15], already replicate-clamped; +1 / +TS stay inside
          toff[k] = valid ? clampi(Y - oy, 0, 14) * TS + clampi(Xq - ox, 0, 14) : 16 * TS;
          mx[k] = xin ? 0xffffffffu : 0u;
          mv[k] = (yin || !xin) ? 0xffffffffu : 0u;
        }
      }
      const unsigned ax = __float_as_uint(S.a[0][jx]), ax1 = __float_as_uint(S.a1[0][jx]);
      const unsigned ay = __float_as_uint(S.a[1][jy]), ay1 = __float_as_uint(S.a1[1][jy]);
      const float axf = __uint_as_float(ax), ayf = __uint_as_float(ay);
      const unsigned one = 0x3f800000u;
      float v[SFE_SLOTS];
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) {
        const float* t = S.tile + toff[k];
        const unsigned r00 = __float_as_uint(t[0]), r01 = __float_as_uint(t[1]);
        const unsigned r10 = __float_as_uint(t[TS]), r11 = __float_as_uint(t[TS + 1]);
        const float t00 = __uint_as_float(r00), t01 = __uint_as_float(r01 & mx[k]);
        const float t10 = __uint_as_float(r10 & mv[k]), t11 = __uint_as_float(r11 & mx[k] & mv[k]);
        const float A1 = __uint_as_float((ax1 & mx[k]) | (one & ~mx[k])), B1 = __uint_as_float((ay1 & mv[k]) | (one & ~mv[k]));
        v[k] = fmaf(t11, axf * ayf, fmaf(t10, A1 * ayf, fmaf(t01, axf * B1, t00 * (A1 * B1))));
      }
      float part = 0.f;
      if (MODE == MODE_HESSIAN) {
        float m, q;
        patch_stats(v, m, q);
        if (is_tmpl) {
#pragma unroll
          for (int k = 0; k < SFE_SLOTS; ++k) T[k] = v[k];
          tmean = m;
          tsumsq = q;
          return 0.f;
        }
        const float alpha = sqrtf(tsumsq / q);
        const float beta = tmean - alpha * m;
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {
          float diff = fmaf(-v[k], alpha, T[k]) - beta;
          diff = diff * diff;
          const float t = fmaf(diff, mk[k], part);
          part = (T[k] == 0.f || v[k] == 0.f) ? part : t;
        }
      } else {
        if (is_tmpl) {
#pragma unroll
          for (int k = 0; k < SFE_SLOTS; ++k) T[k] = v[k];
          return 0.f;
        }
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {
          const float diff = T[k] - v[k];
          const float t = fmaf(diff * diff, mk[k], part);
          part = (T[k] == 0.f || v[k] == 0.f) ? part : t;
        }
      }
      const float tot = warp_sum(part);
      if (lane == 0) S.score[s] = tot;
    }
    __syncwarp();
    sc = S.score[(lane >> 2) & 7];
  }
  finite_differences<MODE>(sc, lane, d);
  return __shfl_sync(SFE_FULL, sc, 0);
}

__device__ __forceinline__ void load_mask(const float* __restrict__ mask, int lane, float (&mk)[SFE_SLOTS]) {
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) mk[k] = (mask && lane + 32 * k < SFE_PLEN) ? __ldg(mask + lane + 32 * k) : 0.f;
}

__device__ __forceinline__ void init_scratch(WarpScratch& S, int lane) {
  if (lane < TS + 2) S.zero[lane] = 0.f;
  __syncwarp();
}

// GetPatches (hessian.h:175-183 / klt.h:249-256) on the template pyramid + TrackFeature
// (hessian.h:243-264 / klt.h:403-424) with Track (hessian.h:185-241 / klt.h:258-401) on the search
// pyramid.  (x,y) is updated only on success.
template <int MODE>
__device__ __forceinline__ int track_feature(WarpScratch& S, const PyrView& tp, int tframe, float tx, float ty,
                                             const PyrView& sp, int sframe, int levels, float thr, int maxit,
                                             const float (&mk)[SFE_SLOTS], float& x, float& y, int lane, int& steps) {
  int lv = min(tp.depth, sp.depth);
  if (MODE == MODE_HESSIAN) lv = min(lv, levels);
  const float margin = MODE == MODE_HESSIAN ? 0.01f : 0.1f;
  float px = x * (float)(1. / (1 << (lv - 1))), py = y * (float)(1. / (1 << (lv - 1)));
#pragma unroll 1
  for (int i = lv - 1; i >= 0; --i) {
    const float sc = (float)(1. / (1 << i));  // pt *= 0.5 i times (exact)
    float T[SFE_SLOTS], tmean = 0.f, tsumsq = 0.f;
    const float th = (MODE == MODE_KLT && i > 0) ? thr * 50.f : thr;  // klt.h:413
    const ImgView tim = img_of(tp, 0, i, tframe), sim = img_of(sp, 0, i, sframe);
    // it == -1 extracts the template patch of this level (GetPatches); it >= 0 are the Newton steps
#pragma unroll 1
    for (int it = -1; it < maxit; ++it) {
      const bool is_tmpl = it < 0;
      if (!is_tmpl && (px < margin || py < margin || (px + margin) > (float)sim.w || (py + margin) > (float)sim.h))
        return SFE_OUT_OF_BOUNDS;
      ImgView im;
      im.p = is_tmpl ? tim.p : sim.p;
      im.w = sim.w; im.h = sim.h; im.pitch = sim.pitch;  // both pyramids have the same geometry
      float d[6];
      evaluate<MODE>(S, im, is_tmpl, T, tmean, tsumsq, mk, is_tmpl ? tx * sc : px, is_tmpl ? ty * sc : py, lane, d);
      if (is_tmpl) continue;
      ++steps;
      float dx, dy;
      newton_step(d[0], d[1], d[2], d[3], d[4], d[5], dx, dy);
      px += clamp1(dx);
      py += clamp1(dy);
      if (MODE == MODE_HESSIAN) {
        if (fabsf(dx) < th && fabsf(dy) < th) break;
      } else {  // klt.h:392: float |d| against the double threshold/10.
        const double t10 = __ddiv_rn((double)th, 10.0);
        if ((double)fabsf(dx) < t10 && (double)fabsf(dy) < t10) break;
      }
    }
    if (i > 0) { px *= 2.f; py *= 2.f; }
  }
  x = px;
  y = py;
  return SFE_OK;
}

template <int MODE>
// Persistent: the grid is sized to the machine (TRK_MINB CTAs per SM) and every warp pulls the next
// feature from a global counter, so no warp slot idles while a CTA-mate finishes a feature that needs
// more Newton steps (features take 8..80 steps; with static assignment a quarter of the slots idled).
__global__ void __launch_bounds__(32 * TRK_WARPS, TRK_MINB) track_fb_kernel(PyrView from, PyrView to, TrackArgs a,
                                                                           const float* __restrict__ mask,
                                                                           int* __restrict__ next_feature) {
  __shared__ WarpScratch scratch[TRK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpScratch& S = scratch[warp];
  init_scratch(S, lane);
  float mk[SFE_SLOTS];
  load_mask(mask, lane, mk);
#pragma unroll 1
  for (;;) {
  int i = 0;
  if (lane == 0) i = atomicAdd(next_feature, 1);
  i = __shfl_sync(SFE_FULL, i, 0);
  if (i >= a.n) break;
  const int pair = i / a.n_per_pair;
  const int ff = a.from_first + pair, tf = a.to_first + pair;
  const float fx = a.from_xy[2 * i], fy = a.from_xy[2 * i + 1];
  float tx = a.to_xy[2 * i], ty = a.to_xy[2 * i + 1];
  const int lv = max(1, a.levels ? a.levels[i] : a.default_levels);  // a level count < 1 is tracked as 1 (host entries reject it)
  int steps = 0;
  int st[2];
  float bx = fx, by = fy;  // matcher.cpp:181
  // dir 0: template from `from` at from_pt, search `to` from the seed (matcher.cpp:175-176)
  // dir 1: template from `to` at the forward result, search `from` from from_pt (matcher.cpp:180-182)
  st[1] = SFE_OK;
#pragma unroll 1
  for (int dir = 0; dir < a.ndir; ++dir) {
    const PyrView& tp = dir == 0 ? from : to;
    const PyrView& sp = dir == 0 ? to : from;
    float x = dir == 0 ? tx : bx, y = dir == 0 ? ty : by;
    const int s = track_feature<MODE>(S, tp, dir == 0 ? ff : tf, dir == 0 ? fx : tx, dir == 0 ? fy : ty, sp,
                                      dir == 0 ? tf : ff, lv, a.thr, a.maxit, mk, x, y, lane, steps);
    if (dir == 0) { tx = x; ty = y; st[0] = s; } else { bx = x; by = y; st[1] = s; }
  }
  bool ok = !(st[0] || st[1]);  // matcher.cpp:192
  if (ok && a.ndir == 2) {
    float ddx = fx - bx, ddy = fy - by;
    double nrm = sqrt(__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)));
    if (nrm > a.fb_max) ok = false;  // matcher.cpp:201
  }
  if (lane == 0) {
    a.to_xy[2 * i] = tx;
    a.to_xy[2 * i + 1] = ty;
    if (a.back_xy) { a.back_xy[2 * i] = bx; a.back_xy[2 * i + 1] = by; }
    if (a.status_fwd) a.status_fwd[i] = st[0];
    if (a.status_bwd) a.status_bwd[i] = st[1];
    if (a.accepted) a.accepted[i] = ok ? 1 : 0;
    if (a.steps) a.steps[i] = steps;
  }
  }
}

template <int MODE>
int launch_track_fb(const PyrView& from, const PyrView& to, const TrackArgs& a, const float* mask, int* counter,
                    int num_sms, cudaStream_t s) {
  if (a.n <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return -(int)e;
  int blocks = min((a.n + TRK_WARPS - 1) / TRK_WARPS, num_sms * TRK_MINB);
  track_fb_kernel<MODE><<<blocks, 32 * TRK_WARPS, 0, s>>>(from, to, a, mask, counter);
  e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace
