// track_hessian.cu -- P1, the live path: HessianTracker (hessian.h) driven forward and backward as
// matcher.cpp:173-206 does, plus the GetPatch / BruteHessian parity accessors.
//
// One warp per feature, persistent grid with a dynamic feature queue; the whole chain (template
// patches, coarse-to-fine Newton iterations, backward track, consistency gate) runs inside one
// launch with no host round trips.
//
// The hot loop is BruteHessian (hessian.h:147-172): six 13x13 bilinear patches at offsets of
// +-0.02 px around the current point, each scored against the template.  The six patches share one
// 16x16 footprint (origin floor(x)-7), staged once per Newton step in shared memory with replicate
// clamping applied at load time.  Three sampling routes feed ONE common tail (patch statistics ->
// packed warp reduction -> lane-parallel alpha/beta -> scores -> packed reduction -> lane-parallel
// finite differences):
//   fast      all six shifts share their integer taps, no GetPatch clipping, footprint inside the
//             image and strictly positive: every lane loads the 4 taps of its <= 6 patch pixels once
//             and evaluates six weight sets (fully unrolled, patch values stay in registers);
//   plain     as fast, but a shift crosses an integer boundary or the footprint holds a zero: a compact
//             runtime loop over the shifts re-reads the taps and writes the patch values to a per-warp
//             shared scratch; template patches (GetPatch, one shift) take one iteration of it;
//   general   image borders (cv::getRectSubPix's per-pixel 2-tap rules and its top-right
//             irregularity) and GetPatch's left/top clipping, same scratch.
// The exact-zero skip of ScorePatchMatch (hessian.h:134) is folded away where it cannot trigger:
// template zeros become zero mask weights once per level (x*0 adds an exact zero), and candidate
// zeros are impossible when the staged footprint is strictly positive and nothing is clipped; the
// remaining cases run a compact rolled score loop with the select.
//
// Code size is a first-class constraint: every warp sits at a different point of a long dependent
// chain, so the frequently executed body has to stay inside the 32 KB L1.5 instruction cache.  Measured
// (profiles/README.md): +376 SASS instructions across that line cost +14 % run time although they
// removed 5 % of the executed instructions; below the line, fewer executed instructions win again.
#include "patch.cuh"

namespace {

#ifndef TRK_WARPS_D
#define TRK_WARPS_D 4  // warps (= features in flight) per CTA
#endif
constexpr int TRK_WARPS = TRK_WARPS_D;
#ifndef TRK_MINB
#define TRK_MINB 4  // resident CTAs per SM the register allocator must allow (4 -> <=128 registers)
#endif
#ifndef TRK_TS
#define TRK_TS 45
#endif
// Tile row stride (floats).  45 = 13 + 32: patch pixel i = 13 pr + pc of a slot then sits in bank (i + const) mod 32, so the
// four tap loads of the 32 lanes are conflict-free; with a stride of 16 rows r and r + 2 share their banks and every tap load
// takes two wavefronts (the LSU / shared-memory path is this kernel's busiest unit: 66 %, profiles/track_r2e_summary.txt).
constexpr int TS = TRK_TS;
constexpr int ZOFF = 16 * TS;   // taps of unused patch entries point at the zero region behind the tile

struct WarpScratch {
  float tile[16 * TS];
  float zero[TS + 2];                // must directly follow tile
  float pad[32 - (TS + 2) % 32];
  float2 v2[3 * SFE_SLOTS * 32];     // non-fast routes: patch values of the shift pair (2p, 2p+1), slot k, lane l at [(p*6+k)*32 + l]
  float2 TM[SFE_SLOTS * 32];         // {template patch value, its effective mask weight}, slot k of lane l at [k*32 + l]
  // transposing reductions: lane l parks partial sum j at [j * RS + l]; rows 6, 7, 14, 15 of the statistics block and
  // rows 6, 7 of the score block are never written and stay zero (they feed the idle lanes of the read-back)
  float red[16 * 36];                // statistics; the six scores reuse rows 0-5 (a __syncwarp apart)
  // lane-parallel results that every lane needs, handed over with one store and a few broadcast loads instead of one
  // shuffle per value: geo[j] = {a, 1-a, i0, r} of axis variant j (0..2: x, 4..6: y), ab = -alpha (0..5), -beta (8..13)
  float4 geo[8];
  float ab[16];
};
constexpr int RS = 36;  // row stride of the reduction blocks: 16-byte aligned rows, and the float4 read-back of a quarter-warp
                        // (4 rows x 2 halves, or 2 rows x 4 quarters) touches 32 distinct banks

// Per-lane tap offsets, packed: 16-bit field k of (a, b, c) is pr * TS + pc of the lane's slot k (i = lane + 32 k,
// pr = i / 13, pc = i % 13) -- the tap offset inside the footprint.  Three registers for the whole kernel instead of six
// quotients plus the multiply-subtract per slot and evaluation.
struct PixPack {
  unsigned a, b, c;
};
__device__ __forceinline__ PixPack make_pixpack(int lane) {
  unsigned w[3] = {0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    const int i = lane + 32 * k;
    const unsigned f = i < SFE_PLEN ? (unsigned)((i / SFE_PATCH) * TS + i % SFE_PATCH) : 0u;
    w[k >> 1] |= f << (16 * (k & 1));
  }
  return PixPack{w[0], w[1], w[2]};
}
__device__ __forceinline__ int pix_off(const PixPack& p, int k) {   // zero-extended field k: one PRMT
  return (int)__byte_perm(k < 2 ? p.a : (k < 4 ? p.b : p.c), 0u, (k & 1) ? 0x4432u : 0x4410u);
}

// x variants live in lanes 0..2 (p, p-h, p+h), y variants in lanes 4..6; BruteHessian's shifts
// (0,0) (-h,0) (0,-h) (+h,0) (0,+h) (+h,+h) (hessian.h:154-161) pick variant (SXP>>2s)&3 / (SYP>>2s)&3.
constexpr unsigned SXP = 0u | 1u << 2 | 0u << 4 | 2u << 6 | 0u << 8 | 2u << 10;
constexpr unsigned SYP = 0u | 0u << 2 | 1u << 4 | 0u << 6 | 2u << 8 | 2u << 10;

// Lane-parallel geometry of one BruteHessian: lane j evaluates one axis variant.  Shifted coordinates
// are formed in double and rounded to float (cv::Point2f(pt.x - h, pt.y), hessian.h:155-160).
__device__ __forceinline__ AxisGeom lane_geom(float x, float y, int lane) {
  const bool isy = lane & 4;
  const int var = lane & 3;
  const double off = var == 1 ? -0.02 : (var == 2 ? 0.02 : 0.0);
  const float p = isy ? y : x;
  const float ps = (float)((double)p + off);  // + 0.0 is exact
  return axis_geom(ps, true, !isy);
}

// Stage the replicate-clamped 16x16 footprint; returns true when it holds a value <= 0.
__device__ __forceinline__ bool stage_tile(WarpScratch& S, const ImgView& im, int ox, int oy, int lane) {
  const int c = lane & 15, rsub = lane >> 4;
  const float* colp = im.p + clampi(ox + c, 0, im.w - 1);
  float mn = 1.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int row = clampi(oy + 2 * j + rsub, 0, im.h - 1);
    const float val = __ldg(colp + (long long)row * im.pitch);
    S.tile[(2 * j + rsub) * TS + c] = val;
    mn = fminf(mn, val);
  }
  const bool nonpos = __any_sync(SFE_FULL, mn <= 0.f);
  __syncwarp();
  return nonpos;
}

// The six axis variants of one BruteHessian in every lane (read back from WarpScratch::geo).
struct GeoAll {
  float4 x[3], y[3];  // {a, 1-a, i0 bits, r bits}
};
__device__ __forceinline__ int gi(float f) { return __float_as_int(f); }

// General sampling route: `nshift` shifts of the BruteHessian pattern (1 for a template patch),
// cv::getRectSubPix's border rules per pixel.  "full" 4-tap, "vertical" 2-tap (overflow column, and
// corners), "horizontal" 2-tap (overflow row): with the unused taps zeroed and A1' = xin ? 1-a : 1,
// B1' = vert ? 1-b : 1 the weights (A1'B1', aB1', A1'b, ab) reproduce the rule exactly (x*1 is exact
// and a zero tap adds an exact zero), so one FMA chain serves all pixels.  Results go to S.v2.
__device__ __forceinline__ void general_sample(WarpScratch& S, const ImgView& im, int ox, int oy,
                                               int nshift, bool shared_geom, int lane, const PixPack& pix) {
  int toff[SFE_SLOTS];
  unsigned mx[SFE_SLOTS], mv[SFE_SLOTS];  // all-ones where the pixel uses x weights / vertical weights
#pragma unroll 1
  for (int s = 0; s < nshift; ++s) {
    const int jx = (SXP >> (2 * s)) & 3, jy = 4 + ((SYP >> (2 * s)) & 3);
    const float4 gx = S.geo[jx], gy = S.geo[jy];  // broadcast loads
    if (s == 0 || !shared_geom) {
      const int x0 = gi(gx.z), rx = gi(gx.w);
      const int y0 = gi(gy.z), ry = gi(gy.w);
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) {
        const int i = lane + 32 * k;
        const int pr = (int)((unsigned)i / SFE_PATCH), pc = i - SFE_PATCH * pr;
        const bool valid = (k < SFE_SLOTS - 1 || i < SFE_PLEN) && pr >= ry && pc >= rx;
        const int X = x0 + pc, Y = y0 + pr;
        const bool xin = X >= 0 && X + 1 <= im.w - 1, yin = Y >= 0 && Y + 1 <= im.h - 1;
        int Xq = X;
        if (!xin && !yin && Y < 0 && X >= im.w - 1 && im.w >= 2) Xq = im.w - 2;  // OpenCV top-right quirk
        // the tile spans [ox, ox+15] x [oy, oy+15], already replicate-clamped; +1 / +TS stay inside
        toff[k] = valid ? clampi(Y - oy, 0, 14) * TS + clampi(Xq - ox, 0, 14) : ZOFF;
        mx[k] = xin ? 0xffffffffu : 0u;
        mv[k] = (yin || !xin) ? 0xffffffffu : 0u;
      }
    }
    const float axf = gx.x, ayf = gy.x;
    const unsigned ax1 = __float_as_uint(gx.y), ay1 = __float_as_uint(gy.y);
    const unsigned one = 0x3f800000u;
#pragma unroll
    for (int k = 0; k < SFE_SLOTS; ++k) {
      const float* t = S.tile + toff[k];
      const unsigned r01 = __float_as_uint(t[1]), r10 = __float_as_uint(t[TS]), r11 = __float_as_uint(t[TS + 1]);
      const float t00 = t[0], t01 = __uint_as_float(r01 & mx[k]);
      const float t10 = __uint_as_float(r10 & mv[k]), t11 = __uint_as_float(r11 & mx[k] & mv[k]);
      const float A1 = __uint_as_float((ax1 & mx[k]) | (one & ~mx[k])), B1 = __uint_as_float((ay1 & mv[k]) | (one & ~mv[k]));
      reinterpret_cast<float*>(S.v2)[((s >> 1) * SFE_SLOTS + k) * 64 + 2 * lane + (s & 1)] =
          fmaf(t11, axf * ayf, fmaf(t10, A1 * ayf, fmaf(t01, axf * B1, t00 * (A1 * B1))));
    }
  }
}

// General route when all shifts share their integer geometry (the usual case: no +-h shift crosses an integer
// boundary).  Slots outside, shifts inside: a pixel's taps, its border masks and its three A1' / B1' candidates are
// formed ONCE and serve all six shifts, whose weights differ only in which axis variant they take -- 4 loads and the
// border logic per pixel instead of per pixel and shift.  The loop over the slots is rolled (code size), the six
// shifts inside are unrolled with compile-time variant indices.
__device__ __forceinline__ void general_sample_shared(WarpScratch& S, const ImgView& im, const GeoAll& G, int ox, int oy,
                                                      int nshift, int lane, const PixPack& pix) {
  const int x0 = gi(G.x[0].z), rx = gi(G.x[0].w);
  const int y0 = gi(G.y[0].z), ry = gi(G.y[0].w);
  float ax[3], ay[3];
  unsigned ax1[3], ay1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    ax[j] = G.x[j].x;
    ax1[j] = __float_as_uint(G.x[j].y);
    ay[j] = G.y[j].x;
    ay1[j] = __float_as_uint(G.y[j].y);
  }
  // the shifts are evaluated as pairs (2p, 2p+1) with packed FP32 (each half rounds like the scalar form); the a*b
  // weight of a pair does not depend on the pixel
  float2 axp[3], ayp[3], w3p[3];
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const int jxa = (SXP >> (4 * p)) & 3, jya = (SYP >> (4 * p)) & 3, jxb = (SXP >> (4 * p + 2)) & 3, jyb = (SYP >> (4 * p + 2)) & 3;
    axp[p] = make_float2(ax[jxa], ax[jxb]);
    ayp[p] = make_float2(ay[jya], ay[jyb]);
    w3p[p] = mul2(axp[p], ayp[p]);
  }
  const unsigned one = 0x3f800000u;
#pragma unroll 1
  for (int k = 0; k < SFE_SLOTS; ++k) {
    const int i = lane + 32 * k, pr = (int)((unsigned)i / SFE_PATCH), pc = i - SFE_PATCH * pr;
    const bool valid = i < SFE_PLEN && pr >= ry && pc >= rx;
    const int X = x0 + pc, Y = y0 + pr;
    const bool xin = X >= 0 && X + 1 <= im.w - 1, yin = Y >= 0 && Y + 1 <= im.h - 1;
    int Xq = X;
    if (!xin && !yin && Y < 0 && X >= im.w - 1 && im.w >= 2) Xq = im.w - 2;  // OpenCV top-right quirk
    const float* t = S.tile + (valid ? clampi(Y - oy, 0, 14) * TS + clampi(Xq - ox, 0, 14) : ZOFF);
    const unsigned mx = xin ? 0xffffffffu : 0u, mv = (yin || !xin) ? 0xffffffffu : 0u;
    const float t00 = t[0], t01 = __uint_as_float(__float_as_uint(t[1]) & mx);
    const float t10 = __uint_as_float(__float_as_uint(t[TS]) & mv), t11 = __uint_as_float(__float_as_uint(t[TS + 1]) & mx & mv);
    float A1[3], B1[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      A1[j] = __uint_as_float((ax1[j] & mx) | (one & ~mx));
      B1[j] = __uint_as_float((ay1[j] & mv) | (one & ~mv));
    }
    float2* out = S.v2 + k * 32 + lane;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      if (p == 1 && nshift == 1) break;   // template patch: shift 0 only (the pair's other half is not read)
      const int jxa = (SXP >> (4 * p)) & 3, jya = (SYP >> (4 * p)) & 3, jxb = (SXP >> (4 * p + 2)) & 3, jyb = (SYP >> (4 * p + 2)) & 3;
      const float2 A1p = make_float2(A1[jxa], A1[jxb]), B1p = make_float2(B1[jya], B1[jyb]);
      out[p * SFE_SLOTS * 32] = fma2(both(t11), w3p[p], fma2(both(t10), mul2(A1p, ayp[p]), fma2(both(t01), mul2(axp[p], B1p), mul2(both(t00), mul2(A1p, B1p)))));
    }
  }
}

// Plain route: no border rule and no clipping.  A compact runtime loop re-reads the 4 taps per shift, so it
// also serves steps in which a +-h shift crosses an integer boundary (the shifts do not share their taps
// then), template patches (one shift) and footprints that contain a zero.  Results go to S.v2.
__device__ __forceinline__ void straddle_sample(WarpScratch& S, int ox, int oy, int nshift, int lane,
                                                const PixPack& pix) {
  int poff[SFE_SLOTS];
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) poff[k] = (k < SFE_SLOTS - 1 || lane + 32 * k < SFE_PLEN) ? pix_off(pix, k) : -1;
#pragma unroll 1
  for (int s = 0; s < nshift; ++s) {
    const int jx = (SXP >> (2 * s)) & 3, jy = 4 + ((SYP >> (2 * s)) & 3);
    const float4 gx = S.geo[jx], gy = S.geo[jy];  // broadcast loads
    const int base = (gi(gy.z) - oy) * TS + (gi(gx.z) - ox);
    const float ax = gx.x, ax1 = gx.y;
    const float ay = gy.x, ay1 = gy.y;
    const float w0 = ax1 * ay1, w1 = ax * ay1, w2 = ax1 * ay, w3 = ax * ay;
    float* out = reinterpret_cast<float*>(S.v2) + (s >> 1) * SFE_SLOTS * 64 + 2 * lane + (s & 1);  // half s & 1 of pair s >> 1
#pragma unroll
    for (int k = 0; k < SFE_SLOTS; ++k) {
      const float* tp = S.tile + (poff[k] >= 0 ? base + poff[k] : ZOFF);
      out[k * 64] = fmaf(tp[TS + 1], w3, fmaf(tp[TS], w2, fmaf(tp[1], w1, tp[0] * w0)));
    }
  }
}

// Correctly rounded division by a CONSTANT without the division sequence: with y = RN(1/d),
// q = RN(x*y), r = x - q*d (exact in one FMA) and q' = RN(q + r*y), q' equals RN(x/d) (Markstein's
// theorem for a correctly rounded reciprocal; no overflow/underflow at the magnitudes of patch sums and
// scores).  Verified rather than trusted: x/169.f exhaustively over all 2^32 floats (the only difference
// is x = -0, which a sum of non-negative terms never is), x/0.02 over 6.4e9 doubles of the forms that
// occur here (float differences, halves, quotient differences) -- tools/check_const_div.c.  The oracle keeps
// the reference's plain divisions (hessian.h:88-89, :163-169); parity tests compare bit patterns.
__device__ __forceinline__ float div169(float x) {
  const float d = (float)SFE_PLEN, y = 1.0f / (float)SFE_PLEN;
  const float q = x * y;
  const float r = fmaf(-q, d, x);
  return fmaf(r, y, q);
}
__device__ __forceinline__ double div_h(double x) {
  const double h = 0.02, y = 1.0 / 0.02;
  const double q = __dmul_rn(x, y);
  const double r = __fma_rn(-q, h, x);
  return __fma_rn(r, y, q);
}

// Declared reduction order of this tracker (oracle.h, tree32_adj): per-lane partial sums, then the balanced pairwise tree
// that pairs ADJACENT lanes first -- ((l0+l1)+(l2+l3)) ... strides 1,2,4,8,16.  The reference is built with -ffast-math and
// has no order of its own; this one is chosen because it maps onto a transposition through shared memory: every lane
// parks its partial sums (one conflict-free STS each), then a few lanes per value read back 16 (or 8) consecutive lanes'
// partials as float4 and add them as a tree in registers -- 12 STS + 4 LDS.128 + 16 FADD + 1 SHFL for the twelve patch
// statistics, where the transposing shuffle butterfly needed 16 SHFL + 30 selects + 16 FADD (-9 % executed instructions).
__device__ __forceinline__ float tree4(float4 a) { return (a.x + a.y) + (a.z + a.w); }

// 12 statistics (rows 0-5: sums, rows 8-13: sums of squares).  Returns the total of row (lane >> 1).
__device__ __forceinline__ float reduce_stats(WarpScratch& S, const float (&st)[16], int lane) {
  float* w = S.red + lane;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    w[j * RS] = st[j];
    w[(8 + j) * RS] = st[8 + j];
  }
  __syncwarp();
  const float4* r = reinterpret_cast<const float4*>(S.red + (lane >> 1) * RS + (lane & 1) * 16);
  const float4 a = r[0], b = r[1], c = r[2], d = r[3];
  const float h = (tree4(a) + tree4(b)) + (tree4(c) + tree4(d));  // 16 lanes
  return h + __shfl_xor_sync(SFE_FULL, h, 1);
}
// 6 scores.  Returns the total of row (lane >> 2) (lanes 24-31: zero).
__device__ __forceinline__ float reduce_scores(WarpScratch& S, const float (&part)[8], int lane) {
  float* w = S.red + lane;
#pragma unroll
  for (int j = 0; j < 6; ++j) w[j * RS] = part[j];
  __syncwarp();
  const float4* r = reinterpret_cast<const float4*>(S.red + (lane >> 2) * RS + (lane & 3) * 8);
  const float4 a = r[0], b = r[1];
  float q = tree4(a) + tree4(b);  // 8 lanes
  q = q + __shfl_xor_sync(SFE_FULL, q, 1);
  return q + __shfl_xor_sync(SFE_FULL, q, 2);
}
// Two warp sums in one packed butterfly, same tree: even lanes end with the total of `a`, odd lanes with that of `b`.
__device__ __forceinline__ float packed_reduce2(float a, float b, int lane) {
  const bool odd = lane & 1;
  float r = (odd ? b : a) + __shfl_xor_sync(SFE_FULL, odd ? a : b, 1);
  r = r + __shfl_xor_sync(SFE_FULL, r, 2);
  r = r + __shfl_xor_sync(SFE_FULL, r, 4);
  r = r + __shfl_xor_sync(SFE_FULL, r, 8);
  return r + __shfl_xor_sync(SFE_FULL, r, 16);
}

// Finite differences of the six scores (hessian.h:163-169), lane-parallel.  Stage A: lane j < 8
// forms q_j = [0.5 *] (score[P_j] - score[M_j]) / h; stage B: lane j < 4 forms (q[PB_j] - q[MB_j]) / h.
// All arithmetic is IEEE double, operation for operation what the reference evaluates.  `sc` holds
// score s in lanes 4s..4s+3 (the layout packed_reduce8 leaves).  d[6] = dx,dy,dxx,dxy,dyx,dyy.
__device__ __forceinline__ void finite_differences(float sc, int lane, float (&d)[6]) {
  constexpr unsigned PA = 0x55040343u, MA = 0x34201021u;  // minuend / subtrahend score of quotient j (nibbles)
  constexpr unsigned PB = 0x7642u, MB = 0x4253u;          // dxx,dyy,dxy,dyx: minuend / subtrahend quotient
  const int j = lane & 7;
  const double p = (double)__shfl_sync(SFE_FULL, sc, 4 * ((PA >> (4 * j)) & 7));
  const double m = (double)__shfl_sync(SFE_FULL, sc, 4 * ((MA >> (4 * j)) & 7));
  double num = __dsub_rn(p, m);
  if (j < 2) num = __dmul_rn(0.5, num);
  const double q = div_h(num);
  const int jb = lane & 3;
  const double qp = __shfl_sync(SFE_FULL, q, (PB >> (4 * jb)) & 7), qm = __shfl_sync(SFE_FULL, q, (MB >> (4 * jb)) & 7);
  const double r = div_h(__dsub_rn(qp, qm));
  const float qf = (float)q, rf = (float)r;
  d[0] = __shfl_sync(SFE_FULL, qf, 0);
  d[1] = __shfl_sync(SFE_FULL, qf, 1);
  d[2] = __shfl_sync(SFE_FULL, rf, 0);
  d[5] = __shfl_sync(SFE_FULL, rf, 1);
  d[3] = __shfl_sync(SFE_FULL, rf, 2);
  d[4] = __shfl_sync(SFE_FULL, rf, 3);
}

// Statistics of the template patch (hessian.h:32-40).  The patch itself and its effective mask (the mask
// weight, 0 where the template pixel is exactly 0, hessian.h:134) are parked in WarpScratch::TM: they
// are needed only in the score loop, and keeping them out of the registers while the 36 candidate values
// are live is worth 3 % (profiles/README.md).
struct Tmpl {
  float mean, sumsq;
};

// Which footprint the shared tile currently holds (warp-uniform).
struct TileTag {
  const float* img;
  unsigned key;
};

// One patch evaluation at (x,y) of image `im`:
//   is_tmpl   GetPatch (hessian.h:54-93): t <- the patch, its statistics and its effective mask
//   otherwise BruteHessian (hessian.h:147-172) against the template t: the six derivatives
//             d[6] = dx,dy,dxx,dxy,dyx,dyy (rounded to float as the reference stores them through
//             float*); returns sad0.
__device__ __forceinline__ float evaluate(WarpScratch& S, TileTag& tag, const ImgView im, bool is_tmpl, Tmpl& t,
                                          const float* __restrict__ mask, float x, float y, int lane, const PixPack& pix,
                                          float (&d)[6]) {
  __syncwarp();
  const AxisGeom g = lane_geom(x, y, lane);
  const int ox = (int)floorf(x) - 7, oy = (int)floorf(y) - 7;
  // a Newton step moves the point by at most one pixel, so consecutive steps usually share their footprint:
  // restage only when the image or the footprint origin changed
  const unsigned key = (unsigned)(ox + 0x4000) | ((unsigned)(oy + 0x4000) << 15);  // 15 bits each, bit 31 = nonpos
  if (tag.img != im.p || (tag.key & 0x7fffffffu) != key) {
    const bool np = stage_tile(S, im, ox, oy, lane);
    tag.img = im.p;
    tag.key = key | (np ? 0x80000000u : 0u);
  }
  const bool nonpos = tag.key >> 31;
  // the eight geometry lanes publish their axis variant; every lane reads all six back (broadcast loads) and derives
  // the route decisions itself -- no shuffles, no votes
  if (lane < 8) S.geo[lane] = make_float4(g.a, g.a1, __int_as_float(g.i0), __int_as_float(g.r));
  __syncwarp();
  GeoAll G;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    G.x[j] = S.geo[j];
    G.y[j] = S.geo[4 + j];
  }
  const int ix = gi(G.x[0].z), iy = gi(G.y[0].z);
  // shifts do not share their taps when an origin or a clip count differs from the unshifted variant's
  const unsigned differ = (unsigned)((gi(G.x[1].z) ^ ix) | (gi(G.x[2].z) ^ ix) | (gi(G.y[1].z) ^ iy) | (gi(G.y[2].z) ^ iy) |
                                     (gi(G.x[1].w) ^ gi(G.x[0].w)) | (gi(G.x[2].w) ^ gi(G.x[0].w)) |
                                     (gi(G.y[1].w) ^ gi(G.y[0].w)) | (gi(G.y[2].w) ^ gi(G.y[0].w)));
  const unsigned clipped0 = (unsigned)(gi(G.x[0].w) | gi(G.y[0].w));  // the unshifted variant (template patches)
  const unsigned clipped = clipped0 | (unsigned)(gi(G.x[1].w) | gi(G.x[2].w) | gi(G.y[1].w) | gi(G.y[2].w));
  // 0 <= ix <= w - 14 as one unsigned compare per axis (a level narrower than the patch makes the bound 0: never true)
  const bool interior = (unsigned)ix < (unsigned)max(im.w - SFE_PATCH, 0) && (unsigned)iy < (unsigned)max(im.h - SFE_PATCH, 0);

  // fast: no border rule, no clipping, all six shifts share their taps, and no pixel can be exactly 0 (a strictly
  // positive footprint under positive weights that sum to 1) -- patch values stay in registers.
  // plain: no border rule / clipping for any shift (integer origins within +-1 of the unshifted one).
  const bool zeros = nonpos || clipped != 0;
  const bool fast = !is_tmpl && differ == 0 && !zeros && interior;
  const bool plain = is_tmpl ? (interior && clipped0 == 0)
                             : (clipped == 0 && (unsigned)(ix - 1) < (unsigned)max(im.w - SFE_PATCH - 2, 0) &&
                                (unsigned)(iy - 1) < (unsigned)max(im.h - SFE_PATCH - 2, 0));  // 1 <= ix <= w - 15
  if (!fast) {
    int nshift = is_tmpl ? 1 : 6;
    asm volatile("" : "+r"(nshift));  // opaque: one copy of each loop serves both callers (code size)
    if (plain) straddle_sample(S, ox, oy, nshift, lane, pix);
    else if (differ == 0) general_sample_shared(S, im, G, ox, oy, nshift, lane, pix);
    else general_sample(S, im, ox, oy, nshift, false, lane, pix);
  }

  if (is_tmpl) {
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < SFE_SLOTS; ++k) {  // hessian.h:85-91
      const float v = S.v2[k * 32 + lane].x;
      sm = sm + v;
      sq = fmaf(v, v, sq);
      const float m = (mask && lane + 32 * k < SFE_PLEN) ? __ldg(mask + lane + 32 * k) : 0.f;
      S.TM[k * 32 + lane] = make_float2(v, v == 0.f ? 0.f : m);
    }
    const float red = div169(packed_reduce2(sm, sq, lane));
    t.mean = __shfl_sync(SFE_FULL, red, 0);
    t.sumsq = __shfl_sync(SFE_FULL, red, 1);
    return 0.f;
  }

  // ---- sampling + patch statistics of the six candidates (hessian.h:85-91): 12 partial sums per lane
  float st[16];
  float2 v[3][SFE_SLOTS];  // candidate patch values, shifts paired: v[p][k] = (shift 2p, shift 2p+1) of slot k
  if (fast) {
    // every patch pixel reads its 4 taps once; six weight sets
    float ax[3], ax1[3], ay[3], ay1[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      ax[j] = G.x[j].x;
      ax1[j] = G.x[j].y;
      ay[j] = G.y[j].x;
      ay1[j] = G.y[j].y;
    }
    float t00[SFE_SLOTS], t01[SFE_SLOTS], t10[SFE_SLOTS], t11[SFE_SLOTS];
    const int base = (iy - oy) * TS + (ix - ox);
#pragma unroll
    for (int k = 0; k < SFE_SLOTS; ++k) {
      const float* tp = S.tile + ((k < SFE_SLOTS - 1 || lane + 32 * k < SFE_PLEN) ? base + pix_off(pix, k) : ZOFF);
      t00[k] = tp[0];
      t01[k] = tp[1];
      t10[k] = tp[TS];
      t11[k] = tp[TS + 1];
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int jxa = (SXP >> (4 * p)) & 3, jya = (SYP >> (4 * p)) & 3, jxb = (SXP >> (4 * p + 2)) & 3, jyb = (SYP >> (4 * p + 2)) & 3;
      const float2 w0 = make_float2(ax1[jxa] * ay1[jya], ax1[jxb] * ay1[jyb]), w1 = make_float2(ax[jxa] * ay1[jya], ax[jxb] * ay1[jyb]);
      const float2 w2 = make_float2(ax1[jxa] * ay[jya], ax1[jxb] * ay[jyb]), w3 = make_float2(ax[jxa] * ay[jya], ax[jxb] * ay[jyb]);
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k)
        v[p][k] = fma2(both(t11[k]), w3, fma2(both(t10[k]), w2, fma2(both(t01[k]), w1, mul2(both(t00[k]), w0))));
    }
  } else {
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k)
        v[p][k] = S.v2[(p * SFE_SLOTS + k) * 32 + lane];
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    float2 sm = both(0.f), sq = both(0.f);
#pragma unroll
    for (int k = 0; k < SFE_SLOTS; ++k) {
      sm = add2(sm, v[p][k]);
      sq = fma2(v[p][k], v[p][k], sq);
    }
    st[2 * p] = sm.x;
    st[2 * p + 1] = sm.y;
    st[8 + 2 * p] = sq.x;
    st[8 + 2 * p + 1] = sq.y;
  }
  const float red = reduce_stats(S, st, lane);  // lanes 2s,2s+1: sum_s; lanes 16+2s,17+2s: sumsq_s
  const float other = __shfl_xor_sync(SFE_FULL, red, 16);
  // lane-parallel alpha/beta (hessian.h:131-132): lanes 2s (s < 6) hold the values of shift s; lanes >= 12
  // hold padding and get benign operands so the IEEE divide/sqrt stay on their fast paths
  const bool live = lane < 12;
  const float mean = div169(red), sumsq = live ? div169(other) : 1.f;
  const float alpha_l = sqrtf((live ? t.sumsq : 1.f) / sumsq);
  const float beta_l = t.mean - alpha_l * mean;
  float part[8];
  part[6] = part[7] = 0.f;
  float T[SFE_SLOTS], mkT[SFE_SLOTS];  // parked in shared memory while the patch values occupy the registers
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    const float2 tm = S.TM[k * 32 + lane];
    T[k] = tm.x;
    mkT[k] = tm.y;
  }
  // lanes 0, 2, .., 10 publish -alpha and -beta of shifts 0..5 ((-v)*alpha == v*(-alpha) and x - beta == x + (-beta), exactly)
  if (live && !(lane & 1)) {
    S.ab[lane >> 1] = -alpha_l;
    S.ab[8 + (lane >> 1)] = -beta_l;
  }
  __syncwarp();
  if (!zeros) {
    const float4 na03 = *reinterpret_cast<const float4*>(S.ab), nb03 = *reinterpret_cast<const float4*>(S.ab + 8);
    const float2 na45 = *reinterpret_cast<const float2*>(S.ab + 4), nb45 = *reinterpret_cast<const float2*>(S.ab + 12);
    const float2 nas[3] = {make_float2(na03.x, na03.y), make_float2(na03.z, na03.w), na45};
    const float2 nbs[3] = {make_float2(nb03.x, nb03.y), make_float2(nb03.z, nb03.w), nb45};
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const float2 na = nas[p], nb = nbs[p];
      float2 acc = both(0.f);
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) {  // hessian.h:133-139; template zeros are folded into mkT
        float2 diff = add2(fma2(v[p][k], na, both(T[k])), nb);
        diff = mul2(diff, diff);
        acc = fma2(diff, both(mkT[k]), acc);
      }
      part[2 * p] = acc.x;
      part[2 * p + 1] = acc.y;
    }
  } else {
    // candidate pixels may be exactly 0 (hessian.h:134 skips them): rare, so a compact rolled loop over S.v2
#pragma unroll
    for (int j = 0; j < 6; ++j) part[j] = 0.f;
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
      const float alpha = -S.ab[s], beta = -S.ab[8 + s];
      float p = 0.f;
#pragma unroll
      for (int k = 0; k < SFE_SLOTS; ++k) {
        const float vv = reinterpret_cast<const float*>(S.v2)[((s >> 1) * SFE_SLOTS + k) * 64 + 2 * lane + (s & 1)];
        float diff = fmaf(-vv, alpha, T[k]) - beta;
        diff = diff * diff;
        const float q = fmaf(diff, mkT[k], p);
        p = vv == 0.f ? p : q;
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) part[j] = s == j ? p : part[j];
    }
  }
  const float sc = reduce_scores(S, part, lane);  // score s in lanes 4s..4s+3
  finite_differences(sc, lane, d);
  return __shfl_sync(SFE_FULL, sc, 0);
}

__device__ __forceinline__ void init_scratch(WarpScratch& S, int lane) {
  for (int i = lane; i < TS + 2; i += 32) S.zero[i] = 0.f;
  for (int i = lane; i < 16 * RS; i += 32) S.red[i] = 0.f;
  __syncwarp();
}

// GetPatches (hessian.h:175-183) on the template pyramid + TrackFeature (hessian.h:243-264) with Track
// (hessian.h:185-241) on the search pyramid.  (x,y) is updated only on success.
__device__ __forceinline__ int track_feature(WarpScratch& S, TileTag& tag, const PyrView& tp, int tframe, float tx, float ty,
                                             const PyrView& sp, int sframe, int levels, float thr, int maxit,
                                             const float* __restrict__ mask, float& x, float& y, int lane, const PixPack& pix,
                                             int& steps) {
  const int lv = min(min(tp.depth, sp.depth), levels);
  const float margin = 0.01f;
  float px = x * (float)(1. / (1 << (lv - 1))), py = y * (float)(1. / (1 << (lv - 1)));
#pragma unroll 1
  for (int i = lv - 1; i >= 0; --i) {
    const float sc = (float)(1. / (1 << i));  // pt *= 0.5 i times (exact)
    Tmpl t;
    t.mean = t.sumsq = 0.f;
    const ImgView tim = img_of(tp, 0, i, tframe), sim = img_of(sp, 0, i, sframe);
    const float wf = (float)sim.w, hf = (float)sim.h;
    // it == -1 extracts the template patch of this level (GetPatches); it >= 0 are the Newton steps
#pragma unroll 1
    for (int it = -1; it < maxit; ++it) {
      const bool is_tmpl = it < 0;
      if (!is_tmpl && (px < margin || py < margin || (px + margin) > wf || (py + margin) > hf))
        return SFE_OUT_OF_BOUNDS;
      ImgView im;
      im.p = is_tmpl ? tim.p : sim.p;
      im.w = sim.w; im.h = sim.h; im.pitch = sim.pitch;  // both pyramids have the same geometry
      float d[6];
      evaluate(S, tag, im, is_tmpl, t, mask, is_tmpl ? tx * sc : px, is_tmpl ? ty * sc : py, lane, pix, d);
      if (is_tmpl) continue;
      ++steps;
      float dx, dy;
      newton_step(d[0], d[1], d[2], d[3], d[4], d[5], dx, dy);
      px += clamp1(dx);
      py += clamp1(dy);
      if (fabsf(dx) < thr && fabsf(dy) < thr) break;
    }
    if (i > 0) { px *= 2.f; py *= 2.f; }
  }
  x = px;
  y = py;
  return SFE_OK;
}

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
  const PixPack pix = make_pixpack(lane);
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
    TileTag tag{nullptr, 0u};
    float bx = fx, by = fy;  // matcher.cpp:181
    // dir 0: template from `from` at from_pt, search `to` from the seed (matcher.cpp:175-176)
    // dir 1: template from `to` at the forward result, search `from` from from_pt (matcher.cpp:180-182)
    st[1] = SFE_OK;
#pragma unroll 1
    for (int dir = 0; dir < a.ndir; ++dir) {
      const PyrView& tp = dir == 0 ? from : to;
      const PyrView& sp = dir == 0 ? to : from;
      float x = dir == 0 ? tx : bx, y = dir == 0 ? ty : by;
      const int s = track_feature(S, tag, tp, dir == 0 ? ff : tf, dir == 0 ? fx : tx, dir == 0 ? fy : ty, sp,
                                  dir == 0 ? tf : ff, lv, a.thr, a.maxit, mask, x, y, lane, pix, steps);
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

// GetPatch (hessian.h:54-93) for n points of one level
__global__ void __launch_bounds__(32 * TRK_WARPS) get_patches_kernel(PyrView v, int frame, int level, int n,
                                                                     const float* __restrict__ xy, float* __restrict__ patches,
                                                                     float* __restrict__ mean, float* __restrict__ sumsq) {
  __shared__ WarpScratch scratch[TRK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, i = blockIdx.x * TRK_WARPS + warp;
  if (i >= n) return;
  init_scratch(scratch[warp], lane);
  float d[6];
  Tmpl t;
  TileTag tag{nullptr, 0u};
  evaluate(scratch[warp], tag, img_of(v, 0, level, frame), true, t, nullptr, xy[2 * i], xy[2 * i + 1], lane, make_pixpack(lane), d);
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k)
    if (lane + 32 * k < SFE_PLEN) patches[(size_t)i * SFE_PLEN + lane + 32 * k] = scratch[warp].TM[k * 32 + lane].x;
  if (lane == 0) { mean[i] = t.mean; sumsq[i] = t.sumsq; }
}

// BruteHessian (hessian.h:147-172) for n (template point, search point) pairs of one level
__global__ void __launch_bounds__(32 * TRK_WARPS) brute_hessian_kernel(PyrView tv, int tframe, PyrView sv, int sframe,
                                                                       int level, int n, const float* __restrict__ txy,
                                                                       const float* __restrict__ xy, float* __restrict__ out7,
                                                                       const float* __restrict__ mask) {
  __shared__ WarpScratch scratch[TRK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * TRK_WARPS + warp;
  if (i >= n) return;
  init_scratch(scratch[warp], lane);
  Tmpl t;
  float d[6];
  TileTag tag{nullptr, 0u};
  evaluate(scratch[warp], tag, img_of(tv, 0, level, tframe), true, t, mask, txy[2 * i], txy[2 * i + 1], lane, make_pixpack(lane), d);
  const float s0 = evaluate(scratch[warp], tag, img_of(sv, 0, level, sframe), false, t, mask, xy[2 * i], xy[2 * i + 1], lane, make_pixpack(lane), d);
  if (lane == 0) {
    out7[7 * i] = s0;
    for (int k = 0; k < 6; ++k) out7[7 * i + 1 + k] = d[k];
  }
}

}  // namespace

int launch_track_hessian(const PyrView& from, const PyrView& to, const TrackArgs& a, const float* mask, int* counter,
                         int num_sms, cudaStream_t s) {
  if (a.n <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return -(int)e;
  const int blocks = min((a.n + TRK_WARPS - 1) / TRK_WARPS, num_sms * TRK_MINB);
  track_fb_kernel<<<blocks, 32 * TRK_WARPS, 0, s>>>(from, to, a, mask, counter);
  e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

int launch_get_patches(const PyrView& v, int frame, int level, int n, const float* xy, float* patches, float* mean,
                       float* sumsq, cudaStream_t s) {
  if (n <= 0) return 0;
  get_patches_kernel<<<(n + TRK_WARPS - 1) / TRK_WARPS, 32 * TRK_WARPS, 0, s>>>(v, frame, level, n, xy, patches, mean, sumsq);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

int launch_brute_hessian(const PyrView& tv, int tframe, const PyrView& sv, int sframe, int level, int n, const float* txy,
                         const float* xy, float* out7, const float* mask, cudaStream_t s) {
  if (n <= 0) return 0;
  brute_hessian_kernel<<<(n + TRK_WARPS - 1) / TRK_WARPS, 32 * TRK_WARPS, 0, s>>>(tv, tframe, sv, sframe, level, n, txy, xy,
                                                                                  out7, mask);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}
