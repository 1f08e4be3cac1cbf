// pyramid.cu -- MakePyramid of the three reference trackers on the GPU.
//   SFE_HESSIAN  hessian.h:95-126  gray -> f32/255 -> blur(1.1); level i = blur(0.8) o pyrDown
//   SFE_KLT      klt.h:98-137      gray -> f32/255, Scharr/32 grads; level i = blur(0.6) o pyrDown,
//                                   grads = 2 * pyrDown(prev grads)
//   SFE_BRUTE    brute.h:59-80     gray -> f32/255; level i = pyrDown
// The arithmetic (operation order, FMA placement, BORDER_REFLECT_101) is the one of
// oracle/oracle.c, which is pinned bit-for-bit against OpenCV 4.13 (see oracle/oracle.h).
//
// Two kernels, both HBM-streaming:
//   pyr_l0_kernel    one pass over the BGR bytes: gray conversion, /255, and the 5x5 blur (or the
//                    Scharr pair) fused, staged through shared memory; 128-bit stores.
//   pyr_down_kernel  pyrDown and the following 5x5 blur fused: the (2T+11)^2 footprint of the
//                    previous level is staged once in shared memory, both separable passes of both
//                    filters run out of shared memory, the level is written once.
#include <stdlib.h>

#include "pyr_math.cuh"

namespace {

constexpr int L0_TW = 64, L0_TH = 16, L0_THREADS = 256;
constexpr int DN_TW = 32, DN_TH = 16, DN_THREADS = 256;

__device__ __forceinline__ float gray_f32(const uint8_t* px) {
  // cvtColor(RGB2GRAY) on BGR bytes (hessian.h:100) then convertTo(CV_32F, 1/255.) (hessian.h:101)
  int g = (9798 * px[0] + 19235 * px[1] + 3735 * px[2] + (1 << 14)) >> 15;
  return (float)g * (float)(1. / 255.);
}

// ------------------------------------------------------------------------------- level 0
template <int FLAVOR>
__global__ void __launch_bounds__(L0_THREADS) pyr_l0_kernel(PyrView v, const uint8_t* __restrict__ bgr,
                                                            size_t row_stride, size_t frame_stride,
                                                            int first) {
  constexpr int R = (FLAVOR == SFE_HESSIAN) ? 2 : (FLAVOR == SFE_KLT ? 1 : 0);
  constexpr int GW = L0_TW + 2 * R, GH = L0_TH + 2 * R;
  __shared__ float G[GH][GW + 1];
  __shared__ float T[(FLAVOR == SFE_BRUTE) ? 1 : GH][L0_TW + 1];
  __shared__ float T2[(FLAVOR == SFE_KLT) ? GH : 1][L0_TW + 1];

  const int w = v.w[0], h = v.h[0], pitch = v.pitch[0];
  const int frame = first + blockIdx.z;
  const int x0 = blockIdx.x * L0_TW, y0 = blockIdx.y * L0_TH;
  const uint8_t* src = bgr + (size_t)blockIdx.z * frame_stride;
  const int tid = threadIdx.x;

  for (int e = tid; e < GH * GW; e += L0_THREADS) {
    int j = e % GW, i = e / GW;
    int x = reflect101(x0 - R + j, w), y = reflect101(y0 - R + i, h);
    G[i][j] = gray_f32(src + (size_t)y * row_stride + 3 * x);
  }
  __syncthreads();

  float* out = v.base[0][0] + (long long)frame * v.frame_stride[0];
  if (FLAVOR == SFE_BRUTE) {
    for (int e = tid; e < L0_TH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      int x = x0 + j, y = y0 + i;
      if (x < w && y < h) out[(size_t)y * pitch + x] = G[i][j];
    }
    return;
  }
  if (FLAVOR == SFE_HESSIAN) {
    const Taps t = taps_for(0);
    for (int e = tid; e < GH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      T[i][j] = blur_row(G[i][j], G[i][j + 1], G[i][j + 2], G[i][j + 3], G[i][j + 4], t);
    }
    __syncthreads();
    for (int e = tid; e < L0_TH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      int x = x0 + j, y = y0 + i;
      if (x < w && y < h)
        out[(size_t)y * pitch + x] = blur_col(T[i][j], T[i + 1][j], T[i + 2][j], T[i + 3][j], T[i + 4][j], t);
    }
  } else {  // SFE_KLT: image plane unblurred + Scharr/32 gradients (klt.h:104-106)
    const float k3 = 3.f / 32.f, k10 = 10.f / 32.f;
    float* ogx = v.base[1][0] + (long long)frame * v.frame_stride[0];
    float* ogy = v.base[2][0] + (long long)frame * v.frame_stride[0];
    for (int e = tid; e < GH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      float m = G[i][j], c = G[i][j + 1], p = G[i][j + 2];
      T[i][j] = p - m;
      T2[i][j] = fmaf(k10, c, (m + p) * k3);
    }
    __syncthreads();
    for (int e = tid; e < L0_TH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      int x = x0 + j, y = y0 + i;
      if (x < w && y < h) {
        size_t o = (size_t)y * pitch + x;
        out[o] = G[i + 1][j + 1];
        ogx[o] = fmaf(k3, T[i][j] + T[i + 2][j], T[i + 1][j] * k10);
        ogy[o] = T2[i + 2][j] - T2[i][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------- level i>0
// out = post_scale * blur(pyrDown(prev)) for one plane; blur_id < 0 skips the blur.
__global__ void __launch_bounds__(DN_THREADS) pyr_down_kernel(const float* __restrict__ prev_base,
                                                              long long prev_fs, int pw, int ph, int ppitch,
                                                              float* __restrict__ cur_base, long long cur_fs,
                                                              int cw, int ch, int cpitch, int first,
                                                              int blur_id, float post_scale) {
  constexpr int PW = 2 * DN_TW + 11, PH = 2 * DN_TH + 11;  // staged footprint of prev level
  constexpr int DW = DN_TW + 4, DH = DN_TH + 4;            // pyrDown outputs incl. blur halo
  __shared__ float P[PH][PW];
  __shared__ float Hh[PH][DW + 1];
  __shared__ float D[DH][DW + 1];
  __shared__ float B[DH][DN_TW + 1];

  const int frame = first + blockIdx.z;
  const float* prev = prev_base + (long long)frame * prev_fs;
  float* cur = cur_base + (long long)frame * cur_fs;
  const int x0 = blockIdx.x * DN_TW, y0 = blockIdx.y * DN_TH;
  const int tid = threadIdx.x;
  const int halo = (blur_id >= 0) ? 2 : 0;
  // pyrDown outputs needed: columns [lox, hix), rows [loy, hiy) (blur halo clipped to the image:
  // reflected coordinates fall back inside this range)
  const int lox = max(x0 - halo, 0), hix = min(x0 + DN_TW + halo, cw);
  const int loy = max(y0 - halo, 0), hiy = min(y0 + DN_TH + halo, ch);
  const int nx = hix - lox, ny = hiy - loy;
  const int px0 = 2 * lox - 2, py0 = 2 * loy - 2;
  const int pnx = 2 * nx + 3, pny = 2 * ny + 3;

  for (int e = tid; e < pny * pnx; e += DN_THREADS) {
    int j = e % pnx, i = e / pnx;
    int x = reflect101(px0 + j, pw), y = reflect101(py0 + i, ph);
    P[i][j] = prev[(size_t)y * ppitch + x];
  }
  __syncthreads();
  // horizontal pyrDown pass: Hh[i][j] for prev row py0+i, output column lox+j
  for (int e = tid; e < pny * nx; e += DN_THREADS) {
    int j = e % nx, i = e / nx;
    const float* r = &P[i][2 * j];
    Hh[i][j] = ((r[0] + r[4]) + (r[1] + r[3]) * 4.f) + r[2] * 6.f;
  }
  __syncthreads();
  // vertical pass
  for (int e = tid; e < ny * nx; e += DN_THREADS) {
    int j = e % nx, i = e / nx;
    float r0 = Hh[2 * i][j], r1 = Hh[2 * i + 1][j], r2 = Hh[2 * i + 2][j], r3 = Hh[2 * i + 3][j], r4 = Hh[2 * i + 4][j];
    D[i][j] = (((r1 + r3) + r2) * 4.f + ((r0 + r4) + (r2 + r2))) * (1.f / 256.f);
  }
  __syncthreads();
  if (blur_id < 0) {
    for (int e = tid; e < DN_TH * DN_TW; e += DN_THREADS) {
      int j = e % DN_TW, i = e / DN_TW;
      int x = x0 + j, y = y0 + i;
      if (x < cw && y < ch) cur[(size_t)y * cpitch + x] = D[y - loy][x - lox] * post_scale;
    }
    return;
  }
  const Taps t = taps_for(blur_id);
  // blur row pass for rows [loy,hiy), columns [x0, x0+TW)
  for (int e = tid; e < ny * DN_TW; e += DN_THREADS) {
    int j = e % DN_TW, i = e / DN_TW;
    int x = x0 + j;
    if (x < cw) {
      const float* d = D[i];
      B[i][j] = blur_row(d[reflect101(x - 2, cw) - lox], d[reflect101(x - 1, cw) - lox], d[x - lox],
                         d[reflect101(x + 1, cw) - lox], d[reflect101(x + 2, cw) - lox], t);
    }
  }
  __syncthreads();
  for (int e = tid; e < DN_TH * DN_TW; e += DN_THREADS) {
    int j = e % DN_TW, i = e / DN_TW;
    int x = x0 + j, y = y0 + i;
    if (x < cw && y < ch) {
      float r = blur_col(B[reflect101(y - 2, ch) - loy][j], B[reflect101(y - 1, ch) - loy][j], B[y - loy][j],
                         B[reflect101(y + 1, ch) - loy][j], B[reflect101(y + 2, ch) - loy][j], t);
      cur[(size_t)y * cpitch + x] = r * post_scale;
    }
  }
}


// =============================================================================== fast paths
// SFE_HESSIAN flavour (the live tracker), frame width a multiple of 4 and 4-byte aligned BGR rows.
// Same arithmetic as the generic kernels above, different data movement: fixed 2-D thread mappings
// (no runtime div/mod), 32-bit global loads of the BGR bytes (4 pixels = 3 words), 128-bit shared
// and global accesses, four outputs per thread in the row passes and a 4x4 block per thread in the
// column passes.

constexpr int F_TW = 128, F_TH = 32, F_THREADS = 256;
constexpr int F_GW = F_TW + 8;  // gray tile: columns x0-4 .. x0+TW+3 (halo 2, padded to 4 for alignment)
constexpr int F_GH = F_TH + 4;  // rows y0-2 .. y0+TH+1

__device__ __forceinline__ float gray_from(int c0, int c1, int c2) {
  int g = (9798 * c0 + 19235 * c1 + 3735 * c2 + (1 << 14)) >> 15;
  return (float)g * (float)(1. / 255.);
}

__global__ void __launch_bounds__(F_THREADS) pyr_l0_hessian_fast(PyrView v, const uint8_t* __restrict__ bgr,
                                                                 size_t row_stride, size_t frame_stride, int first) {
  __shared__ __align__(16) float G[F_GH][F_GW];
  __shared__ __align__(16) float R[F_GH][F_TW];
  const int w = v.w[0], h = v.h[0], pitch = v.pitch[0];
  const int frame = first + blockIdx.z;
  const int x0 = blockIdx.x * F_TW, y0 = blockIdx.y * F_TH;
  const uint8_t* src = bgr + (size_t)blockIdx.z * frame_stride;
  const int tid = threadIdx.x;

  // ---- phase 1: BGR bytes -> gray/255 (4 pixels = 12 bytes = 3 aligned words per item)
  constexpr int GROUPS = F_GW / 4;  // 34
  for (int it = tid; it < F_GH * GROUPS; it += F_THREADS) {
    const int i = it / GROUPS, j4 = it - i * GROUPS;
    const int y = reflect101(y0 - 2 + i, h);
    const int xs = x0 - 4 + 4 * j4;
    const uint8_t* rowp = src + (size_t)y * row_stride;
    float4 o;
    if (xs >= 0 && xs + 3 < w) {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(rowp + 3 * xs);
      const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
      o.x = gray_from(a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xff);
      o.y = gray_from(a >> 24, b & 0xff, (b >> 8) & 0xff);
      o.z = gray_from((b >> 16) & 0xff, b >> 24, c & 0xff);
      o.w = gray_from((c >> 8) & 0xff, (c >> 16) & 0xff, c >> 24);
    } else {
      float t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint8_t* px = rowp + 3 * reflect101(xs + k, w);
        t[k] = gray_from(px[0], px[1], px[2]);
      }
      o = make_float4(t[0], t[1], t[2], t[3]);
    }
    *reinterpret_cast<float4*>(&G[i][4 * j4]) = o;
  }
  __syncthreads();

  // ---- phase 2: horizontal blur, 4 outputs per item (output column c reads G columns c+2 .. c+6)
  const Taps t = taps_for(0);
  for (int it = tid; it < F_GH * (F_TW / 4); it += F_THREADS) {
    const int i = it >> 5, j4 = it & 31;
    const float2 a = *reinterpret_cast<const float2*>(&G[i][4 * j4 + 2]);
    const float4 b = *reinterpret_cast<const float4*>(&G[i][4 * j4 + 4]);
    const float2 c = *reinterpret_cast<const float2*>(&G[i][4 * j4 + 8]);
    float4 o;
    o.x = blur_row(a.x, a.y, b.x, b.y, b.z, t);
    o.y = blur_row(a.y, b.x, b.y, b.z, b.w, t);
    o.z = blur_row(b.x, b.y, b.z, b.w, c.x, t);
    o.w = blur_row(b.y, b.z, b.w, c.x, c.y, t);
    *reinterpret_cast<float4*>(&R[i][4 * j4]) = o;
  }
  __syncthreads();

  // ---- phase 3: vertical blur, a 4x4 output block per thread, 128-bit coalesced stores
  {
    const int cg = tid & 31, seg = tid >> 5;
    float4 r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = *reinterpret_cast<const float4*>(&R[seg * 4 + k][4 * cg]);
    float* out = v.base[0][0] + (long long)frame * v.frame_stride[0];
    const int x = x0 + 4 * cg;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = y0 + seg * 4 + k;
      if (y < h && x < w) {
        float4 o;
        o.x = blur_col(r[k].x, r[k + 1].x, r[k + 2].x, r[k + 3].x, r[k + 4].x, t);
        o.y = blur_col(r[k].y, r[k + 1].y, r[k + 2].y, r[k + 3].y, r[k + 4].y, t);
        o.z = blur_col(r[k].z, r[k + 1].z, r[k + 2].z, r[k + 3].z, r[k + 4].z, t);
        o.w = blur_col(r[k].w, r[k + 1].w, r[k + 2].w, r[k + 3].w, r[k + 4].w, t);
        *reinterpret_cast<float4*>(out + (size_t)y * pitch + x) = o;  // w % 4 == 0: the group is all in or all out
      }
    }
  }
}

// pyrDown + 5x5 blur, TW x TH outputs per CTA (64x32 for the big levels, 32x16 for the small ones).
constexpr int D_THREADS = 256;
template <int TW, int TH>
struct DownTile {
  static constexpr int PDW = TW + 8;      // pyrDown tile: columns x0-4 .. x0+TW+3 (blur halo 2, padded to 4)
  static constexpr int PDH = TH + 4;      // rows y0-2 .. y0+TH+1
  static constexpr int PW = 2 * PDW + 8;  // previous-level tile: columns 2*x0-12 .. (152 for TW = 64)
  static constexpr int PH = 2 * PDH + 3;  // rows 2*y0-6 .. (75 for TH = 32)
  static constexpr size_t SMEM = sizeof(float) * ((size_t)PH * PW + (size_t)PH * PDW);
};

template <int D_TW, int D_TH>
__global__ void __launch_bounds__(D_THREADS) pyr_down_blur_fast(const float* __restrict__ prev_base, long long prev_fs,
                                                                int pw, int ph, int ppitch, float* __restrict__ cur_base,
                                                                long long cur_fs, int cw, int ch, int cpitch, int first,
                                                                int blur_id) {
  constexpr int D_PDW = DownTile<D_TW, D_TH>::PDW, D_PDH = DownTile<D_TW, D_TH>::PDH;
  constexpr int D_PW = DownTile<D_TW, D_TH>::PW, D_PH = DownTile<D_TW, D_TH>::PH;
  extern __shared__ __align__(16) float smem[];
  float (*P)[D_PW] = reinterpret_cast<float (*)[D_PW]>(smem);                  // [D_PH][D_PW]
  float (*Hh)[D_PDW] = reinterpret_cast<float (*)[D_PDW]>(smem + D_PH * D_PW);  // [D_PH][D_PDW]
  float (*D)[D_PDW] = reinterpret_cast<float (*)[D_PDW]>(smem);                 // [D_PDH][D_PDW], aliases P
  float (*B)[D_TW] = reinterpret_cast<float (*)[D_TW]>(smem + D_PDH * D_PDW);   // [D_PDH][D_TW], aliases P

  const int frame = first + blockIdx.z;
  const float* prev = prev_base + (long long)frame * prev_fs;
  float* cur = cur_base + (long long)frame * cur_fs;
  const int x0 = blockIdx.x * D_TW, y0 = blockIdx.y * D_TH;
  const int tid = threadIdx.x;
  const int pc0 = 2 * x0 - 12, pr0 = 2 * y0 - 6;

  // ---- stage the previous level (reflect-101 on its own borders), 128-bit loads in the interior
  for (int it = tid; it < D_PH * (D_PW / 4); it += D_THREADS) {
    const int i = it / (D_PW / 4), j4 = it - i * (D_PW / 4);
    const int y = reflect101(pr0 + i, ph);
    const int xs = pc0 + 4 * j4;
    const float* rowp = prev + (size_t)y * ppitch;
    float4 o;
    if (xs >= 0 && xs + 3 < pw) {
      o = __ldg(reinterpret_cast<const float4*>(rowp + xs));
    } else {
      o.x = rowp[reflect101(xs, pw)];
      o.y = rowp[reflect101(xs + 1, pw)];
      o.z = rowp[reflect101(xs + 2, pw)];
      o.w = rowp[reflect101(xs + 3, pw)];
    }
    *reinterpret_cast<float4*>(&P[i][4 * j4]) = o;
  }
  __syncthreads();
  // ---- horizontal pyrDown pass: pd column pj (= x0-4+pj) is centred on P column 2*pj+4
  for (int it = tid; it < D_PH * (D_PDW / 4); it += D_THREADS) {
    const int i = it / (D_PDW / 4), j4 = it - i * (D_PDW / 4);
    const float4 a = *reinterpret_cast<const float4*>(&P[i][8 * j4]);       // P cols 8j4 .. +3
    const float4 b = *reinterpret_cast<const float4*>(&P[i][8 * j4 + 4]);   // +4 .. +7
    const float4 c = *reinterpret_cast<const float4*>(&P[i][8 * j4 + 8]);   // +8 .. +11
    const float2 d = *reinterpret_cast<const float2*>(&P[i][8 * j4 + 12]);  // +12, +13
    float4 o;  // outputs pj = 4*j4 + k: taps 2pj+2 .. 2pj+6 = 8j4 + 2k + 2 .. + 6
    o.x = pd_h(a.z, a.w, b.x, b.y, b.z);
    o.y = pd_h(b.x, b.y, b.z, b.w, c.x);
    o.z = pd_h(b.z, b.w, c.x, c.y, c.z);
    o.w = pd_h(c.x, c.y, c.z, c.w, d.x);
    *reinterpret_cast<float4*>(&Hh[i][4 * j4]) = o;
  }
  __syncthreads();
  // ---- vertical pyrDown pass: pd row pi (= y0-2+pi) is centred on Hh row 2*pi+2 (taps 2pi .. 2pi+4)
  for (int it = tid; it < D_PDH * (D_PDW / 4); it += D_THREADS) {
    const int i = it / (D_PDW / 4), j4 = it - i * (D_PDW / 4);
    const float4 r0 = *reinterpret_cast<const float4*>(&Hh[2 * i][4 * j4]);
    const float4 r1 = *reinterpret_cast<const float4*>(&Hh[2 * i + 1][4 * j4]);
    const float4 r2 = *reinterpret_cast<const float4*>(&Hh[2 * i + 2][4 * j4]);
    const float4 r3 = *reinterpret_cast<const float4*>(&Hh[2 * i + 3][4 * j4]);
    const float4 r4 = *reinterpret_cast<const float4*>(&Hh[2 * i + 4][4 * j4]);
    float4 o;
    o.x = pd_v(r0.x, r1.x, r2.x, r3.x, r4.x);
    o.y = pd_v(r0.y, r1.y, r2.y, r3.y, r4.y);
    o.z = pd_v(r0.z, r1.z, r2.z, r3.z, r4.z);
    o.w = pd_v(r0.w, r1.w, r2.w, r3.w, r4.w);
    *reinterpret_cast<float4*>(&D[i][4 * j4]) = o;  // D aliases P: every P read finished before the barrier above
  }
  __syncthreads();
  // ---- blur: rows.  D column index of image column c is c - x0 + 4; border tiles remap their taps by
  // reflect-101 on the level's own coordinates (the reflected column is inside the tile).
  const Taps t = taps_for(blur_id);
  const bool xedge = x0 < 2 || x0 + D_TW + 2 > cw, yedge = y0 < 2 || y0 + D_TH + 2 > ch;
  for (int it = tid; it < D_PDH * (D_TW / 4); it += D_THREADS) {
    const int i = it / (D_TW / 4), j4 = it - i * (D_TW / 4);
    float4 o;
    if (!xedge) {
      const float2 a = *reinterpret_cast<const float2*>(&D[i][4 * j4 + 2]);
      const float4 b = *reinterpret_cast<const float4*>(&D[i][4 * j4 + 4]);
      const float2 c = *reinterpret_cast<const float2*>(&D[i][4 * j4 + 8]);
      o.x = blur_row(a.x, a.y, b.x, b.y, b.z, t);
      o.y = blur_row(a.y, b.x, b.y, b.z, b.w, t);
      o.z = blur_row(b.x, b.y, b.z, b.w, c.x, t);
      o.w = blur_row(b.y, b.z, b.w, c.x, c.y, t);
    } else {
      float q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int x = x0 + 4 * j4 + k;
        q[k] = 0.f;
        if (x < cw) {
          const float* d = D[i] + (4 - x0);
          q[k] = blur_row(d[reflect101(x - 2, cw)], d[reflect101(x - 1, cw)], d[x], d[reflect101(x + 1, cw)], d[reflect101(x + 2, cw)], t);
        }
      }
      o = make_float4(q[0], q[1], q[2], q[3]);
    }
    *reinterpret_cast<float4*>(&B[i][4 * j4]) = o;
  }
  __syncthreads();
  // ---- blur: columns.  B row index of image row y is y - y0 + 2.
  for (int it = tid; it < D_TH * (D_TW / 4); it += D_THREADS) {
    const int r = it / (D_TW / 4), cg = it - r * (D_TW / 4);
    const int x = x0 + 4 * cg, y = y0 + r;
    if (y < ch && x < cw) {
      int i0, i1, i2, i3, i4;
      if (!yedge) { i0 = r; i1 = r + 1; i2 = r + 2; i3 = r + 3; i4 = r + 4; }
      else {
        i0 = reflect101(y - 2, ch) - y0 + 2; i1 = reflect101(y - 1, ch) - y0 + 2; i2 = r + 2;
        i3 = reflect101(y + 1, ch) - y0 + 2; i4 = reflect101(y + 2, ch) - y0 + 2;
      }
      const float4 r0 = *reinterpret_cast<const float4*>(&B[i0][4 * cg]), r1 = *reinterpret_cast<const float4*>(&B[i1][4 * cg]);
      const float4 r2 = *reinterpret_cast<const float4*>(&B[i2][4 * cg]), r3 = *reinterpret_cast<const float4*>(&B[i3][4 * cg]);
      const float4 r4 = *reinterpret_cast<const float4*>(&B[i4][4 * cg]);
      float4 o;
      o.x = blur_col(r0.x, r1.x, r2.x, r3.x, r4.x, t);
      o.y = blur_col(r0.y, r1.y, r2.y, r3.y, r4.y, t);
      o.z = blur_col(r0.z, r1.z, r2.z, r3.z, r4.z, t);
      o.w = blur_col(r0.w, r1.w, r2.w, r3.w, r4.w, t);
      *reinterpret_cast<float4*>(cur + (size_t)y * cpitch + x) = o;  // cw % 4 == 0
    }
  }
}

}  // namespace

namespace {
template <int TW, int TH>
void launch_down_fast(const PyrView& v, int l, int first, int count, int blur_id, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(pyr_down_blur_fast<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DownTile<TW, TH>::SMEM);
    attr_set = true;
  }
  dim3 gf((v.w[l] + TW - 1) / TW, (v.h[l] + TH - 1) / TH, count);
  pyr_down_blur_fast<TW, TH><<<gf, D_THREADS, DownTile<TW, TH>::SMEM, s>>>(
      v.base[0][l - 1], v.frame_stride[l - 1], v.w[l - 1], v.h[l - 1], v.pitch[l - 1], v.base[0][l], v.frame_stride[l],
      v.w[l], v.h[l], v.pitch[l], first, blur_id);
}

int build_chunk(const PyrView& v, int flavor, const uint8_t* bgr, size_t row_stride, size_t frame_stride, int first,
                int count, cudaStream_t s) {
  int launches = 0;
  // streaming path (pyramid_stream.cu): levels 0+1 fused, deeper levels while their geometry qualifies
  static const bool tiled_only = getenv("SFE_PYR_TILED") != nullptr;
  int built = 0;
  if (flavor == SFE_HESSIAN && !tiled_only)
    built = launch_pyr_stream_hessian(v, bgr, row_stride, frame_stride, first, count, s, &launches);
  dim3 g0((v.w[0] + L0_TW - 1) / L0_TW, (v.h[0] + L0_TH - 1) / L0_TH, count);
  const bool fast_l0 = flavor == SFE_HESSIAN && v.w[0] % 4 == 0 && ((uintptr_t)bgr & 3) == 0 && row_stride % 4 == 0 &&
                       frame_stride % 4 == 0;
  if (built > 0) {
    --launches;  // level 0 came with the streaming kernel (the increment below counts this branch's launch)
  } else if (fast_l0) {
    dim3 gf((v.w[0] + F_TW - 1) / F_TW, (v.h[0] + F_TH - 1) / F_TH, count);
    pyr_l0_hessian_fast<<<gf, F_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  } else if (flavor == SFE_HESSIAN) pyr_l0_kernel<SFE_HESSIAN><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  else if (flavor == SFE_KLT) pyr_l0_kernel<SFE_KLT><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  else pyr_l0_kernel<SFE_BRUTE><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  ++launches;
  for (int l = built > 0 ? built : 1; l < v.depth; ++l) {
    dim3 g((v.w[l] + DN_TW - 1) / DN_TW, (v.h[l] + DN_TH - 1) / DN_TH, count);
    int blur_id = flavor == SFE_HESSIAN ? 1 : (flavor == SFE_KLT ? 2 : -1);
    // fast path: blurred plane, both levels 4-float aligned, level at least one blur reach wide/high
    const bool fast_dn = blur_id >= 0 && v.w[l] % 4 == 0 && v.w[l - 1] % 4 == 0 && v.w[l] >= 8 && v.h[l] >= 8;
    // KLT / brute flavours: the strip-streaming down stage (pyramid_stream.cu) when the geometry qualifies
    if (flavor != SFE_HESSIAN && launch_pyr_stream_down(v, 0, l, first, count, blur_id, 1.f, s)) {
    } else if (fast_dn && v.w[l] >= 256) launch_down_fast<64, 32>(v, l, first, count, blur_id, s);
    else if (fast_dn) launch_down_fast<32, 16>(v, l, first, count, blur_id, s);
    else
      pyr_down_kernel<<<g, DN_THREADS, 0, s>>>(v.base[0][l - 1], v.frame_stride[l - 1], v.w[l - 1], v.h[l - 1],
                                               v.pitch[l - 1], v.base[0][l], v.frame_stride[l], v.w[l], v.h[l],
                                               v.pitch[l], first, blur_id, 1.f);
    ++launches;
    if (flavor == SFE_KLT) {
      for (int p = 1; p <= 2; ++p) {  // klt.h:123-124
        if (launch_pyr_stream_down(v, p, l, first, count, -1, 2.f, s)) {
          ++launches;
          continue;
        }
        pyr_down_kernel<<<g, DN_THREADS, 0, s>>>(v.base[p][l - 1], v.frame_stride[l - 1], v.w[l - 1], v.h[l - 1],
                                                 v.pitch[l - 1], v.base[p][l], v.frame_stride[l], v.w[l],
                                                 v.h[l], v.pitch[l], first, -1, 2.f);
        ++launches;
      }
    }
  }
  return launches;
}
}  // namespace

// Optional chunking of the frame batch (SFE_PYR_CHUNK_MB of level-0 planes per chunk) so that a chunk's
// level-l planes are still in L2 when level l+1 reads them.  Measured on B200 it does not pay -- the
// kernels are issue-bound, not L2-bound, and every extra launch adds a tail -- so the default is one
// chunk; the knob stays for experiments (profiles/README.md).
int launch_pyr_build(const PyrView& v, int flavor, const uint8_t* bgr, size_t row_stride,
                     size_t frame_stride, int first, int count, cudaStream_t s) {
  static int chunk_mb = -1;
  if (chunk_mb < 0) {
    const char* e = getenv("SFE_PYR_CHUNK_MB");
    chunk_mb = e ? atoi(e) : 1 << 20;  // measured: fewer, larger launches win (profiles/README.md)
  }
  const size_t l0_bytes = (size_t)v.frame_stride[0] * sizeof(float) * (flavor == SFE_KLT ? 3 : 1);
  int chunk = (int)(((size_t)chunk_mb << 20) / (l0_bytes ? l0_bytes : 1));
  if (chunk < 1) chunk = 1;
  int launches = 0;
  for (int f = 0; f < count; f += chunk) {
    const int n = count - f < chunk ? count - f : chunk;
    launches += build_chunk(v, flavor, bgr + (size_t)f * frame_stride, row_stride, frame_stride, first + f, n, s);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? launches : -(int)e;
}
