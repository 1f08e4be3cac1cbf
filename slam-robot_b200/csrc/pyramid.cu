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
      T[i][j] = blur_row_tail(x0 + j, w) ? blur_sc(G[i][j], G[i][j + 1], G[i][j + 2], G[i][j + 3], G[i][j + 4], t)
                                         : blur_row(G[i][j], G[i][j + 1], G[i][j + 2], G[i][j + 3], G[i][j + 4], t);
    }
    __syncthreads();
    for (int e = tid; e < L0_TH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      int x = x0 + j, y = y0 + i;
      if (x < w && y < h)
        out[(size_t)y * pitch + x] = blur_col_tail(x, w) ? blur_sc(T[i][j], T[i + 1][j], T[i + 2][j], T[i + 3][j], T[i + 4][j], t)
                                                          : blur_col(T[i][j], T[i + 1][j], T[i + 2][j], T[i + 3][j], T[i + 4][j], t);
    }
  } else {  // SFE_KLT: image plane unblurred + Scharr/32 gradients (klt.h:104-106)
    const float k3 = 3.f / 32.f, k10 = 10.f / 32.f;
    float* ogx = v.base[1][0] + (long long)frame * v.frame_stride[0];
    float* ogy = v.base[2][0] + (long long)frame * v.frame_stride[0];
    for (int e = tid; e < GH * L0_TW; e += L0_THREADS) {
      int j = e % L0_TW, i = e / L0_TW;
      float m = G[i][j], c = G[i][j + 1], p = G[i][j + 2];
      T[i][j] = p - m;
      // the Scharr row filter's scalar tail (last column of an odd width) fuses the other product (oracle.c orc_scharr)
      T2[i][j] = blur_row_tail(x0 + j, w) ? fmaf(k3, m + p, c * k10) : fmaf(k10, c, (m + p) * k3);
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
  const int hbody = pd_hbody(pw);
  for (int e = tid; e < pny * nx; e += DN_THREADS) {
    int j = e % nx, i = e / nx;
    const float* r = &P[i][2 * j];
    Hh[i][j] = pd_h_tail(lox + j, hbody) ? pd_h_sc(r[0], r[1], r[2], r[3], r[4]) : pd_h(r[0], r[1], r[2], r[3], r[4]);
  }
  __syncthreads();
  // vertical pass
  for (int e = tid; e < ny * nx; e += DN_THREADS) {
    int j = e % nx, i = e / nx;
    float r0 = Hh[2 * i][j], r1 = Hh[2 * i + 1][j], r2 = Hh[2 * i + 2][j], r3 = Hh[2 * i + 3][j], r4 = Hh[2 * i + 4][j];
    D[i][j] = pd_v_tail(lox + j, cw) ? pd_v_sc(r0, r1, r2, r3, r4) : pd_v(r0, r1, r2, r3, r4);
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
      const float m2 = d[reflect101(x - 2, cw) - lox], m1 = d[reflect101(x - 1, cw) - lox], c = d[x - lox];
      const float p1 = d[reflect101(x + 1, cw) - lox], p2 = d[reflect101(x + 2, cw) - lox];
      B[i][j] = blur_row_tail(x, cw) ? blur_sc(m2, m1, c, p1, p2, t) : blur_row(m2, m1, c, p1, p2, t);
    }
  }
  __syncthreads();
  for (int e = tid; e < DN_TH * DN_TW; e += DN_THREADS) {
    int j = e % DN_TW, i = e / DN_TW;
    int x = x0 + j, y = y0 + i;
    if (x < cw && y < ch) {
      const float m2 = B[reflect101(y - 2, ch) - loy][j], m1 = B[reflect101(y - 1, ch) - loy][j], c = B[y - loy][j];
      const float p1 = B[reflect101(y + 1, ch) - loy][j], p2 = B[reflect101(y + 2, ch) - loy][j];
      const float r = blur_col_tail(x, cw) ? blur_sc(m2, m1, c, p1, p2, t) : blur_col(m2, m1, c, p1, p2, t);
      cur[(size_t)y * cpitch + x] = r * post_scale;
    }
  }
}


}  // namespace

namespace {
int build_chunk(const PyrView& v, int flavor, const uint8_t* bgr, size_t row_stride, size_t frame_stride, int first,
                int count, cudaStream_t s) {
  int launches = 0;
  // streaming path (pyramid_stream.cu): levels 0+1 fused, deeper levels while their geometry qualifies
  static const bool tiled_only = getenv("SFE_PYR_TILED") != nullptr;  // experiments: the tiled kernels for every level
  int built = 0;
  if (flavor == SFE_HESSIAN && !tiled_only)
    built = launch_pyr_stream_hessian(v, bgr, row_stride, frame_stride, first, count, s, &launches);
  dim3 g0((v.w[0] + L0_TW - 1) / L0_TW, (v.h[0] + L0_TH - 1) / L0_TH, count);
  if (built > 0) {
    --launches;  // level 0 came with the streaming kernel (the increment below counts this branch's launch)
  } else if (flavor == SFE_HESSIAN) pyr_l0_kernel<SFE_HESSIAN><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  else if (flavor == SFE_KLT) {
    if (!launch_pyr_stream_klt_l0(v, bgr, row_stride, frame_stride, first, count, s))
      pyr_l0_kernel<SFE_KLT><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  }
  else pyr_l0_kernel<SFE_BRUTE><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  ++launches;
  for (int l = built > 0 ? built : 1; l < v.depth; ++l) {
    dim3 g((v.w[l] + DN_TW - 1) / DN_TW, (v.h[l] + DN_TH - 1) / DN_TH, count);
    int blur_id = flavor == SFE_HESSIAN ? 1 : (flavor == SFE_KLT ? 2 : -1);
    // KLT / brute flavours: the strip-streaming down stage (pyramid_stream.cu) when the geometry qualifies
    if (flavor != SFE_HESSIAN && launch_pyr_stream_down(v, 0, l, first, count, blur_id, 1.f, s)) {
    } else
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
