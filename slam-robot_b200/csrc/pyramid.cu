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
#include "sfe_common.cuh"

namespace {

constexpr int L0_TW = 64, L0_TH = 16, L0_THREADS = 256;
constexpr int DN_TW = 32, DN_TH = 16, DN_THREADS = 256;

__device__ __forceinline__ float gray_f32(const uint8_t* px) {
  // cvtColor(RGB2GRAY) on BGR bytes (hessian.h:100) then convertTo(CV_32F, 1/255.) (hessian.h:101)
  int g = (9798 * px[0] + 19235 * px[1] + 3735 * px[2] + (1 << 14)) >> 15;
  return (float)g * (float)(1. / 255.);
}

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

}  // namespace

int launch_pyr_build(const PyrView& v, int flavor, const uint8_t* bgr, size_t row_stride,
                     size_t frame_stride, int first, int count, cudaStream_t s) {
  int launches = 0;
  dim3 g0((v.w[0] + L0_TW - 1) / L0_TW, (v.h[0] + L0_TH - 1) / L0_TH, count);
  if (flavor == SFE_HESSIAN) pyr_l0_kernel<SFE_HESSIAN><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  else if (flavor == SFE_KLT) pyr_l0_kernel<SFE_KLT><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  else pyr_l0_kernel<SFE_BRUTE><<<g0, L0_THREADS, 0, s>>>(v, bgr, row_stride, frame_stride, first);
  ++launches;
  for (int l = 1; l < v.depth; ++l) {
    dim3 g((v.w[l] + DN_TW - 1) / DN_TW, (v.h[l] + DN_TH - 1) / DN_TH, count);
    int blur_id = flavor == SFE_HESSIAN ? 1 : (flavor == SFE_KLT ? 2 : -1);
    pyr_down_kernel<<<g, DN_THREADS, 0, s>>>(v.base[0][l - 1], v.frame_stride[l - 1], v.w[l - 1], v.h[l - 1],
                                             v.pitch[l - 1], v.base[0][l], v.frame_stride[l], v.w[l], v.h[l],
                                             v.pitch[l], first, blur_id, 1.f);
    ++launches;
    if (flavor == SFE_KLT) {
      for (int p = 1; p <= 2; ++p) {  // klt.h:123-124
        pyr_down_kernel<<<g, DN_THREADS, 0, s>>>(v.base[p][l - 1], v.frame_stride[l - 1], v.w[l - 1], v.h[l - 1],
                                                 v.pitch[l - 1], v.base[p][l], v.frame_stride[l], v.w[l],
                                                 v.h[l], v.pitch[l], first, -1, 2.f);
        ++launches;
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? launches : -(int)e;
}
