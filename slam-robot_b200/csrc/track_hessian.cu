// track_hessian.cu -- P1, the live path: HessianTracker forward/backward (track_impl.cuh,
// MODE_HESSIAN) plus the GetPatch / BruteHessian parity accessors.
#include "track_impl.cuh"

namespace {

// GetPatch (hessian.h:54-93) for n points of one level
__global__ void get_patches_kernel(PyrView v, int frame, int level, int n, const float* __restrict__ xy,
                                   float* __restrict__ patches, float* __restrict__ mean, float* __restrict__ sumsq) {
  __shared__ WarpScratch scratch[TRK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, i = blockIdx.x * TRK_WARPS + warp;
  if (i >= n) return;
  init_scratch(scratch[warp], lane);
  float T[SFE_SLOTS], mk[SFE_SLOTS], d[6], m = 0.f, q = 0.f;
  load_mask(nullptr, lane, mk);
  evaluate<MODE_HESSIAN>(scratch[warp], img_of(v, 0, level, frame), true, T, m, q, mk, xy[2 * i], xy[2 * i + 1], lane, d);
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k)
    if (lane + 32 * k < SFE_PLEN) patches[(size_t)i * SFE_PLEN + lane + 32 * k] = T[k];
  if (lane == 0) { mean[i] = m; sumsq[i] = q; }
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
  float mk[SFE_SLOTS];
  load_mask(mask, lane, mk);
  float T[SFE_SLOTS], m = 0.f, q = 0.f, d[6];
  evaluate<MODE_HESSIAN>(scratch[warp], img_of(tv, 0, level, tframe), true, T, m, q, mk, txy[2 * i], txy[2 * i + 1], lane, d);
  float s0 = evaluate<MODE_HESSIAN>(scratch[warp], img_of(sv, 0, level, sframe), false, T, m, q, mk, xy[2 * i],
                                    xy[2 * i + 1], lane, d);
  if (lane == 0) {
    out7[7 * i] = s0;
    for (int k = 0; k < 6; ++k) out7[7 * i + 1 + k] = d[k];
  }
}

}  // namespace

int launch_track_hessian(const PyrView& from, const PyrView& to, const TrackArgs& a, const float* mask,
                         int* counter, int num_sms, cudaStream_t s) {
  return launch_track_fb<MODE_HESSIAN>(from, to, a, mask, counter, num_sms, s);
}

int launch_get_patches(const PyrView& v, int frame, int level, int n, const float* xy, float* patches,
                       float* mean, float* sumsq, cudaStream_t s) {
  if (n <= 0) return 0;
  get_patches_kernel<<<(n + TRK_WARPS - 1) / TRK_WARPS, 32 * TRK_WARPS, 0, s>>>(v, frame, level, n, xy, patches, mean, sumsq);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

int launch_brute_hessian(const PyrView& tv, int tframe, const PyrView& sv, int sframe, int level, int n,
                         const float* txy, const float* xy, float* out7, const float* mask, cudaStream_t s) {
  if (n <= 0) return 0;
  brute_hessian_kernel<<<(n + TRK_WARPS - 1) / TRK_WARPS, 32 * TRK_WARPS, 0, s>>>(tv, tframe, sv, sframe, level, n,
                                                                                  txy, xy, out7, mask);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}
