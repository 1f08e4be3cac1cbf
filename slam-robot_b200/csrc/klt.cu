// klt.cu -- P2: KLTTracker (klt.h) forward/backward (track_impl.cuh, MODE_KLT) and the
// symmetric-KLT normal equations of klt.h:286-353.
//
// As written, klt.h computes the 2x2 system U d = e from the warp sums A,B,C,RS,VW and then
// overwrites the step with the finite-difference Newton step (klt.h:355-380), so the tracked
// position depends on the image plane only; the tracker below follows that.  The system itself
// is exposed by klt_system_kernel: one warp per feature, sixteen warp-shuffle reductions over
// the 169 patch pixels, then the 2x2 inverse / LU solve in every lane.
#include "track_impl.cuh"

namespace {

__device__ __forceinline__ void full_patch(const ImgView& im, float x, float y, const LanePix& lp, float (&v)[SFE_SLOTS]) {
  PatchGeom g;
  g.x = axis_geom(x, false, true);
  g.y = axis_geom(y, false, false);
  sample_patch_global(im, g, lp, v);
}

__global__ void __launch_bounds__(128) klt_system_kernel(PyrView tv, int tframe, PyrView sv, int sframe, int level, int n,
                                                         const float* __restrict__ txy, const float* __restrict__ xy,
                                                         float* __restrict__ out24, const float* __restrict__ mask) {
  const int lane = threadIdx.x & 31, i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const LanePix lp = lane_pix(lane);
  float mk[SFE_SLOTS];
  load_mask(mask, lane, mk);
  float I[SFE_SLOTS], Igx[SFE_SLOTS], Igy[SFE_SLOTS], J[SFE_SLOTS], Jgx[SFE_SLOTS], Jgy[SFE_SLOTS];
  const float tx = txy[2 * i], ty = txy[2 * i + 1], x = xy[2 * i], y = xy[2 * i + 1];
  full_patch(img_of(tv, 0, level, tframe), tx, ty, lp, I);
  full_patch(img_of(tv, 1, level, tframe), tx, ty, lp, Igx);
  full_patch(img_of(tv, 2, level, tframe), tx, ty, lp, Igy);
  full_patch(img_of(sv, 0, level, sframe), x, y, lp, J);
  full_patch(img_of(sv, 1, level, sframe), x, y, lp, Jgx);
  full_patch(img_of(sv, 2, level, sframe), x, y, lp, Jgy);
  float Im, Iq, Jm, Jq;
  patch_stats(I, Im, Iq);
  patch_stats(J, Jm, Jq);
  const float alpha = sqrtf(Iq / Jq);   // klt.h:289
  const float beta = Im - alpha * Jm;   // klt.h:290
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
#pragma unroll
  for (int k = 0; k < SFE_SLOTS; ++k) {
    if (J[k] == 0.f || I[k] == 0.f) continue;  // klt.h:303 (also skips slots beyond 169)
    const float m = mk[k];
    const float Jv = fmaf(J[k], alpha, beta);
    const float gI0 = Igx[k], gI1 = Igy[k];
    const float gJ0 = Jgx[k] * alpha, gJ1 = Jgy[k] * alpha;
    const float diff = (I[k] - Jv) * m;
    acc[0] = fmaf(gI0 * gI0, m, acc[0]);   acc[1] = fmaf(gI0 * gI1, m, acc[1]);
    acc[2] = fmaf(gI1 * gI0, m, acc[2]);   acc[3] = fmaf(gI1 * gI1, m, acc[3]);
    acc[4] = fmaf(gI0 * gJ0, m, acc[4]);   acc[5] = fmaf(gI0 * gJ1, m, acc[5]);
    acc[6] = fmaf(gI1 * gJ0, m, acc[6]);   acc[7] = fmaf(gI1 * gJ1, m, acc[7]);
    acc[8] = fmaf(gJ0 * gJ0, m, acc[8]);   acc[9] = fmaf(gJ0 * gJ1, m, acc[9]);
    acc[10] = fmaf(gJ1 * gJ0, m, acc[10]); acc[11] = fmaf(gJ1 * gJ1, m, acc[11]);
    acc[12] = fmaf(diff, gI0, acc[12]);    acc[13] = fmaf(diff, gI1, acc[13]);
    acc[14] = fmaf(diff, gJ0, acc[14]);    acc[15] = fmaf(diff, gJ1, acc[15]);
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = warp_sum(acc[k]);
  const float* A = acc; const float* B = acc + 4; const float* C = acc + 8;
  const float* RS = acc + 12; const float* VW = acc + 14;
  const float lambda = .0001f;
  // Di = (B^T)^-1, Eigen 2x2 closed form (klt.h:326)
  const float t00 = B[0], t01 = B[2], t10 = B[1], t11 = B[3];
  const float det = t00 * t11 - t10 * t01;
  const float invdet = 1.f / det;
  const float D0 = t11 * invdet, D1 = -t01 * invdet, D2 = -t10 * invdet, D3 = t00 * invdet;
  const float Al0 = A[0] + lambda, Al1 = A[1], Al2 = A[2], Al3 = A[3] + lambda;
  const float M0 = Al0 * D0 + Al1 * D2, M1 = Al0 * D1 + Al1 * D3, M2 = Al2 * D0 + Al3 * D2, M3 = Al2 * D1 + Al3 * D3;
  float U[4] = {(M0 * C[0] + M1 * C[2]) - 0.5f * B[0], (M0 * C[1] + M1 * C[3]) - 0.5f * B[1],
                (M2 * C[0] + M3 * C[2]) - 0.5f * B[2], (M2 * C[1] + M3 * C[3]) - 0.5f * B[3]};  // klt.h:330
  float e[2] = {(M0 * VW[0] + M1 * VW[1]) - 0.5f * RS[0], (M2 * VW[0] + M3 * VW[1]) - 0.5f * RS[1]};  // klt.h:331
  // U.lu().solve(e) (klt.h:343): 2x2 partial-pivot LU
  float u00 = U[0], u01 = U[1], u10 = U[2], u11 = U[3], e0 = e[0], e1 = e[1];
  if (fabsf(u10) > fabsf(u00)) {
    float t;
    t = u00; u00 = u10; u10 = t;
    t = u01; u01 = u11; u11 = t;
    t = e0; e0 = e1; e1 = t;
  }
  const float l = u10 / u00;
  const float w11 = u11 - l * u01;
  const float y1 = e1 - l * e0;
  const float d1 = y1 / w11;
  const float d0 = (e0 - u01 * d1) / u00;
  if (lane == 0) {
    float* o = out24 + 24 * (size_t)i;
    for (int k = 0; k < 16; ++k) o[k] = acc[k];
    for (int k = 0; k < 4; ++k) o[16 + k] = U[k];
    o[20] = e[0]; o[21] = e[1]; o[22] = d0; o[23] = d1;
  }
}

}  // namespace

int launch_track_klt(const PyrView& from, const PyrView& to, const TrackArgs& a, const float* mask, int* counter,
                     int num_sms, cudaStream_t s) {
  return launch_track_fb<MODE_KLT>(from, to, a, mask, counter, num_sms, s);
}

int launch_klt_system(const PyrView& tv, int tframe, const PyrView& sv, int sframe, int level, int n,
                      const float* txy, const float* xy, float* out24, const float* mask, cudaStream_t s) {
  if (n <= 0) return 0;
  klt_system_kernel<<<(n + 3) / 4, 128, 0, s>>>(tv, tframe, sv, sframe, level, n, txy, xy, out24, mask);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}
