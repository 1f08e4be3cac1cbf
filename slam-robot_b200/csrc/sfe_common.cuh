// sfe_common.cuh -- shared device/host definitions of libslamfe (sm_100a only).
//
// Floating-point contract: the library is compiled with -fmad=false, so the compiler never
// fuses a multiply with an add on its own.  Every fused multiply-add is spelled fmaf() and
// every other operation rounds once (IEEE), exactly as the CPU oracle (oracle/oracle.c,
// built with -ffp-contract=off) -- tracked positions are a chaotic function of the last bits
// of the patch scores, so "close" is not good enough.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slamfe.h"

#define SFE_PLEN 169      // 13*13
#define SFE_SLOTS 6       // ceil(169/32) patch pixels per lane
#define SFE_FULL 0xffffffffu

// Device-side view of a batch of pyramids (passed by value to kernels).
// Plane p of level l of frame f starts at base[p][l] + f*frame_stride[l]; rows are `pitch[l]`
// floats apart.  pitch is w rounded up to 4 floats so rows stay 16-byte aligned.
struct PyrView {
  int depth;
  int batch;
  int w[SFE_MAX_LEVELS];
  int h[SFE_MAX_LEVELS];
  int pitch[SFE_MAX_LEVELS];
  long long frame_stride[SFE_MAX_LEVELS];  // floats
  float* base[3][SFE_MAX_LEVELS];          // [plane][level]; planes 1,2 only for SFE_KLT
};

struct ImgView {  // one plane of one frame of one level
  const float* p;
  int w, h, pitch;
};

__host__ __device__ inline ImgView img_of(const PyrView& v, int plane, int level, int frame) {
  ImgView r;
  r.p = v.base[plane][level] + (long long)frame * v.frame_stride[level];
  r.w = v.w[level];
  r.h = v.h[level];
  r.pitch = v.pitch[level];
  return r;
}

__device__ __forceinline__ int reflect101(int i, int n) {
  // cv::BORDER_REFLECT_101; loops only for images narrower than the filter reach
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
  return i;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// Butterfly sum: lane 0 (and every lane) ends with the pairwise tree 16,8,4,2,1 of the
// per-lane partials -- the order declared in oracle/oracle.h.
__device__ __forceinline__ float warp_sum(float v) {
  v = v + __shfl_xor_sync(SFE_FULL, v, 16);
  v = v + __shfl_xor_sync(SFE_FULL, v, 8);
  v = v + __shfl_xor_sync(SFE_FULL, v, 4);
  v = v + __shfl_xor_sync(SFE_FULL, v, 2);
  v = v + __shfl_xor_sync(SFE_FULL, v, 1);
  return v;
}

// Packed FP32 (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE single-precision operations per instruction, each rounded
// exactly like its scalar form).  The tracking and pyramid kernels are bound by instruction issue, not by the FMA
// pipe, so their element-wise work runs on register pairs (the tracker pairs shifts: .x = shift 2p, .y = shift 2p+1;
// the pyramid pairs neighbouring columns).  A scalar operand is written {s, s}; ptxas turns that into the
// instruction's broadcast form, no move is issued.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 both(float s) { return make_float2(s, s); }

// ---- internal launch interface (capi.cu <-> kernels) -----------------------------------
struct TrackArgs {
  int n, n_per_pair, from_first, to_first;
  const float* from_xy;
  float* to_xy;
  const int32_t* levels;
  int default_levels;
  float thr;
  int maxit;
  double fb_max;  // matcher.cpp:201 compares a double norm with the double literal 0.3
  float* back_xy;
  int32_t* status_fwd;
  int32_t* status_bwd;
  uint8_t* accepted;
  int32_t* steps;
  int ndir;  // 2: forward + backward (matcher.cpp:173-206); 1: forward only (TrackFeature, hessian.h:243-264)
};

// each returns the number of kernels launched (negative cudaError on failure)
int launch_pyr_build(const PyrView& v, int flavor, const uint8_t* bgr, size_t row_stride,
                     size_t frame_stride, int first, int count, cudaStream_t s);
int launch_pyr_stream_hessian(const PyrView& v, const uint8_t* bgr, size_t row_stride, size_t frame_stride, int first,
                              int count, cudaStream_t s, int* launches);
int launch_pyr_stream_klt_l0(const PyrView& v, const uint8_t* bgr, size_t row_stride, size_t frame_stride, int first, int count,
                             cudaStream_t s);
int launch_pyr_stream_down(const PyrView& v, int plane, int l, int first, int count, int blur_id, float scale, cudaStream_t s);
int launch_track_hessian(const PyrView& from, const PyrView& to, const TrackArgs& a,
                         const float* mask, int* counter, int num_sms, cudaStream_t s);
int launch_get_patches(const PyrView& v, int frame, int level, int n, const float* xy,
                       float* patches, float* mean, float* sumsq, cudaStream_t s);
int launch_brute_hessian(const PyrView& tv, int tframe, const PyrView& sv, int sframe, int level,
                         int n, const float* txy, const float* xy, float* out7, const float* mask,
                         cudaStream_t s);
int launch_track_klt(const PyrView& from, const PyrView& to, const TrackArgs& a, const float* mask,
                     int* counter, int num_sms, cudaStream_t s);
int launch_klt_system(const PyrView& tv, int tframe, const PyrView& sv, int sframe, int level, int n,
                      const float* txy, const float* xy, float* out24, const float* mask,
                      cudaStream_t s);
int launch_brute_track(const PyrView& from, const PyrView& to, int from_first, int to_first, int n,
                       int n_per_pair, const float* from_xy, float* to_xy, const float* coarse,
                       int n_coarse, const float* fine, int n_fine, int32_t* status, float* best_sad,
                       unsigned long long* positions, cudaStream_t s);
int hamming_set_impl(int impl);  // hamming.cu: 0 by size, 1 ALU kernel, 2 tensor-core kernel; returns the previous value
int launch_hamming256(const uint32_t* q, int nq, const uint32_t* t, int nt, int batch, int ratio_num,
                      int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass,
                      void** ws, size_t* ws_cap, cudaStream_t s);
int launch_good_features(const uint8_t* bgr, size_t row_stride, size_t frame_stride, int w, int h, int count, int max_corners,
                         double quality, double min_distance, float* eig, unsigned* max_code, int* ncand,
                         unsigned long long* keys, int cap, float* corners, int* ncorners, double* rowsums, cudaStream_t s);
int launch_seed_features(int n, const double* points4, const double* uncertainty, const double* rot, const double* trans,
                         const double* k, const float* from_xy, int cols, int rows, float* seed_xy, int32_t* levels,
                         uint8_t* go, cudaStream_t s);
int launch_yuyv_to_bgr(const uint8_t* yuyv, size_t npixels, uint8_t* bgr, cudaStream_t s);
