// seed.cu -- the two small per-element steps either side of the hot path.
//
//   seed_features_kernel   the head of the FindMatches loop, matcher.cpp:224-245, one thread per feature: pyramid
//                          levels from the map point's uncertainty, the search seed from Frame::Project
//                          (localmap.cpp:18-26 -> project.h:11-54: Eigen's quaternion transform, pinhole + radial
//                          distortion, all in double, operation for operation -- the library is built without
//                          FMA contraction) and the out-of-bounds gate.
//   yuyv_to_bgr_kernel     the integer YUYV -> BGR conversion of the V4L2 capture path, video.cpp:187-223, four
//                          pixels (8 bytes in, 12 bytes out) per thread.
// Both are plain streaming kernels; arithmetic = oracle.c orc_seed_features / orc_yuyv_to_bgr.
#include "sfe_common.cuh"

namespace {

struct Pose {
  double rot[4], trans[3], k[7];
};

__global__ void __launch_bounds__(256) seed_features_kernel(int n, const double* __restrict__ points4,
                                                            const double* __restrict__ uncertainty, const Pose pose,
                                                            const float* __restrict__ from_xy, int cols, int rows,
                                                            float* __restrict__ seed_xy, int32_t* __restrict__ levels,
                                                            uint8_t* __restrict__ go) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = from_xy[2 * i], y = from_xy[2 * i + 1];
  const double unc = uncertainty[i];
  levels[i] = unc > 100 ? 6 : 3;  // matcher.cpp:227-229
  if (unc < 100) {                // matcher.cpp:234
    const double* pt = points4 + 4 * (size_t)i;
    const double w = pt[3];
    const double vx = pt[0] - pose.trans[0] * w, vy = pt[1] - pose.trans[1] * w, vz = pt[2] - pose.trans[2] * w;
    const double qx = pose.rot[0], qy = pose.rot[1], qz = pose.rot[2], qw = pose.rot[3];
    double ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;  // Eigen: uv = q.vec x v; uv += uv
    ux += ux; uy += uy; uz += uz;
    const double px = (vx + qw * ux) + (qy * uz - qz * uy);                         // v + w*uv + q.vec x uv
    const double py = (vy + qw * uy) + (qz * ux - qx * uz);
    const double pz = (vz + qw * uz) + (qx * uy - qy * ux);
    if (!(pz < 0.001 * w)) {  // project.h:27: points behind the lens keep from_pt as the seed
      double xp = px / pz, yp = py / pz;
      const double r2 = xp * xp + yp * yp;
      const double distort = 1.0 + r2 * (pose.k[0] + r2 * (pose.k[1] + r2 * pose.k[2]));
      xp *= distort; yp *= distort;
      xp *= pose.k[3]; yp *= pose.k[4];
      xp += pose.k[5]; yp += pose.k[6];
      x = (float)xp;
      y = (float)yp;
    }
  }
  seed_xy[2 * i] = x;
  seed_xy[2 * i + 1] = y;
  go[i] = !(x < 0 || y < 0 || x >= (float)cols || y > (float)rows);  // matcher.cpp:243 (`>` on y, preserved)
}

__device__ __forceinline__ int sat8(int c) { return min(max(c, 0), 255); }  // video.cpp:186 SAT

__global__ void __launch_bounds__(256) yuyv_to_bgr_kernel(const uint2* __restrict__ in, size_t ngroups, uint32_t* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one group = 8 input bytes = 4 pixels
  if (t >= ngroups) return;
  const uint2 v = __ldg(in + t);
  uint8_t px[12];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t wv = h ? v.y : v.x;
    const int y1 = wv & 0xff, u = (int)((wv >> 8) & 0xff) - 128, y2 = (wv >> 16) & 0xff, vv = (int)(wv >> 24) - 128;
    const int cb = (u * 454) >> 8;
    const int cr = (vv * 359) >> 8;
    const int cg = (u * 88 + vv * 183) >> 8;
    px[6 * h + 0] = (uint8_t)sat8(y1 + cb); px[6 * h + 1] = (uint8_t)sat8(y1 - cg); px[6 * h + 2] = (uint8_t)sat8(y1 + cr);
    px[6 * h + 3] = (uint8_t)sat8(y2 + cb); px[6 * h + 4] = (uint8_t)sat8(y2 - cg); px[6 * h + 5] = (uint8_t)sat8(y2 + cr);
  }
  uint32_t* o = out + 3 * t;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    o[k] = (uint32_t)px[4 * k] | ((uint32_t)px[4 * k + 1] << 8) | ((uint32_t)px[4 * k + 2] << 16) | ((uint32_t)px[4 * k + 3] << 24);
}

}  // namespace

int launch_seed_features(int n, const double* points4, const double* uncertainty, const double* rot, const double* trans,
                         const double* k, const float* from_xy, int cols, int rows, float* seed_xy, int32_t* levels,
                         uint8_t* go, cudaStream_t s) {
  if (n <= 0) return 0;
  Pose p;
  for (int i = 0; i < 4; ++i) p.rot[i] = rot[i];
  for (int i = 0; i < 3; ++i) p.trans[i] = trans[i];
  for (int i = 0; i < 7; ++i) p.k[i] = k[i];
  seed_features_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, points4, uncertainty, p, from_xy, cols, rows, seed_xy, levels, go);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}

int launch_yuyv_to_bgr(const uint8_t* yuyv, size_t npixels, uint8_t* bgr, cudaStream_t s) {
  const size_t ngroups = npixels / 4;
  if (ngroups == 0) return 0;
  yuyv_to_bgr_kernel<<<(unsigned)((ngroups + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint2*>(yuyv), ngroups,
                                                                        reinterpret_cast<uint32_t*>(bgr));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}
