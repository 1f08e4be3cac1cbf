// gftt.cu -- corner seeding: cv::goodFeaturesToTrack(grey, corners, maxCorners, quality, minDistance) with its
// defaults (min-eigenvalue response, blockSize 3, Sobel 3), as matcher.cpp:123-130 calls it on the RGB2GRAY image
// of matcher.cpp:313 -- SURVEY.md 8f rank 1, the step that follows tracking on keyframes.
//
// Arithmetic = oracle/oracle.c orc_min_eigen_val / orc_good_features, which is pinned bit-for-bit against
// cv2 4.13 (response map and corner lists).  Three kernels per batch of frames:
//   gftt_eig_kernel     one warp per (frame, 28-column strip), one pixel per lane, walking down the rows:
//                       gray -> Sobel (row parts exchanged by shuffles, three-row register windows) ->
//                       covariance products -> 3x3 box sums in double -- OpenCV's RUNNING column sum
//                       (S += row(y+1); out; S -= row(y-1)), whose history from the top of the image shows
//                       in the last bit, so a strip is never split into row bands -> min eigenvalue;
//                       also the per-frame maximum (atomicMax on an order-preserving integer code);
//   gftt_nms_kernel     threshold at quality*max, 3x3 non-maximum suppression on interior pixels, candidates
//                       appended as 64-bit keys (response code << 32 | pixel offset);
//   gftt_select_kernel  one CTA per frame: bitonic sort of the keys (descending: larger response first, ties by
//                       higher offset -- featureselect.cpp's greaterThanPtr), then the sequential greedy
//                       minimum-distance selection, each candidate tested against the accepted corners by all
//                       threads at once.
#include "ctx.cuh"

namespace {

constexpr int EIG_USEFUL = 28;  // lanes 2..29; two lanes on each side cover Sobel (1) + box (1)
constexpr int SEL_THREADS = 512;

__device__ __forceinline__ unsigned order_code(float f) {  // monotone float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float order_decode(unsigned c) {
  return __uint_as_float((c & 0x80000000u) ? (c & 0x7fffffffu) : ~c);
}

struct EigArgs {
  const uint8_t* bgr;
  size_t row_stride, frame_stride;
  float* eig;          // [count][h][w]
  unsigned* max_code;  // [count]
  int w, h, strips;
};

// Horizontal parts of the two Sobel filters on one gray row: r = g(x+1) - g(x-1) (exact),
// q = fma(g(x+1), s, fma(g(x), 2s, g(x-1)*s)) (the fused SIMD form of OpenCV's row filter).
struct RowParts { float r, q; };

__global__ void __launch_bounds__(32) gftt_eig_kernel(const EigArgs a) {
  const int lane = threadIdx.x;
  const int strip = blockIdx.x % a.strips, frame = blockIdx.x / a.strips;
  const int x = EIG_USEFUL * strip - 2 + lane;
  const bool inimg = x >= 0 && x < a.w;
  const bool ledge = x == 0, redge = x == a.w - 1;
  const bool useful = inimg && lane >= 2 && lane <= 29;
  const int xc = min(max(x, 0), a.w - 1);
  const uint8_t* px = a.bgr + (size_t)frame * a.frame_stride + 3 * (size_t)xc;
  float* out = a.eig + ((size_t)frame * a.h) * a.w + xc;
  const double sd = 1.0 / (4.0 * 3.0 * 255.0);
  const float k0 = (float)sd, k1 = (float)(2.0 * sd);

  auto gray_row = [&](int row) -> float {  // cvtColor(RGB2GRAY) on BGR bytes (matcher.cpp:313)
    const uint8_t* p = px + (size_t)row * a.row_stride;
    const int g = (9798 * (int)__ldg(p) + 19235 * (int)__ldg(p + 1) + 3735 * (int)__ldg(p + 2) + (1 << 14)) >> 15;
    return (float)g;
  };
  auto row_parts = [&](float g) -> RowParts {
    float gl = __shfl_up_sync(SFE_FULL, g, 1), gr = __shfl_down_sync(SFE_FULL, g, 1);
    if (ledge) gl = gr;  // BORDER_REFLECT_101: g(-1) = g(1)
    if (redge) gr = gl;  //                     g(w) = g(w-2)
    RowParts p;
    p.r = gr - gl;
    p.q = fmaf(gr, k0, fmaf(g, k1, gl * k0));
    return p;
  };

  // Input rows arrive in the order 1, 0, 1, 2, ..., h-1, h-2 (reflection at both ends); after the third
  // arrival every new row completes one Sobel row, i.e. one row of covariance row sums R.
  RowParts win[3];
  double R[3][3];  // ring of the last three covariance row sums [slot][channel]
  double S[3] = {0.0, 0.0, 0.0};
  float vmax = -3.0e38f;
  const int nin = a.h + 2;
#pragma unroll 1
  for (int jb = 0; jb < nin; jb += 3) {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int j = jb + u;             // arrival index; gray row = reflect101(j - 1) within [0, h)
      if (j >= nin) break;
      int row = j - 1;
      row = row < 0 ? -row : row;
      row = row >= a.h ? 2 * (a.h - 1) - row : row;
      win[u] = row_parts(gray_row(row));
      if (j < 2) continue;
      // Sobel row yc = j-2 from arrivals j-2, j-1, j = slots (u+1)%3, (u+2)%3, u
      const RowParts &p0 = win[(u + 1) % 3], &p1 = win[(u + 2) % 3], &p2 = win[u];
      const float dx = fmaf(p0.r + p2.r, k0, p1.r * k1);
      const float dy = p2.q - p0.q;
      float c[3] = {dx * dx, dx * dy, dy * dy};
      const int yc = j - 2;             // covariance row just produced; its ring slot is (yc % 3) == (u+1)%3
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float cl = __shfl_up_sync(SFE_FULL, c[ch], 1), cr = __shfl_down_sync(SFE_FULL, c[ch], 1);
        if (ledge) cl = cr;
        if (redge) cr = cl;
        R[(u + 1) % 3][ch] = __dadd_rn(__dadd_rn((double)cl, (double)c[ch]), (double)cr);
      }
      if (yc == 0) continue;
      // running column sum: on arrival of R(yc) emit box row y = yc-1:  D = S + R(y+1);  S = D - R(y-1)
      // slots: R(yc) -> (u+1)%3, R(yc-1) -> u%3 (previous), R(yc-2) -> (u+2)%3
      auto emit = [&](int y, const double (&Rp)[3], const double (&Rm)[3]) {
        float box[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const double s = __dadd_rn(S[ch], Rp[ch]);
          box[ch] = (float)s;
          S[ch] = __dsub_rn(s, Rm[ch]);
        }
        const float aa = box[0] * 0.5f, bb = box[1], cc = box[2] * 0.5f;
        const float t = aa - cc;
        const float e = (aa + cc) - sqrtf(t * t + bb * bb);  // corner.cpp calcMinEigenVal
        if (useful) {
          out[(size_t)y * a.w] = e;
          vmax = fmaxf(vmax, e);
        }
      };
      if (yc == 1) {
        // S = (0 + R(-1)) + R(0) with R(-1) = R(1); then row 0: D = S + R(1), S = D - R(-1)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) S[ch] = __dadd_rn(__dadd_rn(0.0, R[(u + 1) % 3][ch]), R[u % 3][ch]);
        emit(0, R[(u + 1) % 3], R[(u + 1) % 3]);
      } else {
        emit(yc - 1, R[(u + 1) % 3], R[(u + 2) % 3]);
      }
      if (yc == a.h - 1) emit(a.h - 1, R[u % 3], R[u % 3]);  // last row: R(h) = R(h-2), S -= R(h-2) is moot
    }
  }
  // per-frame maximum (cv::minMaxLoc over the whole map)
#pragma unroll
  for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(SFE_FULL, vmax, o));
  if (lane == 0) atomicMax(a.max_code + frame, order_code(vmax));
}

struct NmsArgs {
  const float* eig;
  const unsigned* max_code;
  unsigned long long* keys;  // [count][cap]
  int* ncand;                // [count]
  int w, h, cap;
  double quality;
};

__global__ void __launch_bounds__(256) gftt_nms_kernel(const NmsArgs a) {
  const int frame = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x < 1 || y < 1 || x >= a.w - 1 || y >= a.h - 1) return;  // featureselect.cpp scans interior pixels only
  const float* e = a.eig + ((size_t)frame * a.h + y) * a.w + x;
  const float thr = (float)((double)order_decode(a.max_code[frame]) * a.quality);  // threshold(THRESH_TOZERO), float compare
  const float v = e[0];
  if (!(v > thr)) return;
  // dilate 3x3 of the thresholded map: neighbours at or below the threshold count as 0
  float m = v;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const float n = e[dy * a.w + dx];
      m = fmaxf(m, n > thr ? n : 0.f);
    }
  if (v != m || v == 0.f) return;
  const int slot = atomicAdd(a.ncand + frame, 1);
  if (slot < a.cap)
    a.keys[(size_t)frame * a.cap + slot] = ((unsigned long long)order_code(v) << 32) | (unsigned)(y * a.w + x);
}

struct SelArgs {
  unsigned long long* keys;
  const int* ncand;
  float* corners;  // [count][max_corners][2]
  int* ncorners;   // [count]
  int w, cap, max_corners;
  double min_distance;
};

// One CTA per frame.  Keys are sorted in place in global memory (the working set of a frame stays in L2).
__global__ void __launch_bounds__(SEL_THREADS) gftt_select_kernel(const SelArgs a) {
  const int frame = blockIdx.x, tid = threadIdx.x;
  unsigned long long* keys = a.keys + (size_t)frame * a.cap;
  if (a.ncand[frame] > a.cap) {  // candidate list overflowed: report instead of returning a wrong list
    if (tid == 0) a.ncorners[frame] = -1;
    return;
  }
  const int n = a.ncand[frame];
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = n + tid; i < np2; i += SEL_THREADS) keys[i] = 0ull;  // pad: sorts to the end (descending)
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += SEL_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long ki = keys[i], kl = keys[l];
          const bool desc = (i & k) == 0;
          if (desc ? ki < kl : ki > kl) { keys[i] = kl; keys[l] = ki; }
        }
      }
      __syncthreads();
    }
  // greedy selection (featureselect.cpp): a candidate is kept unless an accepted corner lies closer than
  // min_distance.  Sequential by nature; one warp walks the sorted candidates and tests each against the
  // accepted corners 32 at a time.
  extern __shared__ float acc[];  // [max_corners][2]
  if (tid >= 32) return;
  float* out = a.corners + (size_t)frame * a.max_corners * 2;
  const bool check = a.min_distance >= 1.0;
  const double md2 = a.min_distance * a.min_distance;
  int cnt = 0;
  for (int c = 0; c < n && cnt < a.max_corners; ++c) {
    const unsigned ofs = (unsigned)keys[c];
    const int y = ofs / a.w, x = ofs - y * a.w;
    bool bad = false;
    if (check)
      for (int k = tid; k < cnt; k += 32) {
        const float dx = (float)x - acc[2 * k], dy = (float)y - acc[2 * k + 1];
        bad |= (double)(dx * dx + dy * dy) < md2;
      }
    if (__any_sync(SFE_FULL, bad)) continue;
    if (tid == 0) {
      acc[2 * cnt] = (float)x;
      acc[2 * cnt + 1] = (float)y;
      out[2 * cnt] = (float)x;
      out[2 * cnt + 1] = (float)y;
    }
    ++cnt;
    __syncwarp();
  }
  if (tid == 0) a.ncorners[frame] = cnt;
}

}  // namespace

// Enqueues the three kernels for `count` frames.  Workspace (device): eig [count*w*h] floats, max_code [count],
// ncand [count], keys [count*cap].  Returns the number of launches or a negative cudaError.
int launch_good_features(const uint8_t* bgr, size_t row_stride, size_t frame_stride, int w, int h, int count, int max_corners,
                         double quality, double min_distance, float* eig, unsigned* max_code, int* ncand,
                         unsigned long long* keys, int cap, float* corners, int* ncorners, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(max_code, 0, sizeof(unsigned) * count, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(ncand, 0, sizeof(int) * count, s);
  if (e != cudaSuccess) return -(int)e;
  EigArgs ea{bgr, row_stride, frame_stride, eig, max_code, w, h, (w + EIG_USEFUL - 1) / EIG_USEFUL};
  gftt_eig_kernel<<<ea.strips * count, 32, 0, s>>>(ea);
  NmsArgs na{eig, max_code, keys, ncand, w, h, cap, quality};
  gftt_nms_kernel<<<dim3((w + 31) / 32, (h + 7) / 8, count), 256, 0, s>>>(na);
  SelArgs sa{keys, ncand, corners, ncorners, w, cap, max_corners, min_distance};
  const size_t smem = sizeof(float) * 2 * (size_t)max_corners;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(gftt_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  gftt_select_kernel<<<count, SEL_THREADS, smem, s>>>(sa);
  e = cudaGetLastError();
  return e == cudaSuccess ? 3 : -(int)e;
}
