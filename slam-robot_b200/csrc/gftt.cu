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
#include "device_once.cuh"

namespace {

constexpr int EIG_USEFUL = 28;  // lanes 2..29; two lanes on each side cover Sobel (1) + box (1)
constexpr int SEL_THREADS = 1024;

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

  // The three bytes of the lane's pixel are loaded EIG_AHEAD arrivals before they are used: a strip walks its rows
  // serially (see above), so with few frames in flight -- one frame on a keyframe of the live robot -- every row would
  // otherwise wait for a full memory round trip (measured: 335 us for one VGA frame, 0.7 us per row).
  struct Bytes { int b, g, r; };
  auto arrival_row = [&](int j) -> int {      // gray row of arrival j: reflect101(j - 1) within [0, h)
    int row = j - 1;
    row = row < 0 ? -row : row;
    row = row >= a.h ? 2 * (a.h - 1) - row : row;
    return min(max(row, 0), a.h - 1);         // the clamp only matters for prefetches past the last arrival
  };
  auto load_row = [&](int j) -> Bytes {
    const uint8_t* p = px + (size_t)arrival_row(j) * a.row_stride;
    return Bytes{(int)__ldg(p), (int)__ldg(p + 1), (int)__ldg(p + 2)};
  };
  auto gray_of = [&](const Bytes& v) -> float {  // cvtColor(RGB2GRAY) on BGR bytes (matcher.cpp:313)
    return (float)((9798 * v.b + 19235 * v.g + 3735 * v.r + (1 << 14)) >> 15);
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
  constexpr int EIG_AHEAD = 6;          // a multiple of the unroll factor, so the ring slots are register names
  Bytes pre[EIG_AHEAD];
#pragma unroll
  for (int u = 0; u < EIG_AHEAD; ++u) pre[u] = load_row(u);
#pragma unroll 1
  for (int jb = 0; jb < nin; jb += EIG_AHEAD) {
#pragma unroll
    for (int u6 = 0; u6 < EIG_AHEAD; ++u6) {
      const int u = u6 % 3;
      const int j = jb + u6;            // arrival index; gray row = reflect101(j - 1) within [0, h)
      if (j >= nin) break;
      const float gnow = gray_of(pre[u6]);
      pre[u6] = load_row(j + EIG_AHEAD);
      win[u] = row_parts(gnow);
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

// ---- the same response map for a FEW frames (a keyframe of the live robot) -------------------------------------------
// gftt_eig_kernel walks each 28-column strip down the rows in one warp: right for a batch (one pass, everything in
// registers), but for one VGA frame that is 23 warps taking 482 serial row steps -- 224 us.  The only part that has to be
// serial is the running column sum; everything before it is independent per pixel.  So for small batches:
//   gftt_rowsum_kernel   tiled, one thread per pixel: gray -> Sobel row parts -> covariance products -> the three
//                        horizontal 3-sums R(y, x) in double, stored (24 bytes per pixel);
//   gftt_colsum_kernel   one thread per column and channel: S += R(y+1); box = (float)S; S -= R(y-1) down the rows -- two
//                        dependent double additions per row;
//   gftt_eigval_kernel   per pixel: the eigenvalue from the three box sums, and the frame maximum.
// Operation for operation what gftt_eig_kernel computes (same border rules, same order of the double additions), so
// the response maps are bit-identical; tests run both.
constexpr int RS_TW = 32, RS_TH = 8;

__global__ void __launch_bounds__(RS_TW * RS_TH) gftt_rowsum_kernel(const EigArgs a, double* __restrict__ R /* [count][3][h][w] */) {
  __shared__ float G[RS_TH + 2][RS_TW + 4];   // gray, rows y0-1 .. y0+TH (reflected), columns x0-2 .. x0+TW+1 (clamped)
  __shared__ float Pr[RS_TH + 2][RS_TW + 2];  // Sobel row parts at columns x0-1 .. x0+TW
  __shared__ float Pq[RS_TH + 2][RS_TW + 2];
  __shared__ float C[3][RS_TH][RS_TW + 2];    // covariance products at columns x0-1 .. x0+TW
  const int frame = blockIdx.z, x0 = blockIdx.x * RS_TW, y0 = blockIdx.y * RS_TH, tid = threadIdx.x;
  const uint8_t* src = a.bgr + (size_t)frame * a.frame_stride;
  const double sd = 1.0 / (4.0 * 3.0 * 255.0);
  const float k0 = (float)sd, k1 = (float)(2.0 * sd);
  for (int e = tid; e < (RS_TH + 2) * (RS_TW + 4); e += RS_TW * RS_TH) {
    const int i = e / (RS_TW + 4), j = e % (RS_TW + 4);
    int y = y0 - 1 + i;
    y = y < 0 ? -y : y;
    y = y >= a.h ? 2 * (a.h - 1) - y : y;
    y = min(max(y, 0), a.h - 1);                       // rows past the reflection belong to pixels outside the image
    const int x = min(max(x0 - 2 + j, 0), a.w - 1);
    const uint8_t* p = src + (size_t)y * a.row_stride + 3 * (size_t)x;
    G[i][j] = (float)((9798 * (int)__ldg(p) + 19235 * (int)__ldg(p + 1) + 3735 * (int)__ldg(p + 2) + (1 << 14)) >> 15);
  }
  __syncthreads();
  for (int e = tid; e < (RS_TH + 2) * (RS_TW + 2); e += RS_TW * RS_TH) {
    const int i = e / (RS_TW + 2), j = e % (RS_TW + 2), X = x0 - 1 + j;
    float gl = G[i][j], gr = G[i][j + 2];
    const float g = G[i][j + 1];
    if (X == 0) gl = gr;             // BORDER_REFLECT_101: g(-1) = g(1)
    if (X == a.w - 1) gr = gl;       //                     g(w) = g(w-2)
    Pr[i][j] = gr - gl;
    Pq[i][j] = fmaf(gr, k0, fmaf(g, k1, gl * k0));
  }
  __syncthreads();
  for (int e = tid; e < RS_TH * (RS_TW + 2); e += RS_TW * RS_TH) {
    const int i = e / (RS_TW + 2), j = e % (RS_TW + 2);
    const float dx = fmaf(Pr[i][j] + Pr[i + 2][j], k0, Pr[i + 1][j] * k1);
    const float dy = Pq[i + 2][j] - Pq[i][j];
    C[0][i][j] = dx * dx;
    C[1][i][j] = dx * dy;
    C[2][i][j] = dy * dy;
  }
  __syncthreads();
  const int i = tid / RS_TW, j = tid % RS_TW, X = x0 + j, Y = y0 + i;
  if (X < a.w && Y < a.h) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float cl = C[ch][i][j], cr = C[ch][i][j + 2];
      const float c = C[ch][i][j + 1];
      if (X == 0) cl = cr;
      if (X == a.w - 1) cr = cl;
      R[(((size_t)frame * 3 + ch) * a.h + Y) * a.w + X] = __dadd_rn(__dadd_rn((double)cl, (double)c), (double)cr);
    }
  }
}

// One thread per (column, covariance channel): the three running sums of a pixel are independent chains, so they run in
// three threads (60 warps per VGA frame instead of 20, a third of the instructions per row each).
__global__ void __launch_bounds__(128) gftt_colsum_kernel(const EigArgs a, const double* __restrict__ R, float* __restrict__ box /* [count][3][h][w] */) {
  const int frame = blockIdx.z, ch = blockIdx.y, x = blockIdx.x * 128 + threadIdx.x;
  if (x >= a.w) return;
  const size_t plane = (size_t)a.h * a.w;
  const double* r = R + ((size_t)frame * 3 + ch) * plane + x;
  float* out = box + ((size_t)frame * 3 + ch) * plane + x;
  auto row = [&](int y) -> double {         // R(y) with R(-1) = R(1) and R(h) = R(h-2)
    y = y < 0 ? -y : y;
    y = y >= a.h ? 2 * (a.h - 1) - y : y;
    return __ldg(r + (size_t)y * a.w);
  };
  constexpr int AHEAD = 16;                 // rows requested ahead of the running sum
  double nxt[AHEAD];
  double prev = row(-1), cur = row(0);
  double S = __dadd_rn(__dadd_rn(0.0, prev), cur);   // (0 + R(-1)) + R(0)
#pragma unroll
  for (int u = 0; u < AHEAD; ++u) nxt[u] = row(1 + u);
#pragma unroll 1
  for (int yb = 0; yb < a.h; yb += AHEAD) {
#pragma unroll
    for (int u = 0; u < AHEAD; ++u) {
      const int y = yb + u;
      if (y >= a.h) break;
      const double rp = nxt[u];             // R(y+1)
      nxt[u] = row(y + 1 + AHEAD);
      const double s = __dadd_rn(S, rp);
      out[(size_t)y * a.w] = (float)s;
      S = __dsub_rn(s, prev);               // minus R(y-1)
      prev = cur;
      cur = rp;
    }
  }
}

// Min eigenvalue of the 2x2 box-summed covariance (corner.cpp calcMinEigenVal) and the frame maximum.
__global__ void __launch_bounds__(256) gftt_eigval_kernel(const EigArgs a, const float* __restrict__ box) {
  const int frame = blockIdx.y;
  const size_t plane = (size_t)a.h * a.w, i = (size_t)blockIdx.x * 256 + threadIdx.x;
  float e = -3.0e38f;
  if (i < plane) {
    const float* b = box + (size_t)frame * 3 * plane + i;
    const float aa = b[0] * 0.5f, bb = b[plane], cc = b[2 * plane] * 0.5f;
    const float t = aa - cc;
    e = (aa + cc) - sqrtf(t * t + bb * bb);
    a.eig[(size_t)frame * plane + i] = e;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) e = fmaxf(e, __shfl_xor_sync(SFE_FULL, e, o));
  __shared__ float wmax[8];
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = wmax[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = fmaxf(m, wmax[k]);
    atomicMax(a.max_code + frame, order_code(m));
  }
}

struct NmsArgs {
  const float* eig;
  const unsigned* max_code;
  unsigned long long* keys;  // [count][cap]
  int* ncand;                // [count]
  int w, h, cap;
  double quality;
};

constexpr int NMS_ROWS = 16;  // rows per thread: a CTA covers 256 columns x 16 rows
__global__ void __launch_bounds__(256) gftt_nms_kernel(const NmsArgs a) {
  __shared__ unsigned long long s_keys[256 * NMS_ROWS];  // a CTA's candidates, appended to the frame's list in one piece
  __shared__ int s_n, s_base;
  const int frame = blockIdx.z;
  const int x = blockIdx.x * 256 + threadIdx.x, y0 = blockIdx.y * NMS_ROWS;
  const float thr = (float)((double)order_decode(a.max_code[frame]) * a.quality);  // threshold(THRESH_TOZERO), float compare
  const bool xin = x >= 1 && x < a.w - 1;  // featureselect.cpp scans interior pixels only
  const float* base = a.eig + (size_t)frame * a.h * a.w + min(max(x, 1), a.w - 2);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  // thresholded 3-pixel row maxima of the rows above / at / below, carried down the column
  auto tz = [&](float v) { return v > thr ? v : 0.f; };
  auto rowmax = [&](int y, float& centre) {
    const float* e = base + (size_t)y * a.w;
    centre = e[0];
    return fmaxf(fmaxf(tz(e[-1]), tz(centre)), tz(e[1]));
  };
  float c0, c1, c2;
  float m0 = rowmax(max(y0 - 1, 0), c0), m1 = rowmax(min(y0, a.h - 1), c1);
  for (int y = y0; y < min(y0 + NMS_ROWS, a.h - 1); ++y) {
    const float m2 = rowmax(y + 1, c2);
    const float v = c1;
    // dilate 3x3 of the thresholded map: v is a candidate iff it survives the threshold and equals the local maximum
    const bool cand = xin && y >= 1 && v > thr && v != 0.f && v == fmaxf(fmaxf(m0, m1), m2);
    const unsigned peers = __ballot_sync(SFE_FULL, cand);
    if (peers) {  // one shared-memory atomic per warp: the surviving lanes take consecutive slots
      const int leader = __ffs(peers) - 1;
      int slot0 = 0;
      if (lane == leader) slot0 = atomicAdd(&s_n, __popc(peers));
      slot0 = __shfl_sync(SFE_FULL, slot0, leader);
      if (cand) s_keys[slot0 + __popc(peers & ((1u << lane) - 1))] = ((unsigned long long)order_code(v) << 32) | (unsigned)(y * a.w + x);
    }
    m0 = m1; m1 = m2; c1 = c2;
  }
  __syncthreads();
  if (threadIdx.x == 0) s_base = s_n ? atomicAdd(a.ncand + frame, s_n) : 0;  // one global atomic per CTA
  __syncthreads();
  for (int i = threadIdx.x; i < s_n; i += 256)
    if (s_base + i < a.cap) a.keys[(size_t)frame * a.cap + s_base + i] = s_keys[i];
}

struct SelArgs {
  unsigned long long* keys;
  const int* ncand;
  float* corners;  // [count][max_corners][2]
  int* ncorners;   // [count]
  int w, cap, max_corners;
  double min_distance;
};

// One CTA per frame.  Bitonic sort, descending: stages whose partner distance is below SEL_CHUNK run on
// chunks staged in shared memory (one load and one store per chunk and pass), the few wider ones in global
// memory (a frame's keys stay in L2).
constexpr int SEL_CHUNK = 16384;
__device__ __forceinline__ void bitonic_chunk(unsigned long long* sk, int chunk_base, int len, int k_lo, int k_hi, int j_hi, int tid) {
  // all stages (k, j) with k in [k_lo, k_hi] and j <= min(k/2, j_hi), on `len` keys whose global index starts at chunk_base
  for (int k = k_lo; k <= k_hi; k <<= 1)
    for (int j = min(k >> 1, j_hi); j > 0; j >>= 1) {
      for (int t = tid; t < (len >> 1); t += SEL_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const unsigned long long ki = sk[i], kl = sk[l];
        const bool desc = ((chunk_base + i) & k) == 0;
        if (desc ? ki < kl : ki > kl) { sk[i] = kl; sk[l] = ki; }
      }
      __syncthreads();
    }
}

// Greedy minimum-distance selection (featureselect.cpp) over the first `n` keys of the descending list `list`: a
// candidate is kept unless an accepted corner lies closer than min_distance.  Sequential in the accepted corners only;
// warp 0 does it.  Returns the number of corners (in every lane).
__device__ __forceinline__ int greedy_select(const SelArgs& a, const unsigned long long* list, int n, float* acc, float* out, int lane) {
  const bool check = a.min_distance >= 1.0;
  const double md2 = a.min_distance * a.min_distance;
  int cnt = 0;
  // 32 candidates at a time, one per lane.  First every lane tests its candidate against the corners accepted in earlier
  // batches (no dependency between lanes); then the survivors are settled in list order: the first one is accepted, the
  // later ones test against it, and so on -- one round per ACCEPTED corner instead of one per candidate (round 1 walked
  // the ~1000 candidates of a frame one by one: 75 of the 95 us of a keyframe's selection).
  for (int c0 = 0; c0 < n && cnt < a.max_corners; c0 += 32) {
    const int c = c0 + lane;
    bool alive = c < n;
    float fx = 0.f, fy = 0.f;
    if (alive) {
      const unsigned ofs = (unsigned)list[c];
      const int y = ofs / a.w, x = ofs - y * a.w;
      fx = (float)x;
      fy = (float)y;
    }
    if (check && alive) {
      bool bad = false;
      for (int k = 0; k < cnt; ++k) {
        const float dx = fx - acc[2 * k], dy = fy - acc[2 * k + 1];
        bad |= (double)(dx * dx + dy * dy) < md2;
      }
      alive = !bad;
    }
    unsigned m = __ballot_sync(SFE_FULL, alive);
    while (m != 0u && cnt < a.max_corners) {
      const int j = __ffs(m) - 1;
      const float jx = __shfl_sync(SFE_FULL, fx, j), jy = __shfl_sync(SFE_FULL, fy, j);
      if (lane == j) {
        acc[2 * cnt] = fx;
        acc[2 * cnt + 1] = fy;
        out[2 * cnt] = fx;
        out[2 * cnt + 1] = fy;
      }
      ++cnt;
      if (check && alive && lane > j) {
        const float dx = fx - jx, dy = fy - jy;
        if ((double)(dx * dx + dy * dy) < md2) alive = false;
      }
      m = __ballot_sync(SFE_FULL, alive && lane > j);
    }
    __syncwarp();
  }
  return cnt;
}

// The selection stops after max_corners corners, which on a textured frame takes the first few hundred of ~20 000
// candidates -- sorting all of them is most of a single frame's latency.  So the kernel first tries a PREFIX: a
// histogram over 12 bits of the response code (exponent + 4 mantissa bits, 16 bins per octave) finds the cut bin above
// which about SEL_TOPK candidates lie; only those are compacted into shared memory, sorted and fed to the greedy
// selection.  If that yields max_corners corners the result is, by construction, the one the full list gives (the
// prefix of a descending list does not depend on what follows it).  Otherwise -- few strong corners, a large
// max_corners -- the kernel falls back to sorting everything.
constexpr int SEL_TOPK = 1024;
constexpr int SEL_BINS = 4096;
__device__ __forceinline__ int sel_bin(unsigned long long key) { return (int)((key >> (32 + 19)) & (SEL_BINS - 1)); }

__global__ void __launch_bounds__(SEL_THREADS) gftt_select_kernel(const SelArgs a) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  float* acc = reinterpret_cast<float*>(sel_smem);                                                    // [max_corners][2]
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(sel_smem + ((8 * (size_t)a.max_corners + 15) & ~(size_t)15));
  __shared__ int s_hist[SEL_BINS];
  __shared__ int s_warp[SEL_THREADS / 32];
  __shared__ int s_cut, s_m, s_cnt;
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  unsigned long long* keys = a.keys + (size_t)frame * a.cap;
  float* out = a.corners + (size_t)frame * a.max_corners * 2;
  if (a.ncand[frame] > a.cap) {  // candidate list overflowed: report instead of returning a wrong list
    if (tid == 0) a.ncorners[frame] = -1;
    return;
  }
  const int n = a.ncand[frame];

  // ---- prefix attempt
  if (n > 4 * SEL_TOPK) {
    for (int i = tid; i < SEL_BINS; i += SEL_THREADS) s_hist[i] = 0;
    if (tid == 0) { s_cut = 0; s_m = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) atomicAdd(&s_hist[sel_bin(keys[i])], 1);
    __syncthreads();
    // suffix sums from the top bin: thread t owns bins 4t..4t+3 counted from the top
    int own[4], mine = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { own[q] = s_hist[SEL_BINS - 1 - (4 * tid + q)]; mine += own[q]; }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(SFE_FULL, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < (tid >> 5); ++w) before += s_warp[w];
    int run = before + incl - mine;   // candidates in the bins above this thread's
    if (run < SEL_TOPK) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int prev = run;
        run += own[q];
        if (prev < SEL_TOPK && run >= SEL_TOPK) s_cut = SEL_BINS - 1 - (4 * tid + q);   // exactly one thread and bin
      }
    }
    __syncthreads();
    const int cut = s_cut;   // 0 when all candidates together are fewer than SEL_TOPK: everything passes
    for (int i = tid; i < n; i += SEL_THREADS) {
      const unsigned long long k = keys[i];
      if (sel_bin(k) >= cut) {
        const int pos = atomicAdd(&s_m, 1);
        if (pos < SEL_CHUNK) sk[pos] = k;
      }
    }
    __syncthreads();
    const int m = s_m;
    if (m <= SEL_CHUNK) {
      int mp2 = 1;
      while (mp2 < m) mp2 <<= 1;
      for (int i = m + tid; i < mp2; i += SEL_THREADS) sk[i] = 0ull;
      __syncthreads();
      bitonic_chunk(sk, 0, mp2, 2, mp2, mp2, tid);
      if (tid < 32) {
        const int cnt = greedy_select(a, sk, m, acc, out, lane);
        if (tid == 0) s_cnt = cnt;
      }
      __syncthreads();
      if (s_cnt >= a.max_corners || m == n) {
        if (tid == 0) a.ncorners[frame] = s_cnt;
        return;
      }
    }
    __syncthreads();
  }

  // ---- the whole list
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = n + tid; i < np2; i += SEL_THREADS) keys[i] = 0ull;  // pad: sorts to the end (descending)
  __syncthreads();
  const int chunk = min(np2, SEL_CHUNK);
  // pass 1: every chunk fully sorted (alternating directions come from the global index)
  for (int c0 = 0; c0 < np2; c0 += chunk) {
    for (int i = tid; i < chunk; i += SEL_THREADS) sk[i] = keys[c0 + i];
    __syncthreads();
    bitonic_chunk(sk, c0, chunk, 2, chunk, chunk, tid);
    for (int i = tid; i < chunk; i += SEL_THREADS) keys[c0 + i] = sk[i];
    __syncthreads();
  }
  // wider merges: global steps while the partner distance spans chunks, then one shared-memory pass per chunk
  for (int k = chunk << 1; k <= np2; k <<= 1) {
    for (int j = k >> 1; j >= chunk; j >>= 1) {
      for (int t = tid; t < (np2 >> 1); t += SEL_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const unsigned long long ki = keys[i], kl = keys[l];
        const bool desc = (i & k) == 0;
        if (desc ? ki < kl : ki > kl) { keys[i] = kl; keys[l] = ki; }
      }
      __syncthreads();
    }
    for (int c0 = 0; c0 < np2; c0 += chunk) {
      for (int i = tid; i < chunk; i += SEL_THREADS) sk[i] = keys[c0 + i];
      __syncthreads();
      bitonic_chunk(sk, c0, chunk, k, k, chunk >> 1, tid);
      for (int i = tid; i < chunk; i += SEL_THREADS) keys[c0 + i] = sk[i];
      __syncthreads();
    }
  }
  if (tid >= 32) return;
  const int cnt = greedy_select(a, keys, n, acc, out, lane);
  if (tid == 0) a.ncorners[frame] = cnt;
}

}  // namespace

// Enqueues the kernels for `count` frames.  Workspace (device): eig [count*w*h] floats, max_code [count],
// ncand [count], keys [count*cap], rowsums [count*3*h*w] doubles followed by as many floats, or NULL (selects the
// one-pass response kernel).
// Returns the number of launches or a negative cudaError.
int launch_good_features(const uint8_t* bgr, size_t row_stride, size_t frame_stride, int w, int h, int count, int max_corners,
                         double quality, double min_distance, float* eig, unsigned* max_code, int* ncand,
                         unsigned long long* keys, int cap, float* corners, int* ncorners, double* rowsums, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(max_code, 0, sizeof(unsigned) * count, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(ncand, 0, sizeof(int) * count, s);
  if (e != cudaSuccess) return -(int)e;
  EigArgs ea{bgr, row_stride, frame_stride, eig, max_code, w, h, (w + EIG_USEFUL - 1) / EIG_USEFUL};
  int launches = 3;
  if (rowsums && h >= 3) {  // a few frames: the two-pass form (the caller provides 24 bytes per pixel of workspace for it)
    gftt_rowsum_kernel<<<dim3((w + RS_TW - 1) / RS_TW, (h + RS_TH - 1) / RS_TH, count), RS_TW * RS_TH, 0, s>>>(ea, rowsums);
    float* box = reinterpret_cast<float*>(rowsums + 3 * (size_t)w * h * count);   // the caller's workspace holds both
    gftt_colsum_kernel<<<dim3((w + 127) / 128, 3, count), 128, 0, s>>>(ea, rowsums, box);
    gftt_eigval_kernel<<<dim3((unsigned)(((size_t)w * h + 255) / 256), count), 256, 0, s>>>(ea, box);
    launches = 5;
  } else {
    gftt_eig_kernel<<<ea.strips * count, 32, 0, s>>>(ea);
  }
  NmsArgs na{eig, max_code, keys, ncand, w, h, cap, quality};
  gftt_nms_kernel<<<dim3((w + 255) / 256, (h + NMS_ROWS - 1) / NMS_ROWS, count), 256, 0, s>>>(na);
  SelArgs sa{keys, ncand, corners, ncorners, w, cap, max_corners, min_distance};
  const size_t smem = ((sizeof(float) * 2 * (size_t)max_corners + 15) & ~(size_t)15) + sizeof(unsigned long long) * SEL_CHUNK;
  static PerDeviceOnce attr{};
  if (first_use_on_device(attr))
    cudaFuncSetAttribute(gftt_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  // + 17 KB static (histogram) stays under 227 KB
  gftt_select_kernel<<<count, SEL_THREADS, smem, s>>>(sa);
  e = cudaGetLastError();
  return e == cudaSuccess ? launches : -(int)e;
}
