// brute.cu -- P3: BruteTracker (brute.h), the reference's only "brute-force matching": an
// exhaustive multi-resolution window search minimising the alpha/beta-normalised SSD between
// the 13x13 template and the patch at every candidate position (brute.h:82-117,129-164).
//
// One CTA per feature.  For each {window,res} pass of the schedule the CTA stages the window's
// whole footprint ((2w+16)^2 pixels, replicate-clamped) in shared memory once, the candidate
// positions are dealt round-robin to the CTA's warps, each warp evaluates a position
// cooperatively (bilinear 13x13 sample, patch statistics, SSD -- same lane/tree summation order
// as the oracle) and keeps its running arg-min; a block reduction applies the reference's
// "last minimum wins" rule (`if (sad > best) continue`, brute.h:108).
// The candidate offsets reproduce the reference's float loop counters (x += res accumulates
// rounding, brute.h:105-106): they are generated sequentially by one thread per pass.
#include "patch.cuh"

namespace {

constexpr int BR_WARPS = 8;
constexpr int BR_THREADS = 32 * BR_WARPS;
constexpr int BR_MAXOFF = 2048;        // offsets per pass (the 8/0.01 debug pass has 1601)
constexpr int BR_TDIM = 36;            // staged footprint is at most BR_TDIM x BR_TDIM (window <= 10)
constexpr int BR_TS = BR_TDIM + 1;
constexpr int BR_MAXSCHED = 8;

struct BruteSched {
  float coarse[2 * BR_MAXSCHED];
  float fine[2 * BR_MAXSCHED];
  int n_coarse, n_fine;
};

struct StridedTileFetch {
  const float* t;
  int ox, oy;
  __device__ __forceinline__ float operator()(int Y, int X) const { return t[(Y - oy) * BR_TS + (X - ox)]; }
};

struct Best {
  float sad;
  int p;
};

// sequential semantics of brute.h:102-113 for one more candidate
__device__ __forceinline__ void consider(Best& b, float sad, int p) {
  if (!(sad > b.sad)) { b.sad = sad; b.p = p; }
}
// merge of two disjoint, internally ordered candidate subsequences (no NaNs): smaller sad wins,
// equal sad -> later position wins
__device__ __forceinline__ Best merge(Best a, Best b) {
  if (b.p < 0) return a;
  if (a.p < 0) return b;
  if (b.sad < a.sad || (b.sad == a.sad && b.p > a.p)) return b;
  return a;
}

__global__ void __launch_bounds__(BR_THREADS) brute_track_kernel(PyrView from, PyrView to, int from_first, int to_first,
                                                                 int n, int n_per_pair, const float* __restrict__ from_xy,
                                                                 float* __restrict__ to_xy, BruteSched sched,
                                                                 int32_t* __restrict__ status, float* __restrict__ best_sad,
                                                                 unsigned long long* __restrict__ positions) {
  __shared__ float tile[BR_TDIM * BR_TS];
  __shared__ float offs[BR_MAXOFF];
  __shared__ int s_cnt;
  __shared__ float s_cx, s_cy, s_sad, s_last;
  __shared__ Best s_best[BR_WARPS];
  __shared__ int s_nan[BR_WARPS];

  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LanePix lp = lane_pix(lane);
  const int pair = i / n_per_pair;
  const int ff = from_first + pair, tf = to_first + pair;
  const float fx = from_xy[2 * i], fy = from_xy[2 * i + 1];
  const float x0 = to_xy[2 * i], y0 = to_xy[2 * i + 1];
  const int lv = min(from.depth, to.depth);
  unsigned long long npos = 0;

  // brute.h:137-142
  const float margin = 13.f;
  if (x0 < margin || y0 < margin || (x0 + margin) > (float)to.w[0] || (y0 + margin) > (float)to.h[0]) {
    if (threadIdx.x == 0) {
      if (status) status[i] = SFE_OUT_OF_BOUNDS;
      if (best_sad) best_sad[i] = 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) {
    s_cx = x0 * (float)(1. / (1 << (lv - 1)));  // brute.h:144
    s_cy = y0 * (float)(1. / (1 << (lv - 1)));
    s_sad = 0.f;
  }
  int result = SFE_OK;
  for (int level = lv - 1; level >= 0 && result == SFE_OK; --level) {
    const float sc = (float)(1. / (1 << level));
    const ImgView tim = img_of(from, 0, level, ff), sim = img_of(to, 0, level, tf);
    // template patch (brute.h:120-127), one copy per warp
    float T[SFE_SLOTS], tmean, tsumsq;
    {
      PatchGeom g;
      g.x = axis_geom(fx * sc, false, true);
      g.y = axis_geom(fy * sc, false, false);
      sample_patch_global(tim, g, lp, T);
      patch_stats(T, tmean, tsumsq);
    }
    const int npass = level > 0 ? sched.n_coarse : sched.n_fine;
    const float* sp = level > 0 ? sched.coarse : sched.fine;
    for (int pass = 0; pass < npass; ++pass) {
      const float window = sp[2 * pass], res = sp[2 * pass + 1];
      __syncthreads();
      if (threadIdx.x == 0) {  // the reference's float loop counter, verbatim
        int c = 0;
        for (float x = -window; x <= window && c < BR_MAXOFF; x += res) offs[c++] = x;
        s_cnt = c;
      }
      __syncthreads();
      const int cnt = s_cnt;
      const float cx = s_cx, cy = s_cy;
      // stage the footprint of every candidate: columns floor(cx-window)-7 .. floor(cx+window)+8
      const int ox = (int)floorf(cx - window) - 7, oy = (int)floorf(cy - window) - 7;
      const int tw = (int)floorf(cx + window) + 8 - ox + 1, th = (int)floorf(cy + window) + 8 - oy + 1;
      const bool staged = tw <= BR_TDIM && th <= BR_TDIM;
      if (staged) {
        for (int e = threadIdx.x; e < tw * th; e += BR_THREADS) {
          int r = e / tw, c = e - r * tw;
          tile[r * BR_TS + c] = __ldg(sim.p + (long long)clampi(oy + r, 0, sim.h - 1) * sim.pitch + clampi(ox + c, 0, sim.w - 1));
        }
      }
      __syncthreads();
      Best best{1e6f, -1};
      int nan_seen = 0;
      float last_sad = 0.f;
      const int total = cnt * cnt;
      for (int p = warp; p < total; p += BR_WARPS) {
        const int ixo = p / cnt, iyo = p - ixo * cnt;  // x is the outer loop (brute.h:105-106)
        const float px = cx + offs[ixo], py = cy + offs[iyo];
        PatchGeom g;
        g.x = axis_geom(px, false, true);
        g.y = axis_geom(py, false, false);
        float v[SFE_SLOTS];
        if (staged) {
          StridedTileFetch f{tile, ox, oy};
          const bool interior = g.x.i0 >= 0 && g.x.i0 + SFE_PATCH <= sim.w - 1 && g.y.i0 >= 0 && g.y.i0 + SFE_PATCH <= sim.h - 1;
          if (interior) {
            const float w0 = g.x.a1 * g.y.a1, w1 = g.x.a * g.y.a1, w2 = g.x.a1 * g.y.a, w3 = g.x.a * g.y.a;
            const float* t0 = tile + (g.y.i0 - oy) * BR_TS + (g.x.i0 - ox);
#pragma unroll
            for (int k = 0; k < SFE_SLOTS; ++k) {
              float r = 0.f;
              if (lp.pr[k] < SFE_PATCH) {
                const float* t = t0 + lp.pr[k] * BR_TS + lp.pc[k];
                r = fmaf(t[BR_TS + 1], w3, fmaf(t[BR_TS], w2, fmaf(t[1], w1, t[0] * w0)));
              }
              v[k] = r;
            }
          } else {
#pragma unroll
            for (int k = 0; k < SFE_SLOTS; ++k) {
              float r = 0.f;
              if (lp.pr[k] < SFE_PATCH)
                r = sample_general(f, g.x.i0 + lp.pc[k], g.y.i0 + lp.pr[k], sim.w, sim.h, g.x.a, g.x.a1, g.y.a, g.y.a1);
              v[k] = r;
            }
          }
        } else {
          sample_patch_global(sim, g, lp, v);
        }
        float m, q;
        patch_stats(v, m, q);
        const float alpha = sqrtf(tsumsq / q);  // brute.h:83-84
        const float beta = tmean - alpha * m;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < SFE_SLOTS; ++k) {
          float diff = fmaf(-v[k], alpha, T[k]) - beta;
          float t = fmaf(diff, diff, s);
          s = (T[k] == 0.f || v[k] == 0.f) ? s : t;
        }
        const float sad = warp_sum(s);
        nan_seen |= (sad != sad);
        if (p == total - 1) last_sad = sad;
        consider(best, sad, p);
      }
      if (lane == 0) {
        s_best[warp] = best;
        s_nan[warp] = nan_seen;
        if (total > 0 && (total - 1) % BR_WARPS == warp) s_last = last_sad;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        Best b{1e6f, -1};
        int any_nan = 0;
        for (int w = 0; w < BR_WARPS; ++w) {
          b = merge(b, s_best[w]);
          any_nan |= s_nan[w];
        }
        // a NaN score is never "> best", so it is accepted and every later candidate with it
        if (any_nan) { b.p = total - 1; b.sad = s_last; }
        if (b.p >= 0) {
          const int bx = b.p / cnt, by = b.p - bx * cnt;
          s_cx = cx + offs[bx];  // brute.h:110-111
          s_cy = cy + offs[by];
        }
        s_sad = b.sad;
      }
      npos += (unsigned long long)total;
    }
    __syncthreads();
    if (s_sad > 100.f) result = SFE_OUT_OF_BOUNDS;  // brute.h:149,159
    if (level > 0 && result == SFE_OK) {
      __syncthreads();
      if (threadIdx.x == 0) { s_cx *= 2.f; s_cy *= 2.f; }  // brute.h:151
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (result == SFE_OK) { to_xy[2 * i] = s_cx; to_xy[2 * i + 1] = s_cy; }
    if (status) status[i] = result;
    if (best_sad) best_sad[i] = s_sad;
    if (positions) atomicAdd(positions, npos);
  }
}

}  // namespace

int launch_brute_track(const PyrView& from, const PyrView& to, int from_first, int to_first, int n, int n_per_pair,
                       const float* from_xy, float* to_xy, const float* coarse, int n_coarse, const float* fine,
                       int n_fine, int32_t* status, float* best_sad, unsigned long long* positions, cudaStream_t s) {
  if (n <= 0) return 0;
  BruteSched sc;
  sc.n_coarse = n_coarse;
  sc.n_fine = n_fine;
  for (int k = 0; k < 2 * BR_MAXSCHED; ++k) {
    sc.coarse[k] = k < 2 * n_coarse ? coarse[k] : 0.f;
    sc.fine[k] = k < 2 * n_fine ? fine[k] : 0.f;
  }
  brute_track_kernel<<<n, BR_THREADS, 0, s>>>(from, to, from_first, to_first, n, n_per_pair, from_xy, to_xy, sc, status,
                                              best_sad, positions);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 1 : -(int)e;
}
