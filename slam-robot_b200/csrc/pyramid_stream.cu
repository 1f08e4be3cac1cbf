// pyramid_stream.cu -- MakePyramid (hessian.h:95-126) as a register-streaming pipeline.
//
// Why: the tiled kernels in pyramid.cu spend their time on instruction issue (index arithmetic, shared
// memory round trips, 1.3-1.4x halo recomputation) and re-read level 0 from HBM to build level 1.
// Here one WARP owns a column strip of a frame (128 level-0 columns, 112 of them useful) and walks down
// its rows; nothing is staged in shared memory:
//   * horizontal filters exchange the two or three neighbouring pixels with warp shuffles;
//   * vertical filters keep their last five rows in registers (the row loop is unrolled ten-fold so the
//     window slots are compile-time register names);
//   * level 0 (gray -> /255 -> 5x5 blur) and level 1 (pyrDown -> 5x5 blur) are fused: the level-0 row a
//     lane has just produced feeds the level-1 pipeline directly, so the BGR bytes are read once and
//     every level is written once (algorithmic HBM traffic); deeper levels run the same down stage
//     from the previous level's floats.
// Borders (cv::BORDER_REFLECT_101): above the image the strip simply starts eight rows early on the
// reflected rows -- every filter is symmetric in its taps, so the extension reproduces the reflected
// values bit for bit.  That trick does not carry across a downsampling step at the right/bottom edge
// (the mirror axes of the two grids differ), so there the last in-image lane takes its out-of-image
// neighbours from the mirrored pixels it already holds, and the last two level-1 rows take their
// out-of-image pyrDown rows from the mirrored rows still in the window.
// The arithmetic per pixel is pyr_math.cuh's, i.e. bit-identical to the tiled kernels and the oracle -- including
// OpenCV's scalar-path operation orders in the columns that take them: with widths that are multiples of 8 (the
// only ones these kernels accept) those are the horizontal pyrDown pass in output column 0 and in the last
// columns behind its vector body (pd_h_tail), and the level-1 column blur when w/2 is not a multiple of 8
// (blur_col_tail).  Only the warps that own such columns execute the second form (a warp-uniform branch).
#include <stdlib.h>

#include "pyr_math.cuh"
#include "device_once.cuh"

namespace {

constexpr int STRIP_USEFUL = 112;  // level-(l-1) columns a strip contributes: lanes 2..29, 4 columns each
constexpr int STRIP_HALO = 8;      // two dead lanes on each side absorb the 8-column reach of blur o pyrDown o blur

struct StreamArgs {
  const uint8_t* bgr;      // FROM_BGR input: 8-bit interleaved frames (frame 0 = first frame of this call)
  size_t row_stride, frame_stride;
  const float* in;         // !FROM_BGR input: level l-1 planes of the pyramid batch
  long long in_fs;
  int in_pitch;
  float* out0;             // FROM_BGR: level 0 planes
  long long out0_fs;
  int out0_pitch;
  float* out1;             // the downsampled level
  long long out1_fs;
  int out1_pitch;
  int w, h;                // size of the input level (level 0 for FROM_BGR); w % 4 == 0
  int w1, h1;              // size of the downsampled level: w/2, (h+1)/2
  int first;               // first frame slot in the pyramid batch
  int strips, bands, band_rows, nunits;
  float scale;             // DOWN_ONLY: factor applied to the pyrDown result (klt.h:123-124 doubles the gradients)
  int hbody;               // pd_hbody(w): last output column of cv::pyrDown's horizontal vector body
  float *out_gx, *out_gy;  // klt_l0_stream_kernel: the Scharr planes of level 0 (frame stride / pitch of out0)
};

__device__ __forceinline__ float gray_px(uint32_t p) {
  // cvtColor(RGB2GRAY) on BGR bytes (hessian.h:100), then convertTo(CV_32F, 1/255.) (hessian.h:101)
  unsigned s = __dp2a_lo(9798u | (19235u << 16), p, 1u << 14);
  s = __dp2a_hi(3735u, p, s);
  return (float)(int)(s >> 15) * (float)(1. / 255.);
}

__device__ __forceinline__ int reflect_row(int t, int h) {
  t = t < 0 ? -t : t;
  t = t >= h ? 2 * (h - 1) - t : t;
  return min(max(t, 0), h - 1);  // the clamp only matters for ticks whose results are never stored
}

// Input rows reach the lanes through a per-warp shared-memory ring filled by cp.async (LDGSTS) eight ticks
// ahead: the copies are asynchronous and ordered only by commit/wait groups, so neither nvcc nor ptxas can
// sink them next to their first use (plain or volatile loads were rescheduled five ticks late, exposing the
// full DRAM latency on every tick -- profiles/README.md).  Each lane copies and later reads only its own
// 12 (BGR) or 16 (float4) bytes, so no warp synchronisation is needed; out-of-image lanes copy zeros.
constexpr int RING = 10, PREFETCH = 8;

template <bool FROM_BGR>
__device__ __forceinline__ void issue_row(uint32_t slot_addr, const StreamArgs& a, const uint8_t* bgr_px, const float* in_px,
                                          int t, int src_bytes) {
  const int ry = reflect_row(t, a.h);
  if constexpr (FROM_BGR) {
    const uint8_t* p = bgr_px + (size_t)ry * a.row_stride;
    asm volatile(
        "cp.async.ca.shared.global [%0], [%1], 4, %2;\n\t"
        "cp.async.ca.shared.global [%0+4], [%1+4], 4, %2;\n\t"
        "cp.async.ca.shared.global [%0+8], [%1+8], 4, %2;\n\t"
        "cp.async.commit_group;" ::"r"(slot_addr), "l"(p), "r"(src_bytes) : "memory");
  } else {
    const float* p = in_px + (size_t)ry * a.in_pitch;
    asm volatile(
        "cp.async.ca.shared.global [%0], [%1], 16, %2;\n\t"
        "cp.async.commit_group;" ::"r"(slot_addr), "l"(p), "r"(src_bytes) : "memory");
  }
}

// One warp = one (frame, row band, column strip) unit.
// BLUR0 / BLUR1 select the Gaussian tap sets at compile time: as immediates the taps let FMUL/FFMA issue at the full
// rate (the three-register forms issue every other cycle, and this kernel is FP-pipe bound).
// DOWN_ONLY (with !FROM_BGR): the level is the bare pyrDown of its input times a.scale -- the gradient planes of klt.h
// (:123-124) and every level of brute.h (:72-77); the blur stage is skipped and the pyrDown row itself is stored.
template <bool FROM_BGR, int BLUR0, int BLUR1, bool DOWN_ONLY = false>
__global__ void __launch_bounds__(32) pyr_stream_kernel(const StreamArgs a) {
  const int lane = threadIdx.x;
  const int unit = blockIdx.x;
  const int strip = unit % a.strips;
  const int band = (unit / a.strips) % a.bands;
  const int frame = unit / (a.strips * a.bands);

  const int g = STRIP_USEFUL * strip - STRIP_HALO + 4 * lane;  // first input column of this lane
  const bool inimg = g >= 0 && g < a.w;                        // w % 4 == 0: the group is all in or all out
  const bool ledge = g == 0, redge = g == a.w - 4;
  const bool useful = inimg && lane >= 2 && lane <= 29;
  // OpenCV's scalar-path columns (pyr_math.cuh): this lane's two pyrDown columns g/2, g/2+1, and the level-1 column blur
  const bool tail_x = inimg && pd_h_tail(g >> 1, a.hbody), tail_y = inimg && pd_h_tail((g >> 1) + 1, a.hbody);
  const bool warp_tail = __any_sync(SFE_FULL, tail_x || tail_y);
  const bool ctail = inimg && blur_col_tail(g >> 1, a.w1);
  const bool warp_ctail = __any_sync(SFE_FULL, ctail);
  const int r0 = band * a.band_rows;                           // band of input rows [r0, r1); band_rows is even
  const int r1 = min(r0 + a.band_rows, 2 * a.h1);
  const int q_hi = min(r1, a.h), j_lo = r0 >> 1, j_hi = min(r1 >> 1, a.h1);
  const Taps k0 = taps_for(BLUR0), k1 = taps_for(BLUR1);

  const uint8_t* bgr_px = nullptr;
  const float* in_px = nullptr;
  float* out0_px = nullptr;
  if constexpr (FROM_BGR) {
    bgr_px = a.bgr + (size_t)frame * a.frame_stride + 3 * (size_t)max(g, 0);
    out0_px = a.out0 + (long long)(a.first + frame) * a.out0_fs + max(g, 0);
  } else {
    in_px = a.in + (long long)(a.first + frame) * a.in_fs + max(g, 0);
  }
  float* out1_px = a.out1 + (long long)(a.first + frame) * a.out1_fs + (max(g, 0) >> 1);

  // Ticks: at tick t the input row t arrives (FROM_BGR: its horizontally blurred gray row, which completes
  // level-0 row q = t-2; otherwise the previous level's row q = t).  pyrDown row i needs rows 2i-2..2i+2 and is
  // formed at q = 2i+2; downsampled row j needs pyrDown rows j-2..j+2 and is formed with i = j+2.
  constexpr int LAG = FROM_BGR ? 2 : 0;
  const int t_begin = r0 - 8;  // even, so q is even exactly when the unrolled tick index is
  const int t_last = r1 + 4 + LAG;
  constexpr int LANE_BYTES = FROM_BGR ? 12 : 16;
  __shared__ __align__(16) unsigned char ring[RING][32 * LANE_BYTES];
  const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(&ring[0][lane * LANE_BYTES]);
  const int src_bytes = inimg ? (FROM_BGR ? 4 : 16) : 0;
#pragma unroll
  for (int u = 0; u < PREFETCH; ++u)
    issue_row<FROM_BGR>(ring_addr + u * 32 * LANE_BYTES, a, bgr_px, in_px, t_begin + u, src_bytes);

  float4 hbw[5];   // FROM_BGR: horizontally blurred gray rows t-4..t
  float2 phw[5];   // horizontally pyrDown-filtered rows q-4..q
  float2 bhw[5];   // horizontally blurred pyrDown rows i-4..i
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    hbw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    phw[u] = bhw[u] = make_float2(0.f, 0.f);
  }

#pragma unroll 1
  for (int tb = t_begin; tb <= t_last; tb += 10) {
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int t = tb + u;
      const int q = t - LAG;
      // row t has landed once at most PREFETCH-1 younger copy groups are still in flight; then refill the slot
      // that was consumed two ticks ago with row t + PREFETCH
      asm volatile("cp.async.wait_group %0;" ::"n"(PREFETCH - 1) : "memory");
      issue_row<FROM_BGR>(ring_addr + ((u + PREFETCH) % RING) * 32 * LANE_BYTES, a, bgr_px, in_px, t + PREFETCH, src_bytes);
      float4 row;
      if constexpr (FROM_BGR) {
        const uint32_t* rr = reinterpret_cast<const uint32_t*>(&ring[u % RING][lane * LANE_BYTES]);
        const uint32_t ra_ = rr[0], rb_ = rr[1], rc_ = rr[2];
        float4 gr;
        gr.x = gray_px(ra_);
        gr.y = gray_px(__funnelshift_r(ra_, rb_, 24));
        gr.z = gray_px(__funnelshift_r(rb_, rc_, 16));
        gr.w = gray_px(rc_ >> 8);
        // GaussianBlur rows (sigma 1.1): columns g-2, g-1 from the left lane, g+4, g+5 from the right lane
        float lz = __shfl_up_sync(SFE_FULL, gr.z, 1), lw = __shfl_up_sync(SFE_FULL, gr.w, 1);
        float rx = __shfl_down_sync(SFE_FULL, gr.x, 1), ry = __shfl_down_sync(SFE_FULL, gr.y, 1);
        if (ledge) { lz = gr.z; lw = gr.y; }   // columns -2, -1 mirror to 2, 1
        if (redge) { rx = gr.z; ry = gr.y; }   // columns w, w+1 mirror to w-2, w-3
        float4 hb;
        hb.x = blur_row(lz, lw, gr.x, gr.y, gr.z, k0);
        hb.y = blur_row(lw, gr.x, gr.y, gr.z, gr.w, k0);
        hb.z = blur_row(gr.x, gr.y, gr.z, gr.w, rx, k0);
        hb.w = blur_row(gr.y, gr.z, gr.w, rx, ry, k0);
        hbw[u % 5] = hb;
        // GaussianBlur columns: level-0 row q = t-2 from rows t-4..t
        const float4 &h0 = hbw[(u + 1) % 5], &h1 = hbw[(u + 2) % 5], &h2 = hbw[(u + 3) % 5], &h3 = hbw[(u + 4) % 5], &h4 = hbw[u % 5];
        const float2 rxy = blur_col2(make_float2(h0.x, h0.y), make_float2(h1.x, h1.y), make_float2(h2.x, h2.y),
                                     make_float2(h3.x, h3.y), make_float2(h4.x, h4.y), k0);
        const float2 rzw = blur_col2(make_float2(h0.z, h0.w), make_float2(h1.z, h1.w), make_float2(h2.z, h2.w),
                                     make_float2(h3.z, h3.w), make_float2(h4.z, h4.w), k0);
        row = make_float4(rxy.x, rxy.y, rzw.x, rzw.y);
        if (useful && q >= r0 && q < q_hi) *reinterpret_cast<float4*>(out0_px + (size_t)q * a.out0_pitch) = row;
      } else {
        row = *reinterpret_cast<const float4*>(&ring[u % RING][lane * LANE_BYTES]);
      }

      // pyrDown rows: pyrDown columns c0 = g/2 and c0+1 need input columns g-2..g+4
      {
        float lz = __shfl_up_sync(SFE_FULL, row.z, 1), lw = __shfl_up_sync(SFE_FULL, row.w, 1);
        float rx = __shfl_down_sync(SFE_FULL, row.x, 1);
        if (ledge) { lz = row.z; lw = row.y; }
        if (redge) rx = row.z;
        float2 ph = make_float2(pd_h(lz, lw, row.x, row.y, row.z), pd_h(row.x, row.y, row.z, row.w, rx));
        if (warp_tail) {
          if (tail_x) ph.x = pd_h_sc(lz, lw, row.x, row.y, row.z);
          if (tail_y) ph.y = pd_h_sc(row.x, row.y, row.z, row.w, rx);
        }
        phw[u % 5] = ph;
      }
      if (u % 2 == 0) {
        // pyrDown columns: row i = (q-2)/2 from input rows q-4..q
        const float2 &p0 = phw[(u + 1) % 5], &p1 = phw[(u + 2) % 5], &p2 = phw[(u + 3) % 5], &p3 = phw[(u + 4) % 5], &p4 = phw[u % 5];
        const float2 pab = pd_v2(p0, p1, p2, p3, p4);
        const float pa = pab.x, pb = pab.y;
        const int i = (q - 2) >> 1;
        if constexpr (DOWN_ONLY) {
          if (useful && i >= j_lo && i < j_hi)
            *reinterpret_cast<float2*>(out1_px + (size_t)i * a.out1_pitch) = make_float2(pa * a.scale, pb * a.scale);
          continue;
        }
        // GaussianBlur rows on the pyrDown row: columns c0-2, c0-1 from the left lane, c0+2, c0+3 from the right lane
        float la = __shfl_up_sync(SFE_FULL, pa, 1), lb = __shfl_up_sync(SFE_FULL, pb, 1);
        float ra = __shfl_down_sync(SFE_FULL, pa, 1), rb = __shfl_down_sync(SFE_FULL, pb, 1);
        if (ledge) { la = ra; lb = pb; }   // columns -2, -1 mirror to 2, 1
        if (redge) { ra = pa; rb = lb; }   // columns w1, w1+1 mirror to w1-2, w1-3
        float2 bh = make_float2(blur_row(la, lb, pa, pb, ra, k1), blur_row(lb, pa, pb, ra, rb, k1));
        const int v = (u / 2) % 5;
        // below the image the pyrDown rows mirror about row h1-1: row h1 is row h1-2, row h1+1 is row h1-3
        if (i == a.h1) bh = bhw[(v + 3) % 5];
        if (i == a.h1 + 1) bh = bhw[(v + 1) % 5];
        bhw[v] = bh;
        const float2 &b0 = bhw[(v + 1) % 5], &b1 = bhw[(v + 2) % 5], &b2 = bhw[(v + 3) % 5], &b3 = bhw[(v + 4) % 5], &b4 = bhw[v];
        const int j = i - 2;
        float2 o = blur_col2(b0, b1, b2, b3, b4, k1);
        if (warp_ctail) {
          if (ctail) o = make_float2(blur_sc(b0.x, b1.x, b2.x, b3.x, b4.x, k1), blur_sc(b0.y, b1.y, b2.y, b3.y, b4.y, k1));
        }
        if (useful && j >= j_lo && j < j_hi) *reinterpret_cast<float2*>(out1_px + (size_t)j * a.out1_pitch) = o;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------
// Row-CTA variant of the fused level-0 + level-1 kernel, for widths of NW * 128 columns (SEGS = 1) or for wider rows cut
// into SEGS column segments with halo lanes (1920 columns = 4 segments of a 4-warp CTA).
//
// The strip kernel above spends 4 of its 32 lanes per warp on halo columns and (at 640 columns) rounds 5.7 strips
// up to 6: 768 lane-columns work for 640 image columns.  Here a CTA owns the FULL width of a row band -- warp k
// the columns [128k, 128k+128), no halo lanes -- and the neighbour pixels of the three horizontal filters travel
// through row buffers in shared memory instead of warp shuffles, which cannot cross a warp boundary.  To pay one
// barrier per tick instead of one per horizontal stage, every horizontal stage consumes the row its producer
// published ONE TICK EARLIER (the own-lane values simply stay in registers for a tick): a tick is
//     barrier; all stages, each reading last tick's row buffers and writing this tick's (double-buffered).
// The lag also makes the stages of one tick independent of each other -- instruction-level parallelism the strip
// kernel's dependent chain does not have.  Image borders (BORDER_REFLECT_101) cost no selects: the row buffers
// have one extra slot on each side, which the first / last lane of the row fill with their own mirrored pixels.
// Per-pixel arithmetic and operation order are unchanged (pyr_math.cuh): bit-identical to the strip kernel.

// gray value of one pixel from the doubled-coefficient dot product (the gray byte sits in bits 16..23):
// 0x4b0000XX is the float 8388608 + XX, and fma(8388608 + v, c, -8388608 c) rounds the exact product v*c once,
// i.e. it equals (float)v * c bit for bit (hessian.h:100-101) without the shift and the integer conversion.
__device__ __forceinline__ float gray_from_dot(unsigned s2) {
  const float c = (float)(1. / 255.);
  return fmaf(__uint_as_float(__byte_perm(s2, 0x4b000000u, 0x7652)), c, -8388608.f * c);
}

// the four pixels of 12 BGR bytes; the 16-bit coefficient pairs are placed so that no byte has to be shifted
__device__ __forceinline__ float4 gray4(uint32_t ra, uint32_t rb, uint32_t rc) {
  constexpr unsigned C0 = 2 * 9798u, C1 = 2 * 19235u, C2 = 2 * 3735u, RND = 1u << 15;
  float4 g;
  g.x = gray_from_dot(__dp2a_hi(C2, ra, __dp2a_lo(C0 | (C1 << 16), ra, RND)));          // ra.b0 ra.b1 ra.b2
  g.y = gray_from_dot(__dp2a_lo(C1 | (C2 << 16), rb, __dp2a_hi(C0 << 16, ra, RND)));    // ra.b3 rb.b0 rb.b1
  g.z = gray_from_dot(__dp2a_lo(C2, rc, __dp2a_hi(C0 | (C1 << 16), rb, RND)));          // rb.b2 rb.b3 rc.b0
  g.w = gray_from_dot(__dp2a_hi(C1 | (C2 << 16), rc, __dp2a_lo(C0 << 16, rc, RND)));    // rc.b1 rc.b2 rc.b3
  return g;
}

// Level 0 of the klt.h flavour (klt.h:98-106): the gray plane as it is and its Scharr/32 gradient pair, streamed like the
// strips above -- one warp per (frame, row band, 128-column strip), every lane four columns, walking down the rows with the
// last three rows of the two separable row passes in registers.  The 3x3 filters reach one column, so one halo lane on
// each side (120 useful columns per strip) and one row above / below the band suffice.  Arithmetic and operation order are
// pyr_l0_kernel<SFE_KLT>'s (pyramid.cu), i.e. the oracle's orc_scharr for even widths.
constexpr int KLT_USEFUL = 120;
__global__ void __launch_bounds__(32) klt_l0_stream_kernel(const StreamArgs a) {
  const int lane = threadIdx.x;
  const int unit = blockIdx.x;
  const int strip = unit % a.strips;
  const int band = (unit / a.strips) % a.bands;
  const int frame = unit / (a.strips * a.bands);
  const int g = KLT_USEFUL * strip - 4 + 4 * lane;
  const bool inimg = g >= 0 && g < a.w;
  const bool ledge = g == 0, redge = g == a.w - 4;
  const bool useful = inimg && lane >= 1 && lane <= 30;
  const int r0 = band * a.band_rows, r1 = min(r0 + a.band_rows, a.h);
  const float k3 = 3.f / 32.f, k10 = 10.f / 32.f;

  const uint8_t* bgr_px = a.bgr + (size_t)frame * a.frame_stride + 3 * (size_t)max(g, 0);
  const long long plane_off = (long long)(a.first + frame) * a.out0_fs + max(g, 0);
  float *o_img = a.out0 + plane_off, *o_gx = a.out_gx + plane_off, *o_gy = a.out_gy + plane_off;

  __shared__ __align__(16) unsigned char ring[RING][32 * 12];
  const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(&ring[0][lane * 12]);
  const int src_bytes = inimg ? 4 : 0;
  // tick t: BGR row t (reflected outside the image) arrives; output row y = t - 1 leaves
  const int t_begin = r0 - 1, t_last = r1;
#pragma unroll
  for (int u = 0; u < PREFETCH; ++u) issue_row<true>(ring_addr + u * 32 * 12, a, bgr_px, nullptr, t_begin + u, src_bytes);

  float4 dxw[3], smw[3];   // rows t-2..t of the horizontal difference p - m and of the horizontal smoothing 3,10,3
  float4 gprev = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int u = 0; u < 3; ++u) dxw[u] = smw[u] = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
  for (int tb = t_begin; tb <= t_last; tb += 30) {
#pragma unroll
    for (int u = 0; u < 30; ++u) {
      const int t = tb + u;
      asm volatile("cp.async.wait_group %0;" ::"n"(PREFETCH - 1) : "memory");
      issue_row<true>(ring_addr + ((u + PREFETCH) % RING) * 32 * 12, a, bgr_px, nullptr, t + PREFETCH, src_bytes);
      const uint32_t* rr = reinterpret_cast<const uint32_t*>(&ring[u % RING][lane * 12]);
      const float4 gr = gray4(rr[0], rr[1], rr[2]);
      float m = __shfl_up_sync(SFE_FULL, gr.w, 1), p = __shfl_down_sync(SFE_FULL, gr.x, 1);
      if (ledge) m = gr.y;   // column -1 mirrors to 1
      if (redge) p = gr.z;   // column w mirrors to w-2
      dxw[u % 3] = make_float4(gr.y - m, gr.z - gr.x, gr.w - gr.y, p - gr.z);
      smw[u % 3] = make_float4(fmaf(k10, gr.x, (m + gr.y) * k3), fmaf(k10, gr.y, (gr.x + gr.z) * k3),
                               fmaf(k10, gr.z, (gr.y + gr.w) * k3), fmaf(k10, gr.w, (gr.z + p) * k3));
      const float4 &d0 = dxw[(u + 1) % 3], &d1 = dxw[(u + 2) % 3], &d2 = dxw[u % 3];
      const float4 &s0 = smw[(u + 1) % 3], &s2 = smw[u % 3];
      const int y = t - 1;
      if (useful && y >= r0 && y < r1) {
        const size_t o = (size_t)y * a.out0_pitch;
        *reinterpret_cast<float4*>(o_img + o) = gprev;
        *reinterpret_cast<float4*>(o_gx + o) = make_float4(fmaf(k3, d0.x + d2.x, d1.x * k10), fmaf(k3, d0.y + d2.y, d1.y * k10),
                                                           fmaf(k3, d0.z + d2.z, d1.z * k10), fmaf(k3, d0.w + d2.w, d1.w * k10));
        *reinterpret_cast<float4*>(o_gy + o) = make_float4(s2.x - s0.x, s2.y - s0.y, s2.z - s0.z, s2.w - s0.w);
      }
      gprev = gr;
    }
  }
}

// Shared-memory accesses of the row buffers as PTX: addresses are 32-bit shared-window offsets plus immediates,
// and the border stores stay single predicated instructions (as C++ `if (lane == 0) ...` they become divergent
// branches, which cost more than the selects they replace).
__device__ __forceinline__ void sts2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void sts1_if(int p, uint32_t addr, float a) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.shared.f32 [%1], %2;\n\t}"
               :: "r"(p), "r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float2 lds2(uint32_t addr) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ float lds1(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
  return r;
}

// Row buffers of one CTA (NL = 32 NW lanes, lane L holds columns 4L..4L+3 resp. level-1 columns 2L, 2L+1).
// A row is split into an xy plane and a zw plane so that every access is a conflict-free 8-byte stride:
//   xy[L] = columns 4L, 4L+1 (L = 0..NL; slot NL = the two columns right of the image)
//   zw[L+1] = columns 4L+2, 4L+3 (slot 0 = the two columns left of the image)
template <int NW>
struct RowBufs {
  static constexpr int NL = 32 * NW;
  float2 gxy[2][NL + 1], gzw[2][NL + 1];   // gray rows (double-buffered: written in tick t, read in tick t+1)
  float2 lxy[2][NL + 1], lzw[2][NL + 1];   // level-0 rows
  float2 pd[NL + 2];                       // pyrDown row: pd[L+1]; written in even ticks, read in odd ticks
};

template <int NW, int BLUR0, int BLUR1, bool BULK = false, int SEGS = 1>
#ifndef ROW_MINB5
#define ROW_MINB5 4  // measured: 4 CTAs x 5 warps at 94 registers beat 5 CTAs at 72 (profiles/README.md)
#endif
__global__ void __launch_bounds__(32 * NW, NW <= 5 ? ROW_MINB5 : (NW <= 10 ? 2 : 1)) pyr_row_kernel(const StreamArgs a) {
  using Bufs = RowBufs<NW>;
  extern __shared__ __align__(16) unsigned char row_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, L = threadIdx.x;
  // SEGS > 1 (frames wider than one CTA can hold at a useful residency, e.g. 1920 columns as two 8-warp CTAs): the row is
  // cut into SEGS column segments of w / SEGS useful columns; a CTA's 128 NW lane-columns start at x0 (a multiple of 16, so
  // a segment's BGR bytes stay 16-byte aligned) and the lanes outside [lane_lo, lane_hi) are halo -- they recompute the
  // neighbour's columns (>= 8 are needed for blur o pyrDown o blur) and store nothing.  The row-buffer slots beyond a cut
  // edge are never written; what is read from them only reaches halo lanes.
  const int seg = SEGS > 1 ? blockIdx.x % SEGS : 0;
  const int unit = SEGS > 1 ? blockIdx.x / SEGS : blockIdx.x;
  const int band = unit % a.bands;
  const int frame = unit / a.bands;
  const int seg_w = a.w / SEGS;
  const int x0 = SEGS == 1 || seg == 0 ? 0 : (seg == SEGS - 1 ? a.w - 128 * NW : (seg * seg_w - (128 * NW - seg_w) / 2) & ~15);
  const int lane_lo = (seg * seg_w - x0) >> 2;
  const bool useful = SEGS == 1 || (L >= lane_lo && L < lane_lo + (seg_w >> 2));
  const bool edge_l = SEGS == 1 || seg == 0, edge_r = SEGS == 1 || seg == SEGS - 1;
  // BULK: a CTA owns whole rows, and a row's BGR bytes are one contiguous, 16-byte aligned run of 384 NW bytes -- one
  // cp.async.bulk transaction through the TMA unit per row (issued by thread 0, completed on the slot's mbarrier) instead
  // of three 4-byte cp.async per lane; the ring is then [slot][row bytes] and lane L reads bytes 12 L ...
  unsigned char* ring = BULK ? row_smem + sizeof(Bufs) : row_smem + sizeof(Bufs) + (size_t)warp * RING * 32 * 12;
  __shared__ __align__(8) unsigned long long ring_bar[RING];
  constexpr uint32_t ROW_BYTES = 384 * NW;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(ring_bar);
  if (BULK && threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < RING; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (BULK) __syncthreads();

  const int g = x0 + 4 * L;
  // OpenCV's scalar-path columns of the horizontal pyrDown pass (pyr_math.cuh): column 0 and the last few, i.e. lanes
  // of the first and the last warp only (w1 = 64 NW is a multiple of 8: the level-1 column blur has no tail)
  const bool tail_x = pd_h_tail(g >> 1, a.hbody), tail_y = pd_h_tail((g >> 1) + 1, a.hbody);
  const bool warp_tail = __any_sync(SFE_FULL, tail_x || tail_y);
  const int r0 = band * a.band_rows;
  const int r1 = min(r0 + a.band_rows, 2 * a.h1);
  const int q_hi = min(r1, a.h), j_lo = r0 >> 1, j_hi = min(r1 >> 1, a.h1);
  const Taps k0 = taps_for(BLUR0), k1 = taps_for(BLUR1);

  const uint8_t* bgr_px = a.bgr + (size_t)frame * a.frame_stride + 3 * (size_t)g;
  float* out0_px = a.out0 + (long long)(a.first + frame) * a.out0_fs + g;
  float* out1_px = a.out1 + (long long)(a.first + frame) * a.out1_fs + (g >> 1);

  // every buffer access of this lane is `lb` plus a compile-time offset
  const uint32_t lb = (uint32_t)__cvta_generic_to_shared(row_smem) + 8 * L;
  constexpr uint32_t GXY = offsetof(Bufs, gxy), GZW = offsetof(Bufs, gzw), LXY = offsetof(Bufs, lxy), LZW = offsetof(Bufs, lzw),
                     PD = offsetof(Bufs, pd), BUF = 8 * (Bufs::NL + 1);
  // image borders: the first lane of the row also fills zw slot 0 (columns -2, -1 = 2, 1), the last lane xy slot NL
  // (columns w, w+1 = w-2, w-3); both are (z, y) of the lane.  `mb` is that slot relative to GXY + buffer offset.
  const int mirror = (L == 0 && edge_l) || (L == Bufs::NL - 1 && edge_r);
  const uint32_t mb = L == 0 ? lb + (GZW - GXY) : lb + 8;
  // pyrDown row: slot 0 = columns (2, 1) = (pa of lane 1, pb of lane 0); slot NL+1 = columns (w1-2, w1-3) =
  // (pa of the last lane, pb of the one before): odd lanes store their pa into .x, even lanes their pb into .y
  const int pd_mirror = (L < 2 && edge_l) || (L >= Bufs::NL - 2 && edge_r);
  const uint32_t pmb = (L < 2 ? lb - 8 * L : lb + 8 * (Bufs::NL + 1 - L)) + PD + ((L & 1) ? 0 : 4);

  // Tick t: BGR row t arrives and becomes gray(t); the lagged stages then work on
  //   gray(t-1) -> horizontally blurred row t-1 -> level-0 row q = t-3
  //   level-0 row qd = t-4 -> horizontally pyrDown-filtered row qd -> (qd even) pyrDown row i = (qd-2)/2
  //   (qd odd) the pyrDown row i = (qd-3)/2 of the previous tick -> blurred rows -> level-1 row j = i-2.
  const int t_begin = r0 - 8;   // even: qd is even exactly when the unrolled tick index is
  const int t_last = r1 + 9;
  const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(ring + lane * 12);
  const uint32_t ring_base = (uint32_t)__cvta_generic_to_shared(ring);
  const uint8_t* frame_px = a.bgr + (size_t)frame * a.frame_stride;
  auto bulk_row = [&](int slot, int t) {   // thread 0 only
    const uint8_t* src = frame_px + (size_t)reflect_row(t, a.h) * a.row_stride + 3 * (size_t)x0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * slot), "r"(ROW_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ring_base + slot * ROW_BYTES), "l"(src), "r"(ROW_BYTES), "r"(bar0 + 8 * slot) : "memory");
  };
  if (BULK) {
    if (threadIdx.x == 0)
      for (int u = 0; u < PREFETCH; ++u) bulk_row(u, t_begin + u);
  } else {
#pragma unroll
    for (int u = 0; u < PREFETCH; ++u) issue_row<true>(ring_addr + u * 32 * 12, a, bgr_px, nullptr, t_begin + u, 4);
  }
  uint32_t round = 0;   // how often the ring has wrapped: the parity the slot barriers are waited with
  const int t_exec_last = t_begin + 10 * ((t_last - t_begin) / 10 + 1) - 1;   // the loop runs whole blocks of ten ticks

  float4 hbw[5];   // horizontally blurred gray rows t-5..t-1
  float2 phw[5];   // horizontally pyrDown-filtered rows qd-4..qd
  float2 bhw[5];   // horizontally blurred pyrDown rows i-4..i
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    hbw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    phw[u] = bhw[u] = make_float2(0.f, 0.f);
  }
  float4 gp = make_float4(0.f, 0.f, 0.f, 0.f);     // gray(t-1)
  float4 rowp = make_float4(0.f, 0.f, 0.f, 0.f);   // level-0 row t-4
  float2 pdp = make_float2(0.f, 0.f);              // pyrDown row formed in the previous (even) tick

#pragma unroll 1
  for (int tb = t_begin; tb <= t_last; tb += 10) {
#pragma unroll
    for (int u = 0; u < 10; ++u) {
      const int t = tb + u;
      const uint32_t cur = (u & 1) * BUF, prev = ((u + 1) & 1) * BUF;
      __syncthreads();
      const uint32_t* rr;
      if (BULK) {
        if (threadIdx.x == 0 && t + PREFETCH <= t_exec_last) bulk_row((u + PREFETCH) % RING, t + PREFETCH);   // the slot read two ticks ago
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(bar0 + 8 * (u % RING)), "r"(round & 1u) : "memory");
        rr = reinterpret_cast<const uint32_t*>(ring + (u % RING) * ROW_BYTES + 12 * L);
      } else {
        asm volatile("cp.async.wait_group %0;" ::"n"(PREFETCH - 1) : "memory");
        issue_row<true>(ring_addr + ((u + PREFETCH) % RING) * 32 * 12, a, bgr_px, nullptr, t + PREFETCH, 4);
        rr = reinterpret_cast<const uint32_t*>(ring + ((u % RING) * 32 + lane) * 12);
      }

      // ---- gray(t), published for the next tick
      const float4 gn = gray4(rr[0], rr[1], rr[2]);
      sts2(lb + GXY + cur, gn.x, gn.y);
      sts2(lb + GZW + cur + 8, gn.z, gn.w);
      sts1_if(mirror, mb + GXY + cur, gn.z);
      sts1_if(mirror, mb + GXY + cur + 4, gn.y);

      // ---- GaussianBlur rows on gray(t-1): columns g-2, g-1 and g+4, g+5 from last tick's buffer
      {
        const float2 l = lds2(lb + GZW + prev), r = lds2(lb + GXY + prev + 8);
        float4 hb;
        hb.x = blur_row(l.x, l.y, gp.x, gp.y, gp.z, k0);
        hb.y = blur_row(l.y, gp.x, gp.y, gp.z, gp.w, k0);
        hb.z = blur_row(gp.x, gp.y, gp.z, gp.w, r.x, k0);
        hb.w = blur_row(gp.y, gp.z, gp.w, r.x, r.y, k0);
        hbw[(u + 4) % 5] = hb;
      }
      // ---- GaussianBlur columns: level-0 row q = t-3 from rows t-5..t-1; stored, and published for the next tick
      const int q = t - 3;
      const float4 &h0 = hbw[u % 5], &h1 = hbw[(u + 1) % 5], &h2 = hbw[(u + 2) % 5], &h3 = hbw[(u + 3) % 5], &h4 = hbw[(u + 4) % 5];
      const float2 rxy = blur_col2(make_float2(h0.x, h0.y), make_float2(h1.x, h1.y), make_float2(h2.x, h2.y),
                                   make_float2(h3.x, h3.y), make_float2(h4.x, h4.y), k0);
      const float2 rzw = blur_col2(make_float2(h0.z, h0.w), make_float2(h1.z, h1.w), make_float2(h2.z, h2.w),
                                   make_float2(h3.z, h3.w), make_float2(h4.z, h4.w), k0);
      const float4 row = make_float4(rxy.x, rxy.y, rzw.x, rzw.y);
      if (q >= r0 && q < q_hi && useful) *reinterpret_cast<float4*>(out0_px + (size_t)q * a.out0_pitch) = row;
      sts2(lb + LXY + cur, row.x, row.y);
      sts2(lb + LZW + cur + 8, row.z, row.w);
      sts1_if(mirror, mb + LXY + cur, row.z);
      sts1_if(mirror, mb + LXY + cur + 4, row.y);

      // ---- pyrDown rows on level-0 row qd = t-4: columns g-2, g-1 and g+4 from last tick's buffer
      {
        const float2 l = lds2(lb + LZW + prev);
        const float r = lds1(lb + LXY + prev + 8);
        float2 ph = make_float2(pd_h(l.x, l.y, rowp.x, rowp.y, rowp.z), pd_h(rowp.x, rowp.y, rowp.z, rowp.w, r));
        if (warp_tail) {
          if (tail_x) ph.x = pd_h_sc(l.x, l.y, rowp.x, rowp.y, rowp.z);
          if (tail_y) ph.y = pd_h_sc(rowp.x, rowp.y, rowp.z, rowp.w, r);
        }
        phw[u % 5] = ph;
      }
      if (u % 2 == 0) {
        // pyrDown columns: row i = (qd-2)/2 from level-0 rows qd-4..qd, published for the next tick
        const float2 &p0 = phw[(u + 1) % 5], &p1 = phw[(u + 2) % 5], &p2 = phw[(u + 3) % 5], &p3 = phw[(u + 4) % 5], &p4 = phw[u % 5];
        pdp = pd_v2(p0, p1, p2, p3, p4);
        sts2(lb + PD + 8, pdp.x, pdp.y);
        sts1_if(pd_mirror, pmb, (L & 1) ? pdp.x : pdp.y);
      } else {
        // GaussianBlur rows on the pyrDown row of the previous tick: columns c0-2, c0-1 and c0+2, c0+3
        const int i = (t - 7) >> 1;
        const float2 l = lds2(lb + PD), r = lds2(lb + PD + 16);
        float2 bh = make_float2(blur_row(l.x, l.y, pdp.x, pdp.y, r.x, k1), blur_row(l.y, pdp.x, pdp.y, r.x, r.y, k1));
        const int v = (u / 2) % 5;
        // below the image the pyrDown rows mirror about row h1-1: row h1 is row h1-2, row h1+1 is row h1-3
        if (i == a.h1) bh = bhw[(v + 3) % 5];
        if (i == a.h1 + 1) bh = bhw[(v + 1) % 5];
        bhw[v] = bh;
        const float2 &b0 = bhw[(v + 1) % 5], &b1 = bhw[(v + 2) % 5], &b2 = bhw[(v + 3) % 5], &b3 = bhw[(v + 4) % 5], &b4 = bhw[v];
        const int j = i - 2;
        if (j >= j_lo && j < j_hi && useful)
          *reinterpret_cast<float2*>(out1_px + (size_t)j * a.out1_pitch) = blur_col2(b0, b1, b2, b3, b4, k1);
      }
      gp = gn;
      rowp = row;
    }
    ++round;
  }
}

int g_slots[2] = {0, 0};  // resident warps of the two instantiations on this device (one warp per CTA)

template <bool FROM_BGR>
int stream_slots() {
  int& s = g_slots[FROM_BGR ? 0 : 1];
  if (s == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pyr_stream_kernel<FROM_BGR, 0, 1>, 32, 0);
    s = sms * (per_sm > 0 ? per_sm : 1);
  }
  return s;
}

// Row bands.  Every band repeats ~15 rows of pipeline fill, so long bands are cheaper per pixel, but the
// kernel needs enough warps to keep the issue slots busy: measured on B200 (profiles/README.md) about 1.3x
// the resident warp slots is the sweet spot (256 VGA frames: 3 bands of 160 rows beat 2 and 4).
template <bool FROM_BGR>
void plan_bands(StreamArgs& a, int count) {
  const int slots = stream_slots<FROM_BGR>();
  const int rows = 2 * a.h1;
  int bands = (int)(1.3 * slots / ((double)count * a.strips) + 0.5);
  static const int forced = getenv("SFE_PYR_BANDS") ? atoi(getenv("SFE_PYR_BANDS")) : 0;  // experiments
  if (forced > 0) bands = forced;
  static const int min_band = getenv("SFE_PYR_MIN_BAND") ? atoi(getenv("SFE_PYR_MIN_BAND")) : 40;  // experiments
  const int max_bands = rows / min_band > 0 ? rows / min_band : 1;
  if (bands < 1) bands = 1;
  if (bands > max_bands) bands = max_bands;
  int br = (rows + bands - 1) / bands;
  br += br & 1;
  a.band_rows = br;
  a.bands = (rows + br - 1) / br;
  a.nunits = count * a.strips * a.bands;
}

template <int NW>
size_t row_smem_bytes() { return sizeof(RowBufs<NW>) + (size_t)NW * RING * 32 * 12; }

// Bands of the row-CTA kernel: a CTA is NW warps, a band repeats 18 rows of pipeline fill (which store nothing: about 7
// rows' worth of time in this write-bound kernel).  All CTAs take the same time, so what matters is how the grid
// quantises into waves.  Fitted to measurements on B200 (128..1024 VGA frames, 64..128 frames of 1920 x 1080, 1..12
// bands, profiles/README.md), in units of one CTA time: a grid below one wave costs 0.45 + 0.55 of its fill (fewer CTAs
// per SM hide less latency), full waves cost 1 each, a partial last wave its fraction but at least 0.5 (its CTAs have
// the SMs almost to themselves).  The band count minimises that cost times (band rows + 7): the grids that win hold
// about 1.73 or 2.77 waves, the ones just above a whole number of waves lose up to 15 %.
static int pick_bands(int ctas_per_band, int slots, int rows, int max_bands) {
  int best = 1;
  double best_cost = 1e30;
  for (int b = 1; b <= max_bands; ++b) {
    int br = (rows + b - 1) / b;
    br += br & 1;
    const int nb = (rows + br - 1) / br;
    const double w = (double)ctas_per_band * nb / slots;
    const double full = (double)(long long)w, frac = w - full;
    const double waves = w < 1. ? 0.45 + 0.55 * w : full + (frac > 0. ? (frac > 0.5 ? frac : 0.5) : 0.);
    const double cost = waves * (br + 7);
    if (cost < best_cost * 0.999) best_cost = cost, best = b;
  }
  return best;
}

template <int NW, bool BULK, int SEGS>
void launch_row_impl(StreamArgs& a, int count, cudaStream_t s) {
  static int slots = 0;  // resident CTAs on the device (the GPUs of a box are alike)
  static PerDeviceOnce attr{};
  if (first_use_on_device(attr))
    cudaFuncSetAttribute(pyr_row_kernel<NW, 0, 1, BULK, SEGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem_bytes<NW>());
  if (slots == 0) {
    int dev = 0, per_sm = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pyr_row_kernel<NW, 0, 1, BULK, SEGS>, 32 * NW, row_smem_bytes<NW>());
    slots = sms * (per_sm > 0 ? per_sm : 1);
  }
  const int rows = 2 * a.h1;
  static const double fill = getenv("SFE_PYR_ROW_FILL") ? atof(getenv("SFE_PYR_ROW_FILL")) : 0.;  // experiments: ctas = fill * slots
  static const int min_band = getenv("SFE_PYR_ROW_MIN_BAND") ? atoi(getenv("SFE_PYR_ROW_MIN_BAND")) : 8;  // a live frame: many short bands (latency), the cost model keeps long ones for batches
  const int max_bands = rows / min_band > 0 ? rows / min_band : 1;
  int bands = fill > 0. ? (int)(fill * slots / ((double)count * SEGS) + 0.5) : pick_bands(count * SEGS, slots, rows, max_bands);
  static const int forced = getenv("SFE_PYR_BANDS") ? atoi(getenv("SFE_PYR_BANDS")) : 0;
  if (forced > 0) bands = forced;
  if (bands < 1) bands = 1;
  if (bands > max_bands) bands = max_bands;
  int br = (rows + bands - 1) / bands;
  br += br & 1;
  a.band_rows = br;
  a.bands = (rows + br - 1) / br;
  a.strips = SEGS;
  a.nunits = count * a.bands * SEGS;
  pyr_row_kernel<NW, 0, 1, BULK, SEGS><<<a.nunits, 32 * NW, row_smem_bytes<NW>(), s>>>(a);  // sigma 1.1 then 0.8 (hessian.h:102,113)
}
template <int NW, int SEGS = 1>
void launch_row(StreamArgs& a, int count, cudaStream_t s) {
  // rows through cp.async.bulk (the TMA unit) when every row starts 16-byte aligned: measured 0.628 -> 0.600 ms per 1024
  // VGA frames against three 4-byte cp.async per lane (profiles/README.md); SFE_PYR_BULK=0 switches it off (experiments)
  static const bool want_bulk = !(getenv("SFE_PYR_BULK") && atoi(getenv("SFE_PYR_BULK")) == 0);
  const bool aligned = (((uintptr_t)a.bgr | a.row_stride | a.frame_stride) & 15) == 0;
  if (want_bulk && aligned) launch_row_impl<NW, true, SEGS>(a, count, s);
  else launch_row_impl<NW, false, SEGS>(a, count, s);
}

}  // namespace

// Streams levels 0 and 1 of `count` frames (SFE_HESSIAN flavour) and then every deeper level the down
// stage applies to.  Returns the number of levels built (>= 2) and adds the kernels launched to *launches,
// or 0 when the geometry does not qualify (the caller falls back to the tiled kernels for all levels).
int launch_pyr_stream_hessian(const PyrView& v, const uint8_t* bgr, size_t row_stride, size_t frame_stride, int first,
                              int count, cudaStream_t s, int* launches) {
  // w % 8 == 0: cv::pyrDown's vertical pass then has no scalar columns (pyr_math.cuh); other widths take the tiled kernels
  auto ok_level = [](int w, int h) { return w % 8 == 0 && w >= 16 && h >= 16; };
  if (v.depth < 2 || !ok_level(v.w[0], v.h[0]) || ((uintptr_t)bgr & 3) || row_stride % 4 || frame_stride % 4) return 0;
  int built = 0;
  // A strip warp walks its rows serially, so a level with little work per launch -- the deeper levels of ONE live frame --
  // is latency-bound on this kernel (measured, one VGA frame: 11 us per level against 5 us on the tiled kernel, which
  // covers the level with many small CTAs).  Levels >= 2 below this many output pixels per launch are left to the caller's
  // tiled kernels; both kernels are bit-identical to the oracle.
  static const long long tiled_below = getenv("SFE_PYR_TILED_BELOW") ? atoll(getenv("SFE_PYR_TILED_BELOW")) : 160 * 120 * 8;
  for (int l = 1; l < v.depth; ++l) {
    if (!ok_level(v.w[l - 1], v.h[l - 1])) break;
    if (l >= 2 && (long long)count * v.w[l] * v.h[l] < tiled_below) break;
    StreamArgs a{};
    a.w = v.w[l - 1]; a.h = v.h[l - 1]; a.w1 = v.w[l]; a.h1 = v.h[l];
    a.hbody = pd_hbody(a.w);
    a.first = first;
    a.strips = (a.w + STRIP_USEFUL - 1) / STRIP_USEFUL;
    a.out1 = v.base[0][l]; a.out1_fs = v.frame_stride[l]; a.out1_pitch = v.pitch[l];
    if (l == 1) {
      a.bgr = bgr; a.row_stride = row_stride; a.frame_stride = frame_stride;
      a.out0 = v.base[0][0]; a.out0_fs = v.frame_stride[0]; a.out0_pitch = v.pitch[0];
      static const bool strips_only = getenv("SFE_PYR_STRIPS") != nullptr;  // experiments: force the strip kernel
      static const bool no_segs = getenv("SFE_PYR_NOSEG") != nullptr;        // experiments: 1920 columns on the strip kernel
      static const bool segs2 = getenv("SFE_PYR_SEG") && atoi(getenv("SFE_PYR_SEG")) == 2;    // experiments
      // row-CTA kernel for the widths it was tuned for; 1920 columns would be one 15-warp CTA per SM (slower than strips)
      if (!strips_only && a.w == 640) launch_row<5>(a, count, s);
      else if (!strips_only && a.w == 1280) launch_row<10>(a, count, s);
      // 1920 columns: one 15-warp CTA per SM loses to the strips, so the row is cut into column segments with halo lanes --
      // four 4-warp CTAs of 480 useful columns (5 CTAs = 20 warps per SM, as at 640 columns); two 8-warp CTAs are slower
      else if (!strips_only && !no_segs && a.w == 1920 && segs2) launch_row<8, 2>(a, count, s);
      else if (!strips_only && !no_segs && a.w == 1920) launch_row<4, 4>(a, count, s);
      else {
        plan_bands<true>(a, count);
        pyr_stream_kernel<true, 0, 1><<<a.nunits, 32, 0, s>>>(a);  // sigma 1.1 then 0.8 (hessian.h:102,113)
      }
    } else {
      a.in = v.base[0][l - 1]; a.in_fs = v.frame_stride[l - 1]; a.in_pitch = v.pitch[l - 1];
      plan_bands<false>(a, count);
      pyr_stream_kernel<false, 0, 1><<<a.nunits, 32, 0, s>>>(a);
    }
    ++*launches;
    built = l + 1;
  }
  return built;
}

// One down stage of ONE plane by the strip kernel, for the flavours whose levels are not all "pyrDown + sigma-0.8 blur":
// blur_id 1 / 2 = pyrDown then GaussianBlur sigma 0.8 / 0.6 (klt.h:117-118), blur_id < 0 = bare pyrDown times `scale`
// (klt.h:123-124 gradients, brute.h:72-77).  Returns 1 when launched, 0 when the geometry does not qualify (the caller
// then uses the tiled kernel).
int launch_pyr_stream_down(const PyrView& v, int plane, int l, int first, int count, int blur_id, float scale, cudaStream_t s) {
  static const bool tiled_only = getenv("SFE_PYR_TILED") != nullptr;
  if (tiled_only || l < 1 || v.w[l - 1] % 8 != 0 || v.w[l - 1] < 16 || v.h[l - 1] < 16 || blur_id == 0 || blur_id > 2) return 0;
  StreamArgs a{};
  a.w = v.w[l - 1]; a.h = v.h[l - 1]; a.w1 = v.w[l]; a.h1 = v.h[l];
  a.hbody = pd_hbody(a.w);
  a.first = first;
  a.strips = (a.w + STRIP_USEFUL - 1) / STRIP_USEFUL;
  a.in = v.base[plane][l - 1]; a.in_fs = v.frame_stride[l - 1]; a.in_pitch = v.pitch[l - 1];
  a.out1 = v.base[plane][l]; a.out1_fs = v.frame_stride[l]; a.out1_pitch = v.pitch[l];
  a.scale = scale;
  plan_bands<false>(a, count);
  if (blur_id == 1) pyr_stream_kernel<false, 0, 1><<<a.nunits, 32, 0, s>>>(a);
  else if (blur_id == 2) pyr_stream_kernel<false, 0, 2><<<a.nunits, 32, 0, s>>>(a);
  else pyr_stream_kernel<false, 0, 1, true><<<a.nunits, 32, 0, s>>>(a);
  return 1;
}

// Level 0 of the klt.h flavour by the streaming kernel.  Returns 1 when launched, 0 when the geometry does not qualify
// (odd widths take the Scharr row filter's scalar tail: the tiled kernel handles them).
int launch_pyr_stream_klt_l0(const PyrView& v, const uint8_t* bgr, size_t row_stride, size_t frame_stride, int first, int count,
                             cudaStream_t s) {
  static const bool tiled_only = getenv("SFE_PYR_TILED") != nullptr;
  if (tiled_only || v.w[0] % 4 != 0 || v.w[0] < 8 || v.h[0] < 2 || ((uintptr_t)bgr & 3) || row_stride % 4 || frame_stride % 4) return 0;
  StreamArgs a{};
  a.w = v.w[0]; a.h = v.h[0];
  a.first = first;
  a.bgr = bgr; a.row_stride = row_stride; a.frame_stride = frame_stride;
  a.out0 = v.base[0][0]; a.out_gx = v.base[1][0]; a.out_gy = v.base[2][0];
  a.out0_fs = v.frame_stride[0]; a.out0_pitch = v.pitch[0];
  a.strips = (a.w + KLT_USEFUL - 1) / KLT_USEFUL;
  static int slots = 0;
  if (slots == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, klt_l0_stream_kernel, 32, 0);
    slots = sms * (per_sm > 0 ? per_sm : 1);
  }
  // a band repeats two rows only, so short bands are cheap: enough units for about three waves of resident warps
  static const double fill = getenv("SFE_KLT_L0_FILL") ? atof(getenv("SFE_KLT_L0_FILL")) : 3.0;  // experiments
  int bands = (int)(fill * slots / ((double)count * a.strips) + 0.5);
  const int max_bands = a.h / 16 > 0 ? a.h / 16 : 1;
  if (bands < 1) bands = 1;
  if (bands > max_bands) bands = max_bands;
  a.band_rows = (a.h + bands - 1) / bands;
  a.bands = (a.h + a.band_rows - 1) / a.band_rows;
  a.nunits = count * a.strips * a.bands;
  klt_l0_stream_kernel<<<a.nunits, 32, 0, s>>>(a);
  return 1;
}
