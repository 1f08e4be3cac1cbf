// hamming_mma.cu -- the 256-bit Hamming matcher (P4) as an EXACT int8 contraction on the 5th-generation tensor
// cores (tcgen05.mma kind::i8, accumulators in tensor memory), with the top-2 scan as its epilogue.
//
// The Hamming distance of two bit rows is bilinear once the bits are written as signs: with q' = 1 - 2*qbit and
// t' = 1 - 2*tbit,  sum_k q'_k t'_k = 256 - 2 d.  The query tile is unpacked to a_k = -8 q'_k, the train tile to
// b_k = +8 t'_k (signed bytes), so  sum_k a_k b_k = 128 d - 2^14  -- every partial sum is an integer of at most 15 bits,
// s32 accumulation is exact, there is no rounding anywhere.  One more K block carries the tie rule: a = (1, 0, ...),
// b = (j, 0, ...) with j the train row's index inside its 128-row tile, so the accumulator the tensor core
// hands back is
//        acc[i][j] = 128 d(i,j) + j - 16384        in [-16384, 16511]: a signed 16-bit number
// which orders the columns of a tile by (distance, train index) -- exactly the packed key of hamming.cu
// (lowest train index wins ties, cv2.BFMatcher's rule; oracle.c orc_hamming256_top2).  Because the key fits 16 bits,
// tcgen05.ld.pack::16b delivers TWO columns per register and the scan is a running two-smallest in both halves of a
// register at once (VIMNMX.S16x2 / VIMNMX3.S16x2: 5 instructions per FOUR columns, no decode, no index arithmetic); the
// winners of a tile are decoded once per tile into the global keys dist << 22 | index.
//
// Roles (544 threads, one CTA per SM, persistent over (batch, 256-query tile, train split) items):
//   warps 0-15 unpack the next 128-row train tile (bits -> signed bytes, shared memory in the no-swizzle K-major
//              core-matrix layout the MMA descriptors name), then scan the previous tile's accumulators: warp w reads
//              TMEM lanes 32 (w % 4).. of query half (w / 4) % 2 with tcgen05.ld -- one query row per thread --
//              columns 64 (w / 8)..
//   warp 16    waits for a full stage, issues 2 x 9 tcgen05.mma (M 128, N 128, K 32: the item's two 128-query blocks
//              against the same train tile) into two of the four 128-column accumulators and commits them to the stage's
//              mbarrier.  Two query blocks per train tile halve the unpack work per comparison.
// The train stages, the accumulators and the query tile are double-buffered, so the tensor core works on tile g
// while the workers scan tile g-1 and unpack tile g+1, across item boundaries as well.  What bounds it is the ALU pipe
// of the worker warps (profiles/ham_r2*): the nine UTCIMMA per tile hide completely behind unpack + scan.
#include "sfe_common.cuh"
#include "device_once.cuh"

namespace {

constexpr int MM_M = 128;            // rows of one MMA (TMEM lanes)
constexpr int MM_MQ = 256;           // queries per item: two MMA row blocks share every train tile (halves the unpack per comparison)
constexpr int MM_N = 128;            // train rows per tile (TMEM columns of one accumulator)
constexpr int MM_KB = 9;             // K blocks of 32 bytes: 8 x 32 descriptor bits + the index block
constexpr int MM_WWARPS = 16;        // worker warps: TMEM quadrant w & 3, query half (w >> 2) & 1, column half w >> 3
constexpr int MM_WORKERS = 32 * MM_WWARPS;
constexpr int MM_THREADS = MM_WORKERS + 32;      // warp 16 issues the MMAs
constexpr int MM_PARTS = 2;                      // column halves of a tile, scanned by different warps: key rows per split
constexpr int MM_A_BYTES = MM_MQ * 32 * MM_KB;  // 73,728
constexpr int MM_B_BYTES = MM_N * 32 * MM_KB;   // 36,864
constexpr int MM_SMEM = 2 * MM_A_BYTES + 2 * MM_B_BYTES + 128;   // + barriers and the TMEM address
constexpr int MM_IDX_BITS = 22;
constexpr uint32_t MM_KEY_NONE = 0xffffffffu;
constexpr int MM_BIAS = 16384;                  // acc + MM_BIAS = 128 d + (train row inside its tile)

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol error becomes a trap (an error code at the next synchronisation) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle: 8-row x 16-byte core matrices stored as 128 contiguous bytes;
// the two 16-byte K chunks of an MMA (K = 32 bytes) are `lbo` bytes apart, 8-row groups `sbo` bytes apart
// (cute::UMMA::SmemDescriptor, version 1 = sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = signed 8 bit, both K-major, dense.
constexpr uint32_t MM_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MM_N >> 3) << 17) | ((uint32_t)(MM_M >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(MM_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 64 accumulator columns as 32 registers: register r = low 16 bits of column 2r | low 16 bits of column 2r + 1 << 16
__device__ __forceinline__ void tmem_ld64_packed(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bits -> signed bytes -------------------------------------------------------------------------------------------
// Four bits to four bytes: n * 0x00204081 puts bit b of the 4-bit value n at bit 8b -- the four shifted copies of a 4-bit
// value do not overlap, so nothing carries (a wider operand would: bit 0 << 14 meets bit 7 << 7); masked, that is one
// byte of 0 / 1 per bit; a multiply-add maps {0, 1} to {+8, -8} (train rows, b = +8 t') or {-8, +8} (query rows,
// a = -8 q') without a borrow between the bytes.  Two of the five operations are integer multiply-adds, which issue on
// the FMA pipe -- the ALU pipe is what bounds this kernel.
template <bool QUERY>
__device__ __forceinline__ uint32_t spread4(uint32_t nibble) {
  const uint32_t y = (nibble * 0x00204081u) & 0x01010101u;
  return QUERY ? 0xF8F8F8F8u - y * 0xF0u : y * 0xF0u + 0x08080808u;
}
// One 32-bit descriptor word = one K block of one row: 32 signed bytes, stored as the row's two 16-byte chunks.
template <bool QUERY>
__device__ __forceinline__ void unpack_word(uint32_t w, uint8_t* blk_row /* block base + row-group + row-in-group */) {
  uint4 c0, c1;
  c0.x = spread4<QUERY>(w & 0xfu);         c0.y = spread4<QUERY>((w >> 4) & 0xfu);
  c0.z = spread4<QUERY>((w >> 8) & 0xfu);  c0.w = spread4<QUERY>((w >> 12) & 0xfu);
  c1.x = spread4<QUERY>((w >> 16) & 0xfu); c1.y = spread4<QUERY>((w >> 20) & 0xfu);
  c1.z = spread4<QUERY>((w >> 24) & 0xfu); c1.w = spread4<QUERY>(w >> 28);
  *reinterpret_cast<uint4*>(blk_row) = c0;
  *reinterpret_cast<uint4*>(blk_row + 128) = c1;
}
constexpr bool XOR_A = true;   // query rows
constexpr bool XOR_B = false;  // train rows

// Operand tile layout (R rows): block kb at kb * R * 32; inside a block row r sits at (r / 8) * 256 + (r % 8) * 16, its
// second K chunk 128 bytes further: LBO = 128, SBO = 256.
__device__ __forceinline__ uint32_t row_off(int r) { return (uint32_t)((r >> 3) * 256 + (r & 7) * 16); }

struct Item {
  int b, q0, t0, t1, split;
};
__device__ __forceinline__ Item item_of(int it, int mtiles, int splits, int nt, int per) {
  Item I;
  I.split = it % splits;
  const int r = it / splits;
  I.q0 = (r % mtiles) * MM_MQ;
  I.b = r / mtiles;
  I.t0 = I.split * per;
  I.t1 = min(nt, I.t0 + per);
  return I;
}

__device__ __forceinline__ void top2_u(uint32_t key, uint32_t& k1, uint32_t& k2) {
  const uint32_t hi = max(k1, key);
  k1 = min(k1, key);
  k2 = min(k2, hi);
}

// Two smallest of a stream in both 16-bit halves of a register at once, two registers (four columns) per step:
// 4 x VIMNMX.S16x2 + 1 x VIMNMX3.S16x2.
__device__ __forceinline__ void top2_pair_s16x2(uint32_t a, uint32_t b, uint32_t& k1, uint32_t& k2) {
  const uint32_t lo = __vmins2(a, b), hi = __vmaxs2(a, b);
  const uint32_t t = __vmaxs2(k1, lo);
  k2 = __vimin3_s16x2(k2, hi, t);
  k1 = __vmins2(k1, lo);
}

__global__ void __launch_bounds__(MM_THREADS, 1)
hamming_mma_kernel(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt, int batch, int splits,
                   int per /* train rows per split, a multiple of MM_N */, uint2* __restrict__ keys /* [batch][MM_PARTS*splits][nq] */) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                       // 2 query tiles of 256 rows
  uint8_t* sB = smem + 2 * MM_A_BYTES;      // 2 train stages of 128 rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * MM_A_BYTES + 2 * MM_B_BYTES);  // full[2], done[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_done = smem_u32(bars + 2);

  // ---- one-time setup: barriers, TMEM, the constant index blocks ------------------------------------------------------
  if (tid == 0) {
    mbar_init(bar_full, MM_WORKERS);
    mbar_init(bar_full + 8, MM_WORKERS);
    mbar_init(bar_done, 1);
    mbar_init(bar_done + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MM_WWARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // K block 8: a = (1, 0, ..., 0) for every query row, b = (j, 0, ..., 0) for train row j of a tile
  for (int r = tid; r < 2 * MM_MQ; r += MM_THREADS) {
    uint8_t* p = sA + (r / MM_MQ) * MM_A_BYTES + 8 * MM_MQ * 32 + row_off(r % MM_MQ);
    *reinterpret_cast<uint4*>(p) = make_uint4(1u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(p + 128) = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int r = tid; r < 2 * MM_N; r += MM_THREADS) {
    const int j = r % MM_N;
    uint8_t* p = sB + (r / MM_N) * MM_B_BYTES + 8 * MM_N * 32 + row_off(j);
    *reinterpret_cast<uint4*>(p) = make_uint4((uint32_t)j, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(p + 128) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mtiles = (nq + MM_MQ - 1) / MM_MQ;
  const int nitems = batch * mtiles * splits;

  if (warp == MM_WWARPS) {
    // ===== MMA issuer: per tile 2 x 9 MMAs (query rows 0-127 and 128-255) into the stage's two accumulators =====
    uint32_t g = 0;
    int seq = 0;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++seq) {
      const Item I = item_of(it, mtiles, splits, nt, per);
      const uint32_t a_addr = smem_u32(sA + (seq & 1) * MM_A_BYTES);
      for (int t0 = I.t0; t0 < I.t1; t0 += MM_N, ++g) {
        const uint32_t s = g & 1;
        mbar_wait(bar_full + 8 * s, (g >> 1) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b_addr = smem_u32(sB + s * MM_B_BYTES);
#pragma unroll
          for (int mh = 0; mh < 2; ++mh)
#pragma unroll
            for (int kb = 0; kb < MM_KB; ++kb)
              mma_i8(tmem_base + (2 * s + mh) * MM_N, smem_desc(a_addr + kb * MM_MQ * 32 + mh * MM_M * 32, 128, 256),
                     smem_desc(b_addr + kb * MM_N * 32, 128, 256), kb > 0 ? 1u : 0u);
          mma_commit(bar_done + 8 * s);  // arrives when the MMAs have read their operands and written the accumulators
        }
        __syncwarp();
      }
    }
  } else {
    // ===== workers: unpack tile g, scan tile g-1 =====
    const int quad = warp & 3, mh = (warp >> 2) & 1, part = warp >> 3;
    const int row = mh * MM_M + quad * 32 + lane;           // the query row of the item this thread scans
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint32_t g = 0;
    int seq = 0;
    // state of the tile whose accumulators are scanned next (tile g-1)
    int p_valid = 0, p_tbase = 0, p_b = 0, p_q0 = 0, p_split = 0;
    bool p_first = false, p_last = false, p_long = false, have_prev = false;
    uint32_t K1 = MM_KEY_NONE, K2 = MM_KEY_NONE;
    // this thread's words of the next train tile (row tid & 127, words 2 (tid >> 7) ..) and of the next query tile
    // (row tid & 255, words 4 (tid >> 8) ..), requested one tile / one item ahead
    uint2 nw = make_uint2(0u, 0u);
    uint4 qw = make_uint4(0u, 0u, 0u, 0u);
    bool have_next = false, have_q = false;
    auto load_train = [&](int b, int t0) {
      const int r = tid & (MM_N - 1), pr = tid >> 7;
      nw = __ldg(reinterpret_cast<const uint2*>(t + ((size_t)b * nt + min(t0 + r, nt - 1)) * 8) + pr);
      have_next = true;
    };
    auto load_query = [&](int b, int q0) {
      const int r = tid & (MM_MQ - 1), hw = tid >> 8;
      qw = __ldg(reinterpret_cast<const uint4*>(q + ((size_t)b * nq + min(q0 + r, nq - 1)) * 8) + hw);
      have_q = true;
    };

    auto scan_prev = [&]() {
      const uint32_t s = (g - 1) & 1;
      mbar_wait(bar_done + 8 * s, ((g - 1) >> 1) & 1);
      tc_fence_after();
      if (p_first) K1 = K2 = MM_KEY_NONE;
      // this warp's 64 columns in one packed load; 0x7fff = no candidate in that half-register
      uint32_t k1 = 0x7fff7fffu, k2 = 0x7fff7fffu;
      const int col0 = part * 64;
      if (col0 < p_valid) {  // warp-uniform
        uint32_t v[32];
        tmem_ld64_packed(tmem_base + lane_addr + (2 * s + mh) * MM_N + col0, v);
        tmem_ld_wait();
        if (col0 + 64 > p_valid) {  // the last tile of a train range: columns >= p_valid hold stale rows
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int cc = col0 + 2 * r;
            v[r] = cc + 1 < p_valid ? v[r] : (cc < p_valid ? (v[r] | 0x7fff0000u) & 0x7fffffffu : 0x7fff7fffu);
          }
        }
        // Against a long train range almost no tile holds a column that can still enter a query's two best: one pass of
        // three-input minima finds the tile's smallest key, and the scan proper runs only if some lane of the warp needs
        // it (a tie in distance counts as needed, so the lowest-index rule is untouched).  With a few thousand train rows
        // some lane of the 32 almost always does, so there the test would be pure overhead and is left out.
        bool scan = true;
        if (p_long) {
          uint32_t mm = v[0];
#pragma unroll
          for (int r = 1; r + 1 < 32; r += 2) mm = __vimin3_s16x2(mm, v[r], v[r + 1]);
          mm = __vmins2(mm, v[31]);
          const int cmin = min((int)(short)(mm & 0xffffu), (int)mm >> 16);
          const int reach = (int)((K2 >> MM_IDX_BITS) << 7) + 127 - MM_BIAS;  // largest key a column of this thread's second-best distance can have
          scan = __any_sync(0xffffffffu, cmin <= reach);
        }
        if (scan) {
#pragma unroll
          for (int r = 0; r < 32; r += 2) top2_pair_s16x2(v[r], v[r + 1], k1, k2);
        }
      }
      // even and odd columns each kept their own two smallest: the tile's two winners are the two smallest of those four
      {
        const int a1 = (int)(short)(k1 & 0xffffu), b1 = (int)k1 >> 16, a2 = (int)(short)(k2 & 0xffffu), b2 = (int)k2 >> 16;
        const int m1 = min(a1, b1), m2 = min(max(a1, b1), min(a2, b2));
        if (m1 != 0x7fff) {
          const uint32_t u = (uint32_t)(m1 + MM_BIAS);
          top2_u(((u >> 7) << MM_IDX_BITS) | (uint32_t)(p_tbase + (int)(u & 127u)), K1, K2);
        }
        if (m2 != 0x7fff) {
          const uint32_t u = (uint32_t)(m2 + MM_BIAS);
          top2_u(((u >> 7) << MM_IDX_BITS) | (uint32_t)(p_tbase + (int)(u & 127u)), K1, K2);
        }
      }
      if (p_last && p_q0 + row < nq)
        keys[((size_t)p_b * (MM_PARTS * splits) + MM_PARTS * p_split + part) * nq + p_q0 + row] = make_uint2(K1, K2);
      tc_fence_before();  // the accumulator reads are ordered before the arrive that lets the next MMA overwrite them
    };

    for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++seq) {
      const Item I = item_of(it, mtiles, splits, nt, per);
      for (int t0 = I.t0; t0 < I.t1; t0 += MM_N, ++g) {
        const uint32_t s = g & 1;
        if (t0 == I.t0) {
          // query tile of this item: 256 rows x 8 words, four words per thread
          const int r = tid & (MM_MQ - 1), hw = tid >> 8;
          if (!have_q) load_query(I.b, I.q0);
          have_q = false;
          uint8_t* base = sA + (seq & 1) * MM_A_BYTES + row_off(r);
          const uint32_t w[4] = {qw.x, qw.y, qw.z, qw.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) unpack_word<XOR_A>(w[i], base + (4 * hw + i) * MM_MQ * 32);
        }
        {
          // train tile: 128 rows x 8 words, two words per thread (requested one tile ago: the L2 round trip of these
          // loads was the kernel's top stall when they were issued here)
          const int r = tid & (MM_N - 1), pr = tid >> 7;
          if (!have_next) load_train(I.b, t0);
          uint8_t* base = sB + s * MM_B_BYTES + row_off(r);
          unpack_word<XOR_B>(nw.x, base + (2 * pr + 0) * MM_N * 32);
          unpack_word<XOR_B>(nw.y, base + (2 * pr + 1) * MM_N * 32);
        }
        fence_async_smem();              // generic-proxy stores -> visible to the tensor core's async proxy
        mbar_arrive(bar_full + 8 * s);
        // request the next tile's rows (of this item, or the first tile and the queries of this CTA's next item)
        have_next = false;
        if (t0 + MM_N < I.t1) {
          load_train(I.b, t0 + MM_N);
        } else if (it + (int)gridDim.x < nitems) {
          const Item N = item_of(it + gridDim.x, mtiles, splits, nt, per);
          load_train(N.b, N.t0);
          load_query(N.b, N.q0);
        }
        if (have_prev) scan_prev();
        p_valid = min(MM_N, I.t1 - t0);
        p_tbase = t0;
        p_b = I.b; p_q0 = I.q0; p_split = I.split;
        p_first = t0 == I.t0;
        p_last = t0 + MM_N >= I.t1;
        p_long = I.t1 - I.t0 >= 16384;
        have_prev = true;
      }
    }
    if (have_prev) scan_prev();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MM_WWARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

// Launches the tensor-core matcher; writes [batch][MM_PARTS * splits][nq] keys for hamming_finalize_kernel (hamming.cu),
// which merges them exactly as it merges the ALU kernel's train splits.  Returns the number of key rows per query
// (MM_PARTS * splits) or a negative cudaError.
int launch_hamming_mma(const uint32_t* q, int nq, const uint32_t* t, int nt, int batch, int num_sms, void** ws, size_t* ws_cap,
                       cudaStream_t s) {
  static PerDeviceOnce configured{};
  if (first_use_on_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(hamming_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
    if (e != cudaSuccess) return -(int)e;
  }
  const int mtiles = (nq + MM_MQ - 1) / MM_MQ;
  const int ntiles = (nt + MM_N - 1) / MM_N;
  // split the train set over CTAs when the query tiles alone cannot fill the GPU
  int splits = 1;
  if (batch * mtiles < num_sms) splits = min(ntiles, (num_sms + batch * mtiles - 1) / (batch * mtiles));
  const int per = ((ntiles + splits - 1) / splits) * MM_N;
  splits = (nt + per - 1) / per;
  const size_t need = (size_t)batch * MM_PARTS * splits * nq * sizeof(uint2);
  if (need > *ws_cap) {
    if (*ws) cudaFree(*ws);
    cudaError_t e = cudaMalloc(ws, need);
    if (e != cudaSuccess) { *ws = nullptr; *ws_cap = 0; return -(int)e; }
    *ws_cap = need;
  }
  const int nitems = batch * mtiles * splits;
  hamming_mma_kernel<<<min(nitems, num_sms), MM_THREADS, MM_SMEM, s>>>(q, nq, t, nt, batch, splits, per, (uint2*)*ws);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MM_PARTS * splits : -(int)e;
}
