// hamming_mma.cu -- the 256-bit Hamming matcher (P4) as an EXACT int8 contraction on the 5th-generation tensor
// cores (tcgen05.mma kind::i8, accumulators in tensor memory), with the top-2 scan as its epilogue.
//
// The Hamming distance of two bit rows is bilinear once the bits are written as signs: with q' = 1 - 2*qbit and
// t' = 1 - 2*tbit,  sum_k q'_k t'_k = 256 - 2 d.  The query tile is unpacked to a_k = -64 q'_k, the train tile to
// b_k = +64 t'_k (signed bytes), so  sum_k a_k b_k = 8192 d - 2^20  -- every partial sum is an integer below 2^21,
// s32 accumulation is exact, there is no rounding anywhere.  One more K block carries the tie rule: a = (1, 0, ...),
// b = (j - 128, 0, ...) with j the train row's index inside its 256-row tile, so the accumulator the tensor core
// hands back is
//        acc[i][j] = 8192 d(i,j) + j - 128 - 2^20
// which orders the tile's columns by (distance, train index) -- exactly the packed key of hamming.cu
// (lowest train index wins ties, cv2.BFMatcher's rule; oracle.c orc_hamming256_top2).  The epilogue therefore is
// nothing but a running two-smallest over raw accumulators (5 integer min/max per TWO columns, no decode, no index
// arithmetic); the two winners of a tile are decoded once per tile into the global keys dist << 22 | index.
//
// Roles (288 threads, one CTA per SM, persistent over (batch, 128-query tile, train split) items):
//   warps 0-7  unpack the next 256-row train tile (bits -> signed bytes, shared memory in the no-swizzle K-major
//              core-matrix layout the MMA descriptors name), then scan the previous tile's accumulators: warp w reads
//              TMEM lanes 32 (w % 4).. with tcgen05.ld -- one query row per thread -- columns 128 (w / 4)..
//   warp 8     waits for a full stage, issues 9 x tcgen05.mma (M 128, N 256, K 32) into one half of the 512 TMEM
//              columns and commits them to the stage's mbarrier.
// The train stages, the accumulators and the query tile are double-buffered, so the tensor core works on tile g
// while the workers scan tile g-1 and unpack tile g+1, across item boundaries as well.
#include "sfe_common.cuh"

namespace {

constexpr int MM_M = 128;            // queries per item (TMEM lanes)
constexpr int MM_N = 256;            // train rows per tile (TMEM columns of one accumulator)
constexpr int MM_KB = 9;             // K blocks of 32 bytes: 8 x 32 descriptor bits + the index block
constexpr int MM_WORKERS = 256;      // threads of warps 0-7
constexpr int MM_THREADS = MM_WORKERS + 32;
constexpr int MM_A_BYTES = MM_M * 32 * MM_KB;   // 36,864
constexpr int MM_B_BYTES = MM_N * 32 * MM_KB;   // 73,728
constexpr int MM_SMEM = 2 * MM_A_BYTES + 2 * MM_B_BYTES + 128;   // + barriers and the TMEM address
constexpr int MM_IDX_BITS = 22;
constexpr uint32_t MM_KEY_NONE = 0xffffffffu;
constexpr int MM_BIAS = (1 << 20) + 128;        // acc + MM_BIAS = 8192 d + j

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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bits -> signed bytes -------------------------------------------------------------------------------------------
// Four bits to four bytes: n * 0x10204080 puts bit b of the 4-bit value n at bit 8b + 7 -- the four shifted copies of a
// 4-bit value do not overlap, so nothing carries (a wider operand would: bit 0 << 14 meets bit 7 << 7); masked, that is
// 0x80 per set bit; the xor maps {0, 0x80} to {-64, +64} or the reverse.
template <uint32_t XORC>
__device__ __forceinline__ uint32_t spread4(uint32_t nibble) {
  return ((nibble * 0x10204080u) & 0x80808080u) ^ XORC;
}
// One 32-bit descriptor word = one K block of one row: 32 signed bytes, stored as the row's two 16-byte chunks.
template <uint32_t XORC>
__device__ __forceinline__ void unpack_word(uint32_t w, uint8_t* blk_row /* block base + row-group + row-in-group */) {
  uint4 c0, c1;
  c0.x = spread4<XORC>(w & 0xfu);         c0.y = spread4<XORC>((w >> 4) & 0xfu);
  c0.z = spread4<XORC>((w >> 8) & 0xfu);  c0.w = spread4<XORC>((w >> 12) & 0xfu);
  c1.x = spread4<XORC>((w >> 16) & 0xfu); c1.y = spread4<XORC>((w >> 20) & 0xfu);
  c1.z = spread4<XORC>((w >> 24) & 0xfu); c1.w = spread4<XORC>(w >> 28);
  *reinterpret_cast<uint4*>(blk_row) = c0;
  *reinterpret_cast<uint4*>(blk_row + 128) = c1;
}
constexpr uint32_t XOR_A = 0xC0C0C0C0u;  // query: bit 0 -> -64, bit 1 -> +64   (a = -64 q')
constexpr uint32_t XOR_B = 0x40404040u;  // train: bit 0 -> +64, bit 1 -> -64   (b = +64 t')

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
  I.q0 = (r % mtiles) * MM_M;
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

// Two smallest of a stream, two columns per step: 5 integer min/max (one of them three-input).
__device__ __forceinline__ void top2_pair(int a, int b, int& k1, int& k2) {
  const int lo = min(a, b), hi = max(a, b);
  const int t = max(k1, lo);
  k2 = min(min(k2, hi), t);
  k1 = min(k1, lo);
}

__global__ void __launch_bounds__(MM_THREADS, 1)
hamming_mma_kernel(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt, int batch, int splits,
                   int per /* train rows per split, a multiple of MM_N */, uint2* __restrict__ keys /* [batch][2*splits][nq] */) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                       // 2 query tiles
  uint8_t* sB = smem + 2 * MM_A_BYTES;      // 2 train stages
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
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // K block 8: a = (1, 0, ..., 0) for every query row, b = (j - 128, 0, ..., 0) for train row j of a tile
  for (int r = tid; r < 2 * MM_M; r += MM_THREADS) {
    uint8_t* p = sA + (r / MM_M) * MM_A_BYTES + 8 * MM_M * 32 + row_off(r % MM_M);
    *reinterpret_cast<uint4*>(p) = make_uint4(1u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(p + 128) = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int r = tid; r < 2 * MM_N; r += MM_THREADS) {
    const int j = r % MM_N;
    uint8_t* p = sB + (r / MM_N) * MM_B_BYTES + 8 * MM_N * 32 + row_off(j);
    *reinterpret_cast<uint4*>(p) = make_uint4((uint32_t)((j - 128) & 0xff), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(p + 128) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mtiles = (nq + MM_M - 1) / MM_M;
  const int nitems = batch * mtiles * splits;

  if (warp == 8) {
    // ===== MMA issuer =====
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
          for (int kb = 0; kb < MM_KB; ++kb)
            mma_i8(tmem_base + s * MM_N, smem_desc(a_addr + kb * MM_M * 32, 128, 256), smem_desc(b_addr + kb * MM_N * 32, 128, 256),
                   kb > 0 ? 1u : 0u);
          mma_commit(bar_done + 8 * s);  // arrives when the nine MMAs have read their operands and written the accumulator
        }
        __syncwarp();
      }
    }
  } else {
    // ===== workers: unpack tile g, scan tile g-1 =====
    const uint4* q4 = reinterpret_cast<const uint4*>(q);
    const uint4* t4 = reinterpret_cast<const uint4*>(t);
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;                       // the query row (TMEM lane) this thread scans
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint32_t g = 0;
    int seq = 0;
    // state of the tile whose accumulators are scanned next (tile g-1)
    int p_valid = 0, p_tbase = 0, p_b = 0, p_q0 = 0, p_split = 0;
    bool p_first = false, p_last = false, have_prev = false;
    uint32_t K1 = MM_KEY_NONE, K2 = MM_KEY_NONE;

    auto scan_prev = [&]() {
      const uint32_t s = (g - 1) & 1;
      mbar_wait(bar_done + 8 * s, ((g - 1) >> 1) & 1);
      tc_fence_after();
      if (p_first) K1 = K2 = MM_KEY_NONE;
      int k1 = INT_MAX, k2 = INT_MAX;
      const uint32_t tcol = tmem_base + lane_addr + s * MM_N + half * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = half * 128 + c * 32;
        if (col0 >= p_valid) break;  // warp-uniform
        int v[32];
        tmem_ld32(tcol + c * 32, v);
        tmem_ld_wait();
        if (col0 + 32 > p_valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = col0 + j < p_valid ? v[j] : INT_MAX;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) top2_pair(v[j], v[j + 1], k1, k2);
      }
      // the tile's two winners -> global keys (distance << 22 | train index)
      if (k1 != INT_MAX) {
        const uint32_t u = (uint32_t)(k1 + MM_BIAS);
        top2_u(((u >> 13) << MM_IDX_BITS) | (uint32_t)(p_tbase + (int)(u & 8191u)), K1, K2);
      }
      if (k2 != INT_MAX) {
        const uint32_t u = (uint32_t)(k2 + MM_BIAS);
        top2_u(((u >> 13) << MM_IDX_BITS) | (uint32_t)(p_tbase + (int)(u & 8191u)), K1, K2);
      }
      if (p_last && p_q0 + row < nq)
        keys[((size_t)p_b * (2 * splits) + 2 * p_split + half) * nq + p_q0 + row] = make_uint2(K1, K2);
      tc_fence_before();  // the accumulator reads are ordered before the arrive that lets the next MMA overwrite them
    };

    for (int it = blockIdx.x; it < nitems; it += gridDim.x, ++seq) {
      const Item I = item_of(it, mtiles, splits, nt, per);
      for (int t0 = I.t0; t0 < I.t1; t0 += MM_N, ++g) {
        const uint32_t s = g & 1;
        if (t0 == I.t0) {
          // query tile of this item: thread -> (row tid % 128, words 4 (tid / 128) ..)
          const int r = tid & (MM_M - 1), hw = tid >> 7;
          const uint4 w = __ldg(q4 + ((size_t)I.b * nq + min(I.q0 + r, nq - 1)) * 2 + hw);
          uint8_t* base = sA + (seq & 1) * MM_A_BYTES + row_off(r);
          unpack_word<XOR_A>(w.x, base + (4 * hw + 0) * MM_M * 32);
          unpack_word<XOR_A>(w.y, base + (4 * hw + 1) * MM_M * 32);
          unpack_word<XOR_A>(w.z, base + (4 * hw + 2) * MM_M * 32);
          unpack_word<XOR_A>(w.w, base + (4 * hw + 3) * MM_M * 32);
        }
        {
          // train tile: thread -> row tid, all 8 words
          const size_t tr = (size_t)I.b * nt + min(t0 + tid, nt - 1);
          const uint4 w0 = __ldg(t4 + tr * 2), w1 = __ldg(t4 + tr * 2 + 1);
          uint8_t* base = sB + s * MM_B_BYTES + row_off(tid);
          unpack_word<XOR_B>(w0.x, base + 0 * MM_N * 32);
          unpack_word<XOR_B>(w0.y, base + 1 * MM_N * 32);
          unpack_word<XOR_B>(w0.z, base + 2 * MM_N * 32);
          unpack_word<XOR_B>(w0.w, base + 3 * MM_N * 32);
          unpack_word<XOR_B>(w1.x, base + 4 * MM_N * 32);
          unpack_word<XOR_B>(w1.y, base + 5 * MM_N * 32);
          unpack_word<XOR_B>(w1.z, base + 6 * MM_N * 32);
          unpack_word<XOR_B>(w1.w, base + 7 * MM_N * 32);
        }
        fence_async_smem();              // generic-proxy stores -> visible to the tensor core's async proxy
        mbar_arrive(bar_full + 8 * s);
        if (have_prev) scan_prev();
        p_valid = min(MM_N, I.t1 - t0);
        p_tbase = t0;
        p_b = I.b; p_q0 = I.q0; p_split = I.split;
        p_first = t0 == I.t0;
        p_last = t0 + MM_N >= I.t1;
        have_prev = true;
      }
    }
    if (have_prev) scan_prev();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

// Launches the tensor-core matcher; writes [batch][2 * splits][nq] keys for hamming_finalize_kernel (hamming.cu), which
// merges them exactly as it merges the ALU kernel's train splits.  Returns the number of key rows per query (2 * splits)
// or a negative cudaError.
int launch_hamming_mma(const uint32_t* q, int nq, const uint32_t* t, int nt, int batch, int num_sms, void** ws, size_t* ws_cap,
                       cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hamming_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM);
    if (e != cudaSuccess) return -(int)e;
    configured = true;
  }
  const int mtiles = (nq + MM_M - 1) / MM_M;
  const int ntiles = (nt + MM_N - 1) / MM_N;
  // split the train set over CTAs when the query tiles alone cannot fill the GPU
  int splits = 1;
  if (batch * mtiles < num_sms) splits = min(ntiles, (num_sms + batch * mtiles - 1) / (batch * mtiles));
  const int per = ((ntiles + splits - 1) / splits) * MM_N;
  splits = (nt + per - 1) / per;
  const size_t need = (size_t)batch * 2 * splits * nq * sizeof(uint2);
  if (need > *ws_cap) {
    if (*ws) cudaFree(*ws);
    cudaError_t e = cudaMalloc(ws, need);
    if (e != cudaSuccess) { *ws = nullptr; *ws_cap = 0; return -(int)e; }
    *ws_cap = need;
  }
  const int nitems = batch * mtiles * splits;
  hamming_mma_kernel<<<min(nitems, num_sms), MM_THREADS, MM_SMEM, s>>>(q, nq, t, nt, batch, splits, per, (uint2*)*ws);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 2 * splits : -(int)e;
}
