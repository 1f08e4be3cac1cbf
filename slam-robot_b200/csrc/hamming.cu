// hamming.cu -- brute-force 256-bit Hamming matcher with fused top-2 and ratio test (P4).
//
// There is no descriptor matcher in the reference; the behaviour is specified
// against cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2): distance = popcount(xor), best and second
// best train rows per query, LOWEST train index wins ties (oracle.c orc_hamming256_top2).
//
// Integer-ALU work, not a GEMM: each thread keeps QPT query descriptors in registers (8 words
// each), train descriptors stream through shared memory in tiles and are read as warp-wide
// broadcasts (2 x LDS.128 per descriptor), the distance is 8 LOP3 xor + 6 LOP3 carry-save + 5 POPC and the running
// top-2 is three min/max on packed keys  key = dist << 22 | train_index  -- ordering by key IS the
// tie rule, independent of the order tiles are visited, so the train set can be split over CTAs.
// A finalize kernel merges the per-split keys, decodes them and applies the integer ratio test
// d1 * den < d2 * num.
#include <stdlib.h>

#include "sfe_common.cuh"

namespace {

#ifndef HM_CSA
#define HM_CSA 2  // carry-save adders in front of the popcounts (2: 6 POPC, 3: 5 POPC per comparison; measured 0.775 vs 0.80 ms)
#endif
#ifndef HM_UNROLL
#define HM_UNROLL 8
#endif
constexpr int kUnroll = HM_UNROLL;
constexpr int HM_THREADS = 128;
constexpr int HM_QPT = 2;                    // queries per thread
constexpr int HM_QPB = HM_THREADS * HM_QPT;  // queries per CTA
constexpr int HM_TT = 256;                   // train descriptors per shared-memory tile
constexpr uint32_t KEY_NONE = 0xffffffffu;
constexpr int IDX_BITS = 22;                 // nt <= 4M per call

__device__ __forceinline__ void top2_update(uint32_t key, uint32_t& k1, uint32_t& k2) {
  uint32_t hi = max(k1, key);
  k1 = min(k1, key);
  k2 = min(k2, hi);
}

__global__ void __launch_bounds__(HM_THREADS) hamming_kernel(const uint4* __restrict__ q, int nq,
                                                             const uint4* __restrict__ t, int nt, int splits,
                                                             uint2* __restrict__ keys /* [batch][splits][nq] */) {
  __shared__ uint4 tile[HM_TT * 2];
  const int b = blockIdx.z, split = blockIdx.y;
  const uint4* qb = q + (size_t)b * nq * 2;
  const uint4* tb = t + (size_t)b * nt * 2;
  const int per = (nt + splits - 1) / splits;
  const int t0 = split * per, t1 = min(nt, t0 + per);

  uint4 qa[HM_QPT], qc[HM_QPT];
  uint32_t k1[HM_QPT], k2[HM_QPT];
  int qi[HM_QPT];
#pragma unroll
  for (int u = 0; u < HM_QPT; ++u) {
    qi[u] = blockIdx.x * HM_QPB + u * HM_THREADS + threadIdx.x;
    int qq = min(qi[u], nq - 1);
    qa[u] = __ldg(qb + 2 * (size_t)qq);
    qc[u] = __ldg(qb + 2 * (size_t)qq + 1);
    k1[u] = k2[u] = KEY_NONE;
  }

  for (int base = t0; base < t1; base += HM_TT) {
    const int cnt = min(HM_TT, t1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 2; e += HM_THREADS) tile[e] = __ldg(tb + 2 * (size_t)base + e);
    __syncthreads();
#pragma unroll kUnroll
    for (int j = 0; j < cnt; ++j) {
      const uint4 ta = tile[2 * j], tc = tile[2 * j + 1];
      const uint32_t jj = (uint32_t)(base + j);
#pragma unroll
      for (int u = 0; u < HM_QPT; ++u) {
        // POPC issues at a quarter of the LOP3 rate, so carry-save adders (2 LOP3 each) first fold the eight
        // XOR words into words of weight 1 and weight 2: 6 (or 5) POPC instead of 8, both pipes about equally busy
        const uint32_t w0 = qa[u].x ^ ta.x, w1 = qa[u].y ^ ta.y, w2 = qa[u].z ^ ta.z, w3 = qa[u].w ^ ta.w;
        const uint32_t w4 = qc[u].x ^ tc.x, w5 = qc[u].y ^ tc.y, w6 = qc[u].z ^ tc.z, w7 = qc[u].w ^ tc.w;
        const uint32_t s0 = w0 ^ w1 ^ w2, c0 = (w0 & w1) | (w2 & (w0 | w1));
        const uint32_t s1 = w3 ^ w4 ^ w5, c1 = (w3 & w4) | (w5 & (w3 | w4));
#if HM_CSA == 3
        const uint32_t s2 = s0 ^ s1 ^ w6, c2 = (s0 & s1) | (w6 & (s0 | s1));
        const uint32_t d = (__popc(s2) + __popc(w7)) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
#else
        const uint32_t d = (__popc(s0) + __popc(s1) + __popc(w6) + __popc(w7)) + 2 * (__popc(c0) + __popc(c1));
#endif
        top2_update(d * (1u << IDX_BITS) + jj, k1[u], k2[u]);  // == d << 22 | jj; the multiply-add runs on the FMA pipe
      }
    }
  }
#pragma unroll
  for (int u = 0; u < HM_QPT; ++u)
    if (qi[u] < nq) keys[((size_t)b * splits + split) * nq + qi[u]] = make_uint2(k1[u], k2[u]);
}

__global__ void hamming_finalize_kernel(const uint2* __restrict__ keys, int nq, int splits, int total /*batch*nq*/,
                                        int ratio_num, int ratio_den, int max_dist, int32_t* __restrict__ idx,
                                        int32_t* __restrict__ dist, uint8_t* __restrict__ pass) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  int b = g / nq, i = g - b * nq;
  uint32_t k1 = KEY_NONE, k2 = KEY_NONE;
  for (int s = 0; s < splits; ++s) {
    uint2 k = keys[((size_t)b * splits + s) * nq + i];
    top2_update(k.x, k1, k2);
    top2_update(k.y, k1, k2);
  }
  int i1 = -1, i2 = -1, d1 = 257, d2 = 257;
  if (k1 != KEY_NONE) { i1 = (int)(k1 & ((1u << IDX_BITS) - 1)); d1 = (int)(k1 >> IDX_BITS); }
  if (k2 != KEY_NONE) { i2 = (int)(k2 & ((1u << IDX_BITS) - 1)); d2 = (int)(k2 >> IDX_BITS); }
  idx[2 * (size_t)g] = i1; idx[2 * (size_t)g + 1] = i2;
  dist[2 * (size_t)g] = d1; dist[2 * (size_t)g + 1] = d2;
  if (pass) pass[g] = (uint8_t)(i1 >= 0 && d1 <= max_dist && (long long)d1 * ratio_den < (long long)d2 * ratio_num);
}

}  // namespace

int launch_hamming_mma(const uint32_t* q, int nq, const uint32_t* t, int nt, int batch, int num_sms, void** ws, size_t* ws_cap,
                       cudaStream_t s);  // hamming_mma.cu

namespace {
int device_sms() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
    return 148;
  return sms;
}
// 0 = by size, 1 = the integer-ALU kernel, 2 = the tensor-core kernel.  Process-wide; initialised from SFE_HAMMING=alu|mma,
// changed by sfe_hamming_impl() (tests and measurements run both kernels on the same inputs).
int g_impl = [] {
  const char* e = getenv("SFE_HAMMING");
  if (!e) return 0;
  return e[0] == 'a' ? 1 : (e[0] == 'm' ? 2 : 0);
}();
int forced_impl() { return g_impl; }
}  // namespace

int hamming_set_impl(int impl) {
  const int old = g_impl;
  if (impl >= 0 && impl <= 2) g_impl = impl;
  return old;
}

int launch_hamming256(const uint32_t* q, int nq, const uint32_t* t, int nt, int batch, int ratio_num,
                      int ratio_den, int max_dist, int32_t* idx, int32_t* dist, uint8_t* pass, void** ws, size_t* ws_cap, cudaStream_t s) {
  if (nq <= 0 || batch <= 0) return 0;
  if (nt > (1 << IDX_BITS)) return -(int)cudaErrorInvalidValue;
  const int sms = device_sms();
  const int total = batch * nq;
  // The tensor-core kernel (hamming_mma.cu) pays a 128 x 256 tile per MMA whatever is valid in it; tiny problems stay on
  // the ALU kernel, whose cost is proportional to the comparisons.
  const int impl = forced_impl();
  const bool mma = nt > 0 && (impl == 2 || (impl == 0 && (long long)total * nt >= (1ll << 22)));
  if (mma) {
    const int rows = launch_hamming_mma(q, nq, t, nt, batch, sms, ws, ws_cap, s);
    if (rows < 0) return rows;
    hamming_finalize_kernel<<<(total + 255) / 256, 256, 0, s>>>((const uint2*)*ws, nq, rows, total, ratio_num, ratio_den,
                                                                max_dist, idx, dist, pass);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 2 : -(int)e;
  }
  int qblocks = (nq + HM_QPB - 1) / HM_QPB;
  // split the train set when the query tiles alone cannot fill the GPU (4 CTAs per SM target)
  int splits = 1;
  const int target = sms * 4;
  if (nt > 0 && qblocks * batch < target) splits = min((target + qblocks * batch - 1) / (qblocks * batch), (nt + HM_TT - 1) / HM_TT);
  if (splits < 1) splits = 1;
  // per-context key workspace [batch][splits][nq], grown on demand (calls are stream-ordered)
  size_t need = (size_t)batch * splits * nq * sizeof(uint2);
  if (need > *ws_cap) {
    if (*ws) cudaFree(*ws);
    cudaError_t e = cudaMalloc(ws, need);
    if (e != cudaSuccess) { *ws = nullptr; *ws_cap = 0; return -(int)e; }
    *ws_cap = need;
  }
  uint2* g_keys = (uint2*)*ws;
  dim3 grid(qblocks, splits, batch);
  hamming_kernel<<<grid, HM_THREADS, 0, s>>>((const uint4*)q, nq, (const uint4*)t, nt, splits, g_keys);
  hamming_finalize_kernel<<<(total + 255) / 256, 256, 0, s>>>(g_keys, nq, splits, total, ratio_num, ratio_den,
                                                              max_dist, idx, dist, pass);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 2 : -(int)e;
}
