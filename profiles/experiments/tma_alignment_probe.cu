#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ int wait_bar(unsigned bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 22)) return 0;
  }
  return 1;
}
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int* status, unsigned bytes, int cx, int cy) {
  __shared__ __align__(1024) float tile[1024];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
    if (MODE == 0)   // CUTLASS form with an L2 cache hint
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                   ::"r"(smem_addr(tile)), "l"(&m), "r"(smem_addr(&bar)), "r"(cx), "r"(cy), "l"(0x1000000000000000ull) : "memory");
    if (MODE == 1)   // shared::cta destination
      asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_addr(tile)), "l"(&m), "r"(smem_addr(&bar)), "r"(cx), "r"(cy) : "memory");
    if (MODE == 2)   // plain
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_addr(tile)), "l"(&m), "r"(smem_addr(&bar)), "r"(cx), "r"(cy) : "memory");
  }
  int ok = wait_bar(smem_addr(&bar), 0);
  if (lane == 0) *status = ok;
  if (ok) for (int i = lane; i < 256; i += 32) out[i] = tile[i];
}
int main(int argc, char** argv) {
  const int mode = atoi(argv[1]), desc = atoi(argv[2]), cx = atoi(argv[3]), cy = atoi(argv[4]);
  const int W = 640, H = 480;
  std::vector<float> h((size_t)W * H);
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) h[(size_t)y * W + x] = y * 1000 + x;
  float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m; memset(&m, 0, sizeof(m));
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t strides[1] = {(cuuint64_t)W * 4};
  cuuint32_t box[2] = {16, 16}, es[2] = {1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE; CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (desc == 1) { l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; }
  if (desc == 2) { sw = CU_TENSOR_MAP_SWIZZLE_128B; box[0] = 32; box[1] = 8; l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; }
  if (desc == 3) { dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; dims[0] = W * 4; box[0] = 64; }
  CUresult r = ((EncodeTiledFn)fn)(&m, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  float* out; int* st; cudaMalloc(&out, 1024); cudaMalloc(&st, 4);
  cudaMemset(out, 0, 1024); cudaMemset(st, 0xff, 4);
  if (mode == 0) k<0><<<1, 32>>>(m, out, st, 1024, cx, cy); else if (mode == 1) k<1><<<1, 32>>>(m, out, st, 1024, cx, cy); else k<2><<<1, 32>>>(m, out, st, 1024, cx, cy);
  cudaError_t e = cudaDeviceSynchronize();
  float ho[256]; int hs = -2; cudaMemcpy(ho, out, 1024, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
  printf("c=(%d,%d) mode %d desc %d: enc %d sync=%s status=%d tile[0]=%g tile[17]=%g tile[255]=%g\n", cx, cy, mode, desc, (int)r, cudaGetErrorString(e), hs, ho[0], ho[17], ho[255]);
  return 0;
}
