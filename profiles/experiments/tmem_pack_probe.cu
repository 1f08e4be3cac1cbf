// What does tcgen05.ld ... .pack::16b return?  One warp stores column c of its 32 TMEM lanes as (0x8000 + c) << 16 | (c + 256 * lane),
// loads 32x32b.x16.pack::16b and prints the registers of lanes 0 and 5.
#include <cuda_runtime.h>
#include <stdio.h>
__device__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(unsigned* out) {
  __shared__ unsigned slot;
  const int lane = threadIdx.x;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&slot)), "r"(64) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = slot;
  unsigned v[32];
  for (int c = 0; c < 32; ++c) v[c] = ((0x8000u + c) << 16) | (unsigned)(c + 256 * lane);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               ::"r"(base), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
                 "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(base) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[lane * 16 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64) : "memory");
}
int main() {
  unsigned* d; cudaMalloc(&d, 32 * 16 * 4);
  k<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h[512]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("sync: %s\n", cudaGetErrorString(e));
  for (int lane : {0, 5}) { printf("lane %d:", lane); for (int i = 0; i < 16; ++i) printf(" %08x", h[lane * 16 + i]); printf("\n"); }
  return 0;
}
