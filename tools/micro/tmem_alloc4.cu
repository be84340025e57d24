// Do four CTAs per SM get four disjoint 128-column tensor-memory allocations?  Each CTA writes a CTA-unique pattern over its
// allocation, spins for a while so that all CTAs of the SM are alive together, reads it back and counts mismatches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tmem_alloc4.cu -o tmem_alloc4
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void __launch_bounds__(128, 4) k(uint32_t* bases, uint32_t* bad, int cols, long long spin) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0) { bases[blockIdx.x * 2] = base; bases[blockIdx.x * 2 + 1] = smid; }
  const uint32_t addr = base + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < cols; c++) {
    uint32_t v = (blockIdx.x << 16) | (threadIdx.x << 8) | c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  uint32_t nbad = 0;
  for (int c = 0; c < cols; c++) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(addr + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (v != ((blockIdx.x << 16) | (threadIdx.x << 8) | c)) nbad++;
  }
  if (nbad) atomicAdd(bad, nbad);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
}
int main() {
  uint32_t *bases, *bad;
  cudaMalloc(&bases, 148 * 8 * 2 * 4);
  cudaMalloc(&bad, 4);
  for (int cols = 64; cols <= 256; cols *= 2) {
    const int ctas = 148 * 4;
    cudaMemset(bad, 0, 4);
    k<<<ctas, 128>>>(bases, bad, cols, 2000000);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t hb[148 * 8], nb;
    cudaMemcpy(hb, bases, ctas * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&nb, bad, 4, cudaMemcpyDeviceToHost);
    printf("cols=%d: %s, mismatches %u; SM of CTA 0 = %u, allocations on it:", cols, cudaGetErrorString(e), nb, hb[1]);
    for (int i = 0; i < ctas; i++) if (hb[2 * i + 1] == hb[1]) printf(" cta%d@0x%x", i, hb[2 * i]);
    printf("\n");
  }
  return 0;
}
