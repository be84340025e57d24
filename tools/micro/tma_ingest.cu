// Microbenchmark: how fast can ONE SM ingest operand tiles through TMA (cp.async.bulk.tensor, SWIZZLE_128B, 128-byte
// rows) from L2-resident data, as a function of the bytes kept in flight?  No MMA: a slot is released as soon as it
// lands.  Prints B/clk/SM and GB/s aggregate for 148 CTAs, plus the single-load round-trip latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tma_ingest.cu -o tma_ingest -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// tensor: [rows][C] bf16; box {64, box_rows}.  Each CTA walks its own row range repeatedly.
__global__ void __launch_bounds__(128, 1) ingest(const __grid_constant__ CUtensorMap map, int box_rows, int stages, int iters,
                                                int rows_per_cta, int chunks, long long* clocks, int producers, int variant) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t full_all[4][16];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const uint32_t bytes = box_rows * 128;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; w++) for (int i = 0; i < 16; i++) mbar_init(&full_all[w][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int wp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0 || wp >= producers) return;
  uint64_t* full = full_all[wp];
  smem += (size_t)wp * stages * bytes;
  rows_per_cta /= producers;
  iters /= producers;
  const int row0 = blockIdx.x * rows_per_cta * producers + wp * rows_per_cta;
  const int boxes_per_pass = (rows_per_cta / box_rows) * chunks;
  if (variant & 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map)) : "memory");
  if (variant & 2) {
    // burst mode: issue `stages` loads back to back, then wait for all of them; repeat
    long long tb0 = clock64(), t_issue = 0;
    const int boxes_per_pass2 = (rows_per_cta / box_rows) * chunks;
    for (int it = 0; it < iters / stages; it++) {
      long long a = clock64();
      for (int s2 = 0; s2 < stages; s2++) {
        const int b = (it * stages + s2) % boxes_per_pass2;
        mbar_expect_tx(&full[s2], bytes);
        tma_load_2d(smem + (size_t)s2 * bytes, &map, &full[s2], (b % chunks) * 64, row0 + (b / chunks) * box_rows);
      }
      t_issue += clock64() - a;
      for (int s2 = 0; s2 < stages; s2++) mbar_wait(&full[s2], it & 1);
    }
    if (wp == 0) { clocks[blockIdx.x] = clock64() - tb0; clocks[148 + blockIdx.x] = t_issue; }
    return;
  }
  long long t0 = clock64();
  // prologue: fill the ring
  int issued = 0, done = 0;
  auto issue = [&](int i) {
    const int s = i % stages;
    const int b = i % boxes_per_pass;
    const int chunk = b % chunks, rb = b / chunks;
    mbar_expect_tx(&full[s], bytes);
    tma_load_2d(smem + (size_t)s * bytes, &map, &full[s], chunk * 64, row0 + rb * box_rows);
  };
  for (; issued < stages && issued < iters; issued++) issue(issued);
  for (; done < iters; done++) {
    const int s = done % stages;
    mbar_wait(&full[s], (done / stages) & 1);
    if (issued < iters) issue(issued++);       // slot is free the moment it landed
  }
  long long t1 = clock64();
  if (wp == 0) clocks[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &f, 12000, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)f;
  const int ctas = 148;
  const int rpc = 512, C = 256;
  size_t rows = (size_t)ctas * rpc;
  void* buf; cudaMalloc(&buf, rows * C * 2); cudaMemset(buf, 1, rows * C * 2);
  long long* clk; cudaMalloc(&clk, 2 * ctas * 8);
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int box_rows : {16, 128}) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t ones[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int bytes = box_rows * 128;
    for (int variant : {0, 1, 2, 3}) {
      for (int stages : {2, 4, 8}) {
        const int producers = 1;
        const int iters = 4096 * 8;
        const int grid = 148;
        ingest<<<grid, 128, producers * stages * bytes + 1024>>>(map, box_rows, stages, 64, rpc, C / 64, clk, producers, variant);
        ingest<<<grid, 128, producers * stages * bytes + 1024>>>(map, box_rows, stages, iters, rpc, C / 64, clk, producers, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[296]; cudaMemcpy(h, clk, 2 * grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0, avi = 0; for (int i = 0; i < grid; i++) { avg += (double)h[i]; avi += (double)h[148 + i]; } avg /= grid; avi /= grid;
        printf("box %3d rows, variant %d (1=prefetch desc, 2=burst), stages %d: %.0f clk per box, issue-only %.0f clk per box\n",
               box_rows, variant, stages, avg / iters, (variant & 2) ? avi / iters : 0.0);
      }
    }
  }
  return 0;
}
