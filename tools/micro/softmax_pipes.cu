// What bounds the softmax warps of the attention kernel: per-SM-per-clock rates of the instructions they issue.
//   ex2.approx.ftz.f32 / .f16x2 / .bf16x2 (MUFU), cvt.rn.{bf16x2,f16x2}.f32 (XU), max.f32 2- and 3-input (ALU),
//   fma.rn.f32 (FMA pipe), and tcgen05.ld 32x32b.x32 (TMEM read bytes per clock per SM) with 4 and 8 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 softmax_pipes.cu -o softmax_pipes
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvth2(float a, float b) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ uint32_t cvtb2(float a, float b) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }

template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = -0.001f * (threadIdx.x + i); u[i] = 0xB800B800u + threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = ex2f(a[i]);
      if (MODE == 1) u[i] = ex2h2(u[i]);
      if (MODE == 2) u[i] = ex2b2(u[i]);
      if (MODE == 3) { u[i] += cvth2(a[i], a[(i + 1) & 7]); }
      if (MODE == 4) { u[i] += cvtb2(a[i], a[(i + 1) & 7]); }
      if (MODE == 5) a[i] = fmaxf(a[i], a[(i + 3) & 7] - 0.0f);
      if (MODE == 6) a[i] = max3(a[i], a[(i + 3) & 7], a[(i + 5) & 7]);
      if (MODE == 7) a[i] = fmaf(a[i], a[(i + 3) & 7], a[(i + 5) & 7]);
      if (MODE == 8) { u[i] = ex2h2(cvth2(a[i], a[(i + 1) & 7]) + u[i]); }      // cvt + packed exp chain per 2 elements
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double per_instr) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * 1024 * iters * 8;
  printf("%-34s %7.2f thread-instr/clk/SM  (%5.2f elements/clk/SM)  %.3f ms\n", name, ops / (ms * 1e-3) / 148 / 1.92e9,
         per_instr * ops / (ms * 1e-3) / 148 / 1.92e9, ms);
  cudaFree(out);
}

// ---- TMEM read rate: NW warps of one CTA per SM read 32 lanes x 32 columns x 4 B per instruction, back to back
__global__ void __launch_bounds__(256, 1) tmem_rd(float* out, int iters, int x64) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t r[64];
    const uint32_t col = (uint32_t)((it * 64) & 255);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(base + col));
    if (x64)
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
            "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
            "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]),
            "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
          : "r"(base + col + 32));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // consume every register with static indices (a dynamic index would turn r[] into a local-memory array)
#pragma unroll
    for (int i = 0; i < 32; i++) acc += __uint_as_float(r[i]);
    if (x64) {
#pragma unroll
      for (int i = 32; i < 64; i++) acc += __uint_as_float(r[i]);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)(t1 - t0);
  if (threadIdx.x == 0) out[148 * 256 + blockIdx.x] = (float)(t1 - t0);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

void run_tmem(int warps, int x64) {
  float* out;
  cudaMalloc(&out, (148 * 256 + 148) * 4);
  const int iters = 20000;
  tmem_rd<<<148, warps * 32>>>(out, 100, x64);
  tmem_rd<<<148, warps * 32>>>(out, iters, x64);
  cudaDeviceSynchronize();
  float clk;
  cudaMemcpy(&clk, out + 148 * 256, 4, cudaMemcpyDeviceToHost);
  const double bytes = (double)warps * 32 * 32 * 4 * (x64 ? 2 : 1) * iters;
  printf("tcgen05.ld 32x32b.x32 %s, %d warps/SM: %.1f B/clk/SM (%.0f clk per iteration)  [%s]\n", x64 ? "x2 per wait" : "x1 per wait",
         warps, bytes / clk, clk / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cvt.rn.f16x2.f32 (+iadd)", 2);
  run<4>("cvt.rn.bf16x2.f32 (+iadd)", 2);
  run<5>("max.f32 (2-input, +fadd)", 1);
  run<6>("max.f32 (3-input)", 2);
  run<7>("fma.rn.f32", 1);
  run<8>("cvt.f16x2 + iadd + ex2.f16x2", 2);
  run_tmem(4, 0);
  run_tmem(4, 1);
  run_tmem(8, 1);
  return 0;
}
