// Throughput of ex2.approx.ftz.f32 (MUFU.EX2), of a degree-3 polynomial exp2 on the FMA pipe, and of cvt.rn.bf16x2.f32
// per SM and clock:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 mufu_rate.cu -o mufu_rate
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^x for x <= 0: floor/frac split, cubic on [0,1), exponent by integer add
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);
  const float f = x - fl;
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((int)fl << 23));
}
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = -0.001f * (threadIdx.x + i);
  uint32_t acc = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = ex2(a[i]) - 1.0f;
      if (MODE == 1) a[i] = ex2_poly(a[i]) - 1.0f;
      if (MODE == 2) { __nv_bfloat162 h = __floats2bfloat162_rn(a[i], a[(i + 1) & 7]); acc += *reinterpret_cast<uint32_t*>(&h); a[i] += 1.0f; }
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
}
template <int MODE> void run(const char* name) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(out, iters);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * 1024 * iters * 8;
  printf("%-28s %.1f Gop/s = %.2f per clk per SM at 1.92 GHz (%.3f ms)\n", name, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.92e9, ms);
  cudaFree(out);
}
int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("polynomial exp2 (FMA pipe)");
  run<2>("cvt.rn.bf16x2.f32");
  return 0;
}
