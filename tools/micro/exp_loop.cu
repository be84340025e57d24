// Throughput of the softmax inner loop of attention.cu without tensor memory: per thread a row of 64 fp32 scores in
// registers; P = exp2(s * scale - m) as FFMA2 -> 2 x MUFU.EX2 -> F2FP pack -> FADD2 row sum, repeated.  Variants drop
// pieces of the chain to show what each costs; run with 1, 2 and 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 exp_loop.cu -o exp_loop
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t d; asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_rn(float lo, float hi) { uint32_t y; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t pack_trunc(float lo, float hi) { return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632); }
// Cody-Waite + degree-3 polynomial exp2 on the FMA pipe (x <= 0): 2^x = 2^floor(x) * p(frac)
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);
  const float f = x - fl;
  float p = fmaf(f, 0.05550357f, 0.24022649f);
  p = fmaf(p, f, 0.69314720f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + ((int)fl << 23));
}

// MODE 0: MUFU only; 1: + pack (rn); 2: + FADD2 sum; 3: full (FFMA2 + MUFU + pack + FADD2); 4: full with truncating pack;
// 5: full, 1 of 4 exponentials by polynomial; 6: full, sum accumulated from packed pairs is dropped (ones-column variant);
template <int MODE>
__global__ void __launch_bounds__(512) k(const float* in, uint32_t* out, int iters, float scale) {
  float s[64];
#pragma unroll
  for (int c = 0; c < 64; c++) s[c] = in[(threadIdx.x * 64 + c) & 4095];
  uint32_t acc_pk = 0;
  uint64_t acc2[4] = {0, 0, 0, 0};
  const uint64_t sc2 = f2_pack(scale, scale);
  float m = 1.0f;
  for (int it = 0; it < iters; it++) {
    const uint64_t nm2 = f2_pack(-m, -m);
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 64; c += 2) {
      float x0, x1;
      if (MODE >= 3) {
        f2_unpack(f2_fma(f2_pack(s[c], s[c + 1]), sc2, nm2), x0, x1);
      } else {
        x0 = s[c] - m; x1 = s[c + 1] - m;
      }
      float e0, e1;
      if (MODE == 5 && (c & 6) == 6) { e0 = ex2_poly(x0); e1 = ex2_poly(x1); }
      else { e0 = ex2a(x0); e1 = ex2a(x1); }
      if (MODE == 0) pk[c >> 1] = __float_as_uint(e0) ^ __float_as_uint(e1);
      else if (MODE == 4) pk[c >> 1] = pack_trunc(e0, e1);
      else pk[c >> 1] = pack_rn(e0, e1);
      if (MODE >= 2 && MODE != 6) acc2[(c >> 1) & 3] = f2_add(acc2[(c >> 1) & 3], f2_pack(e0, e1));
    }
#pragma unroll
    for (int c = 0; c < 32; c++) acc_pk ^= pk[c];
    m += 1e-3f;
  }
  float a0, a1, b0, b1;
  f2_unpack(f2_add(acc2[0], acc2[1]), a0, a1);
  f2_unpack(f2_add(acc2[2], acc2[3]), b0, b1);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc_pk ^ __float_as_uint(a0 + a1 + b0 + b1);
}

template <int MODE>
void run(const char* name, const float* in, uint32_t* out) {
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 2000;
    k<MODE><<<148, warps * 32>>>(in, out, 10, 0.2f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(in, out, iters, 0.2f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double el = 148.0 * warps * 32 * 64 * iters;
    printf("%-44s %2d warps/SM: %6.2f elements/clk/SM (at 1.92 GHz)  %.3f ms\n", name, warps, el / (ms * 1e-3) / 148 / 1.92e9, ms);
  }
}
int main() {
  float* in; uint32_t* out;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4);
  float h[4096];
  for (int i = 0; i < 4096; i++) h[i] = -0.01f * (i % 97);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("MUFU.EX2 only", in, out);
  run<1>("MUFU + F2FP pack", in, out);
  run<2>("MUFU + F2FP + FADD2", in, out);
  run<3>("FFMA2 + MUFU + F2FP + FADD2 (kernel's loop)", in, out);
  run<4>("same, truncating PRMT pack", in, out);
  run<5>("same, 1 of 4 exponentials on the FMA pipe", in, out);
  run<6>("FFMA2 + MUFU + F2FP, no row sum", in, out);
  return 0;
}
