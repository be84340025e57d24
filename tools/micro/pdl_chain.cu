// Microbenchmark: per-kernel cost of a dependent chain of small kernels inside a CUDA graph, with and without
// programmatic dependent launch, for three trigger placements.  nvcc -arch=sm_100a -O3 pdl_chain.cu -o pdl_chain
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

template <int MODE>   // 0: no PDL instructions, 1: trigger at start + wait, 2: wait, trigger at end, 3: wait only
__global__ void __launch_bounds__(256) k(const float* __restrict__ in, float* __restrict__ out, int n, int work) {
  __shared__ float s[256];
  s[threadIdx.x] = threadIdx.x;         // "prologue"
  __syncthreads();
  if (MODE == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (MODE != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = i < n ? in[i] : 0.f;
  for (int j = 0; j < work; j++) v = v * 1.0001f + s[(threadIdx.x + j) & 255];
  if (MODE == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (i < n) out[i] = v;
}

template <int MODE>
float run(int blocks, int work, int chain, bool pdl, int smem_bytes) {
  int n = blocks * 256;
  float *a, *b;
  cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
  cudaMemset(a, 0, n * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int c = 0; c < chain; c++) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st; cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    const float* in = (c & 1) ? b : a; float* out = (c & 1) ? a : b;
    cudaLaunchKernelEx(&cfg, k<MODE>, in, out, n, work);
  }
  cudaStreamEndCapture(st, &g);
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed %s\n", cudaGetErrorString(e)); return -1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; w++) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int w = 0; w < 10; w++) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  cudaFree(a); cudaFree(b);
  return ms * 1000.f / (10 * chain);
}

int main() {
  const int chain = 500;
  int cfgs[][3] = {{148, 100, 0}, {148, 2000, 0}, {1184, 100, 0}, {1184, 2000, 0}, {148, 2000, 190 * 1024}, {36, 2000, 190 * 1024}, {148, 20000, 0}};
  for (auto& c : cfgs) {
    printf("blocks %5d work %5d smem %6d | us/kernel: plain %.2f | pdl-attr no-instr %.2f | start-trigger %.2f | end-trigger %.2f | wait-only %.2f\n",
           c[0], c[1], c[2], run<0>(c[0], c[1], chain, false, c[2]), run<0>(c[0], c[1], chain, true, c[2]),
           run<1>(c[0], c[1], chain, true, c[2]), run<2>(c[0], c[1], chain, true, c[2]), run<3>(c[0], c[1], chain, true, c[2]));
  }
  return 0;
}
