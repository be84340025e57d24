// Shared host/device definitions for the onedc_b200 kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace onedc {

enum DType { DT_BF16 = 0, DT_F32 = 1 };
enum Act { ACT_NONE = 0, ACT_LRELU = 1, ACT_SILU = 2, ACT_GELU = 3 };
enum EpiMode { EPI_PLAIN = 0, EPI_PAIR_LRELU = 1, EPI_GEGLU = 2 };
enum StoreMode { ST_NORMAL = 0, ST_PIXSHUF = 1, ST_TRANSPOSED = 2, ST_QUAD = 3 };

// error handling: every C-ABI entry returns 0 or a negative code; message via onedc_last_error().
void set_error(const char* fmt, ...);
#define ONEDC_CHECK(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::onedc::set_error(__VA_ARGS__); \
      return -1;                      \
    }                                 \
  } while (0)
#define ONEDC_CUDA(call)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      ::onedc::set_error("%s:%d CUDA error %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

int sm_count();
void count_launch();          // bumps the launch counter read by onedc_launch_count()
bool pdl_enabled();           // programmatic dependent launch (default off; ONEDC_PDL=1 / onedc_set_pdl enable)

// Programmatic dependent launch (opt-in).  With it every kernel of this library is launched with the
// programmatic-stream-serialization attribute, so in a stream (or a captured graph) kernel i+1 may become resident while kernel i is still running:
// its prologue (barrier init, TMEM allocation, descriptor prefetch, parameter math) overlaps the tail of kernel i.
// Contract kept by every kernel here: ALL threads execute pdl_wait() before the first global-memory access that can
// alias anything an earlier kernel reads or writes (which also makes the chain transitive).  pdl_trigger() is only
// used late (igemm: after the CTA's last MMA is issued), never before a TMEM allocation: a dependent CTA that grabbed
// TMEM first and then blocked in pdl_wait() would starve a co-resident CTA of the still-running primary.  Measured
// on B200 (tools/micro/pdl_chain.cu): inside a graph an early trigger is slower than none; wait-only / late trigger
// saves ~0.3 us per kernel boundary.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// The same with thread-block clusters of `cluster_x` consecutive CTAs: the hardware only starts a cluster when all of its
// CTAs can be resident at once, so CTAs that wait for each other (the splits of one split-K tile) cannot be stranded
// behind kernels of other streams.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// GELU (exact, erf form) for bf16 outputs.  erf(z) = z * P(z^2) on |z| <= 3.2 (degree-8 polynomial in z^2, fitted for
// absolute error: 6.2e-5 in fp32 Horner form; erf(3.2) = 1 - 6e-6 beyond), i.e. 11 FMA-pipe instructions instead of the ~30 of
// erff.  The GEGLU epilogue of the UNet's feed-forward GEMMs (K = 320..1280, 128 x 256 accumulators per tile) was bound by
// exactly that: 10 k clocks of epilogue per tile against 2.6 k clocks of MMA.  The error in gelu(x) is below 0.5 |x| 6.2e-5,
// 1/60 of a bf16 half-ulp of the result.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fmaxf(x * 0.70710678118654752f, -3.2f), 3.2f);
  const float t = z * z;
  float p = 2.804641676e-08f;
  p = fmaf(p, t, -1.468352581e-06f);
  p = fmaf(p, t, 3.372799397e-05f);
  p = fmaf(p, t, -4.514517528e-04f);
  p = fmaf(p, t, 3.961231007e-03f);
  p = fmaf(p, t, -2.439187257e-02f);
  p = fmaf(p, t, 1.101094308e-01f);
  p = fmaf(p, t, -3.747264627e-01f);
  p = fmaf(p, t, 1.128165502e+00f);
  const float hx = 0.5f * x;
  return fmaf(hx, z * p, hx);
}

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  switch (act) {
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_SILU: return v / (1.f + __expf(-v));
    case ACT_GELU: return gelu_erf(v);
    default: return v;
  }
}

}  // namespace onedc
