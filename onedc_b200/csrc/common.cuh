// Shared host/device definitions for the onedc_b200 kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  switch (act) {
    case ACT_LRELU: return v > 0.f ? v : v * slope;
    case ACT_SILU: return v / (1.f + __expf(-v));
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

}  // namespace onedc
