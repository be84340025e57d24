// GroupNorm (two-pass, deterministic), LayerNorm and row softmax.  All HBM-bound: 16-byte vector
// accesses, channel vectors mapped to threadIdx.x so that a warp reads 512 contiguous bytes of a pixel.
#include "../../include/onedc_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace onedc {

__host__ __device__ inline int gn_chunk_pixels(long long hw) { return hw <= 1024 ? 16 : (hw <= 65536 ? 64 : 256); }

__device__ __forceinline__ void load8(const void* base, int dtype, long long elem_off, float* v) {
  if (dtype == DT_BF16) {
    uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
    v[0] = bf16lo(q.x); v[1] = bf16hi(q.x); v[2] = bf16lo(q.y); v[3] = bf16hi(q.y);
    v[4] = bf16lo(q.z); v[5] = bf16hi(q.z); v[6] = bf16lo(q.w); v[7] = bf16hi(q.w);
  } else {
    const float4* f = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    float4 a = __ldg(f), b = __ldg(f + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void store8_bf16(void* base, long long elem_off, const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = q;
}

struct GnSrc {
  const void* x0; const void* x1;
  int c0, c1; long long ld0, ld1;
  int dtype;
};

// partial[((n*chunks + chunk)*C + c)*2 + {0,1}] = sum / sum of squares of channel c over the chunk's pixels
__global__ void __launch_bounds__(256) gn_partial_kernel(GnSrc s, long long hw, int chunk_px, int chunks, float* partial) {
  __shared__ float red[8][32][16];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int C = s.c0 + s.c1;
  const int v = blockIdx.x * 32 + tx;
  const int chunk = blockIdx.y, n = blockIdx.z;
  const bool active = v * 8 < C;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; j++) sum[j] = sq[j] = 0.f;
  if (active) {
    const bool first = v * 8 < s.c0;
    const void* base = first ? s.x0 : s.x1;
    const long long ld = first ? s.ld0 : s.ld1;
    const int ch = first ? v * 8 : v * 8 - s.c0;
    const long long p0 = (long long)chunk * chunk_px;
    long long p1 = p0 + chunk_px;
    if (p1 > hw) p1 = hw;
    for (long long p = p0 + ty; p < p1; p += 8) {
      float x[8];
      load8(base, s.dtype, ((long long)n * hw + p) * ld + ch, x);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        sum[j] += x[j];
        sq[j] += x[j] * x[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    red[ty][tx][j] = sum[j];
    red[ty][tx][8 + j] = sq[j];
  }
  __syncthreads();
  if (ty == 0 && active) {
    float* dst = partial + (((long long)n * chunks + chunk) * C + v * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) {     // fixed order => deterministic
        a += red[k][tx][j];
        b += red[k][tx][8 + j];
      }
      dst[2 * j] = a;
      dst[2 * j + 1] = b;
    }
  }
}

__global__ void __launch_bounds__(128) gn_finalize_kernel(const float* partial, int chunks, int C, int groups, long long hw,
                                                          float eps, float* stats) {
  __shared__ double sh[2][128];
  const int g = blockIdx.x, n = blockIdx.y;
  const int cpg = C / groups;
  const long long entries = (long long)chunks * cpg;
  double a = 0.0, b = 0.0;
  for (long long e = threadIdx.x; e < entries; e += 128) {
    const int chunk = (int)(e / cpg), ch = g * cpg + (int)(e % cpg);
    const float* src = partial + (((long long)n * chunks + chunk) * C + ch) * 2;
    a += (double)src[0];
    b += (double)src[1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double cnt = (double)hw * cpg;
    const double mean = sh[0][0] / cnt;
    double var = sh[1][0] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((long long)n * groups + g) * 2] = (float)mean;
    stats[((long long)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(GnSrc s, long long hw, int groups, const float* stats,
                                                       const float* gamma, const float* beta, int silu, void* out,
                                                       long long out_ld, int px_per_block) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int C = s.c0 + s.c1;
  const int v = blockIdx.x * 32 + tx;
  const int n = blockIdx.z;
  if (v * 8 >= C) return;
  const int cpg = C / groups;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = v * 8 + j, g = c / cpg;
    const float mean = stats[((long long)n * groups + g) * 2], rstd = stats[((long long)n * groups + g) * 2 + 1];
    const float ga = gamma[c];
    sc[j] = rstd * ga;
    sh[j] = beta[c] - mean * rstd * ga;
  }
  const bool first = v * 8 < s.c0;
  const void* base = first ? s.x0 : s.x1;
  const long long ld = first ? s.ld0 : s.ld1;
  const int ch = first ? v * 8 : v * 8 - s.c0;
  const long long p0 = (long long)blockIdx.y * px_per_block;
  long long p1 = p0 + px_per_block;
  if (p1 > hw) p1 = hw;
  for (long long p = p0 + ty; p < p1; p += 8) {
    float x[8];
    load8(base, s.dtype, ((long long)n * hw + p) * ld + ch, x);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float y = x[j] * sc[j] + sh[j];
      if (silu) y = __fdividef(y, 1.f + __expf(-y));
      x[j] = y;
    }
    store8_bf16(out, ((long long)n * hw + p) * out_ld + v * 8, x);
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, C <= 1280 (<= 5 vectors of 8 per lane), bf16 in / bf16 out.
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* x, long long ld, long long rows, int c,
                                                        const float* gamma, const float* beta, float eps,
                                                        __nv_bfloat16* out, long long out_ld) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31, nvec = c >> 3;
  float v[5][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      load8(x, DT_BF16, row * ld + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; j++) sum += v[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)c + eps);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; j++) y[j] = (v[i][j] - mean) * rstd * gamma[vi * 8 + j] + beta[vi * 8 + j];
      store8_bf16(out, row * out_ld + vi * 8, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row softmax: fp32 scores -> bf16 probabilities (one warp per row).  Columns >= valid get 0.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* s, long long ld, long long rows, int cols, int valid,
                                                           const int* valid_per_batch, int rows_per_batch, float scale,
                                                           __nv_bfloat16* out, long long out_ld) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  if (valid_per_batch != nullptr) valid = valid_per_batch[row / rows_per_batch];
  const float* r = s + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < valid; c += 32) m = fmaxf(m, r[c] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int c = lane; c < valid; c += 32) sum += __expf(r[c] * scale - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  __nv_bfloat16* o_ = out + row * out_ld;
  for (int c = lane; c < cols; c += 32)
    o_[c] = __float2bfloat16(c < valid ? __expf(r[c] * scale - m) * inv : 0.f);
}

}  // namespace onedc

using namespace onedc;

extern "C" int64_t onedc_groupnorm_ws_floats(int32_t n_img, int64_t hw, int32_t c_total) {
  const int px = gn_chunk_pixels(hw);
  const int64_t chunks = (hw + px - 1) / px;
  return (int64_t)n_img * chunks * c_total * 2;
}

extern "C" int onedc_groupnorm_stats(const void* x0, int32_t c0, int64_t ld0, const void* x1, int32_t c1, int64_t ld1,
                                     int32_t in_dtype, int32_t n_img, int64_t hw, int32_t groups, float eps,
                                     float* partial, float* stats, void* stream) {
  const int C = c0 + c1;
  ONEDC_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && C % groups == 0 && ld0 % 8 == 0 && ld1 % 8 == 0, "groupnorm: bad channels");
  GnSrc s{x0, x1, c0, c1, ld0, ld1, in_dtype};
  const int px = gn_chunk_pixels(hw);
  const int chunks = (int)((hw + px - 1) / px);
  ONEDC_CHECK(chunks <= 65535 && n_img <= 65535, "groupnorm: grid too large");
  dim3 grid((C / 8 + 31) / 32, chunks, n_img), block(32, 8);
  gn_partial_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(s, hw, px, chunks, partial);
  count_launch();
  gn_finalize_kernel<<<dim3(groups, n_img), 128, 0, (cudaStream_t)stream>>>(partial, chunks, C, groups, hw, eps, stats);
  count_launch();
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_groupnorm_apply(const void* x0, int32_t c0, int64_t ld0, const void* x1, int32_t c1, int64_t ld1,
                                     int32_t in_dtype, int32_t n_img, int64_t hw, int32_t groups, const float* stats,
                                     const float* gamma, const float* beta, int32_t silu, void* out, int64_t out_ld,
                                     void* stream) {
  const int C = c0 + c1;
  ONEDC_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && C % groups == 0 && out_ld % 8 == 0, "groupnorm: bad channels");
  GnSrc s{x0, x1, c0, c1, ld0, ld1, in_dtype};
  const int px = hw <= 4096 ? 16 : 64;
  const int chunks = (int)((hw + px - 1) / px);
  ONEDC_CHECK(chunks <= 65535 && n_img <= 65535, "groupnorm: grid too large");
  dim3 grid((C / 8 + 31) / 32, chunks, n_img), block(32, 8);
  gn_apply_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(s, hw, groups, stats, gamma, beta, silu, out, out_ld, px);
  count_launch();
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_layernorm(const void* x, int64_t ld, int64_t rows, int32_t c, const float* gamma, const float* beta,
                               float eps, void* out, int64_t out_ld, void* stream) {
  ONEDC_CHECK(c % 8 == 0 && c <= 1280 && ld % 8 == 0 && out_ld % 8 == 0, "layernorm: C must be a multiple of 8, <= 1280");
  const int blocks = (int)((rows + 7) / 8);
  layernorm_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, rows, c, gamma, beta, eps,
                                                             (__nv_bfloat16*)out, out_ld);
  count_launch();
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_softmax_rows(const float* scores, int64_t ld, int64_t rows, int32_t cols, int32_t valid, float scale,
                                  void* out, int64_t out_ld, void* stream) {
  const int blocks = (int)((rows + 7) / 8);
  softmax_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scores, ld, rows, cols, valid, nullptr, 1, scale,
                                                                (__nv_bfloat16*)out, out_ld);
  count_launch();
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_softmax_rows_batched(const float* scores, int64_t ld, int64_t rows, int32_t cols,
                                          const int32_t* valid_per_batch, int32_t rows_per_batch, float scale, void* out,
                                          int64_t out_ld, void* stream) {
  const int blocks = (int)((rows + 7) / 8);
  softmax_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scores, ld, rows, cols, cols, valid_per_batch,
                                                                rows_per_batch, scale, (__nv_bfloat16*)out, out_ld);
  count_launch();
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
