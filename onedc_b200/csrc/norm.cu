// GroupNorm (two-pass, deterministic), LayerNorm and row softmax.  All HBM-bound: 16-byte vector
// accesses, channel vectors mapped to threadIdx.x so that a warp reads 512 contiguous bytes of a pixel.
#include "../../include/onedc_b200.h"
#include "common.cuh"
#include <stdlib.h>

#include "ptx.cuh"

namespace onedc {

// pixels per block of the statistics / apply passes.  Small tensors: about one block per SM (every block ends
// with same-address atomics, which serialise); big streaming tensors: ~4 blocks per SM to cover HBM latency.
__host__ __device__ inline int gn_chunk_pixels(long long hw, int slabs, int c_total) {
  const long long bytes = hw * c_total * 2;
  long long target = (bytes >= (64ll << 20) ? 592 : (bytes >= (8ll << 20) ? 296 : 148)) / (slabs > 0 ? slabs : 1);
  if (target < 1) target = 1;
  long long p = (hw + target - 1) / target;
  p = (p + 15) / 16 * 16;
  return (int)(p < 16 ? 16 : p);
}

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
__device__ __forceinline__ void unpack8(const uint4& q, float* v) {
  v[0] = bf16lo(q.x); v[1] = bf16hi(q.x); v[2] = bf16lo(q.y); v[3] = bf16hi(q.y);
  v[4] = bf16lo(q.z); v[5] = bf16hi(q.z); v[6] = bf16lo(q.w); v[7] = bf16hi(q.w);
}
__device__ __forceinline__ void store8_bf16(void* base, long long elem_off, const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]); q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]); q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off) = q;
}

// SiLU through ONE special-function op: y * sigmoid(y) = 0.5 y (1 + tanh(y / 2)).  exp + reciprocal are two MUFU ops
// per element and the XU pipe (16 lanes/clk/SM, shared with the bf16 conversions) is what the apply pass runs out
// of once the tensor is L2-resident; tanh.approx is accurate to ~2^-11, a quarter of a bf16 ulp of the result.
__device__ __forceinline__ float silu_fast(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  return fmaf(0.5f * y, t, 0.5f * y);
}

struct GnSrc {
  const void* x0; const void* x1;
  int c0, c1; long long ld0, ld1;
  int dtype;
};

// Pass 1.  Block = 256 threads = VX channel-vectors (8 channels each) x PY pixel lanes over one chunk of pixels.
// Every block reduces its chunk to per-group (sum, sum of squares) in a fixed order and writes them to
// part[n][chunk*slabs + slab][group][2]; the last block of an image to finish (ticket counter) sums the block
// partials in a fixed order in fp64 and writes mean / rstd: deterministic, no extra launch, no memset.
template <int VX>
__global__ void __launch_bounds__(256, 4) gn_stats_kernel(GnSrc s, long long hw, int chunk_px, int groups, float eps,
                                                       float* part, float* stats, unsigned int* counters,
                                                       const int* valid_px) {
  pdl_wait();
  constexpr int PY = 256 / VX;
  __shared__ float red[PY][VX][16];
  __shared__ float chan[VX * 8][2];
  __shared__ unsigned int ticket_s;
  const int tx = threadIdx.x % VX, ty = threadIdx.x / VX;
  const int C = s.c0 + s.c1;
  const int v = blockIdx.x * VX + tx;
  const int chunk = blockIdx.y, n = blockIdx.z;
  const bool active = v * 8 < C;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; j++) sum[j] = sq[j] = 0.f;
  if (active) {
    const bool first = v * 8 < s.c0;
    const void* base_ptr = first ? s.x0 : s.x1;
    const long long ld = first ? s.ld0 : s.ld1;
    const int ch = first ? v * 8 : v * 8 - s.c0;
    // grid-stride over tiles of 8*PY pixels: at any moment the whole grid reads one compact address window
    // (DRAM page locality); the assignment is fixed, so the summation order is too
    (void)chunk_px;
    constexpr int TILE = 8 * PY;
    for (long long base = (long long)chunk * TILE; base < hw; base += (long long)gridDim.y * TILE) {
      if (base + TILE <= hw && s.dtype == DT_BF16) {
        // raw 16-byte vectors stay packed while in flight (4 registers each): 8 loads outstanding per thread at
        // 4 resident blocks per SM = 128 KB in flight per SM
        uint4 q[8];
        const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(base_ptr);
#pragma unroll
        for (int u = 0; u < 8; u++)
          q[u] = __ldg(reinterpret_cast<const uint4*>(bp + ((long long)n * hw + base + ty + u * PY) * ld + ch));
#pragma unroll
        for (int u = 0; u < 8; u++) {
          float x[8];
          unpack8(q[u], x);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            sum[j] += x[j];
            sq[j] += x[j] * x[j];
          }
        }
      } else {
        const long long pend = base + TILE < hw ? base + TILE : hw;
        for (long long p = base + ty; p < pend; p += PY) {
          float x[8];
          load8(base_ptr, s.dtype, ((long long)n * hw + p) * ld + ch, x);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            sum[j] += x[j];
            sq[j] += x[j] * x[j];
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    red[ty][tx][j] = sum[j];
    red[ty][tx][8 + j] = sq[j];
  }
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < PY; k++) {     // fixed order
        a += red[k][tx][j];
        b += red[k][tx][8 + j];
      }
      chan[tx * 8 + j][0] = a;
      chan[tx * 8 + j][1] = b;
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  const int nblk = gridDim.x * gridDim.y;
  const int blk = blockIdx.y * gridDim.x + blockIdx.x;
  if ((int)threadIdx.x < groups) {
    // this block's contribution to every group (zero for groups outside its channel slab)
    const int g = threadIdx.x;
    const int c0 = blockIdx.x * VX * 8;
    int c1 = c0 + VX * 8;
    if (c1 > C) c1 = C;
    int lo = g * cpg, hi = lo + cpg;
    if (lo < c0) lo = c0;
    if (hi > c1) hi = c1;
    float a = 0.f, b = 0.f;
    for (int c = lo; c < hi; c++) {
      a += chan[c - c0][0];
      b += chan[c - c0][1];
    }
    float2* dst = reinterpret_cast<float2*>(part) + ((long long)n * nblk + blk) * groups + g;
    __stcg(dst, make_float2(a, b));
  }
  // ---- last block of this image finalises
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket_s = atomicAdd(&counters[n], 1u);
  __syncthreads();
  if (ticket_s != (unsigned)nblk - 1) return;
  __threadfence();
  // 256 threads = 8 sub-sums x 32 groups per pass; fixed assignment => deterministic
  for (int gb = 0; gb < groups; gb += 32) {
    const int g = gb + (threadIdx.x & 31), sub = threadIdx.x >> 5;
    double a = 0.0, b = 0.0;
    if (g < groups) {
      const float2* src = reinterpret_cast<const float2*>(part) + (long long)n * nblk * groups + g;
      int i = sub;
      for (; i + 24 < nblk; i += 32) {
        float2 f[4];
#pragma unroll
        for (int u = 0; u < 4; u++) f[u] = __ldcg(src + (long long)(i + 8 * u) * groups);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          a += (double)f[u].x;
          b += (double)f[u].y;
        }
      }
      for (; i < nblk; i += 8) {
        const float2 f = __ldcg(src + (long long)i * groups);
        a += (double)f.x;
        b += (double)f.y;
      }
    }
    __syncthreads();
    double* dred = reinterpret_cast<double*>(&red[0][0][0]);      // 2 x 256 doubles = 4 KB of the 16 KB array
    dred[threadIdx.x] = a;
    dred[256 + threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x < 32 && g < groups) {
      double ta = 0.0, tb = 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        ta += dred[k * 32 + threadIdx.x];
        tb += dred[256 + k * 32 + threadIdx.x];
      }
      // zero-padded pixels (edge attention windows) add nothing to the sums: only the count changes
      const double cnt = (double)(valid_px != nullptr ? (long long)valid_px[n] : hw) * cpg;
      const double mean = ta / cnt;
      double var = tb / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[((long long)n * groups + g) * 2] = (float)mean;
      stats[((long long)n * groups + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  if (threadIdx.x == 0) counters[n] = 0;       // self-cleaning
}

template <int VX>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(GnSrc s, long long hw, int groups, const float* stats,
                                                       const double* acc0, const double* acc1, int acc_per_channel,
                                                       float eps, const float* gamma, const float* beta, int silu,
                                                       void* out, long long out_ld, int px_per_block) {
  pdl_wait();
  constexpr int PY = 256 / VX;
  const int tx = threadIdx.x % VX, ty = threadIdx.x / VX;
  const int C = s.c0 + s.c1;
  const int v = blockIdx.x * VX + tx;
  const int n = blockIdx.z;
  const int cpg = C / groups;
  __shared__ float gstat[64][2];
  if (acc0 != nullptr && acc_per_channel) {
    // statistics were accumulated per CHANNEL by the producing igemm epilogues (one accumulator per source, so a
    // channel concatenation needs nothing extra): 8 threads fold the channels of one group, 32 groups per pass
    for (int gb = 0; gb < groups; gb += 32) {
      const int g = gb + ((int)threadIdx.x >> 3), sub = threadIdx.x & 7;
      double a = 0.0, b = 0.0;
      if (g < groups) {
        for (int c = g * cpg + sub; c < (g + 1) * cpg; c += 8) {
          const double* src = c < s.c0 ? acc0 + ((size_t)n * s.c0 + c) * 2 : acc1 + ((size_t)n * s.c1 + (c - s.c0)) * 2;
          a += __ldcg(src);
          b += __ldcg(src + 1);
        }
      }
#pragma unroll
      for (int o = 4; o >= 1; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (sub == 0 && g < groups) {
        const double cnt = (double)hw * cpg;
        const double mean = a / cnt;
        double var = b / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        gstat[g][0] = (float)mean;
        gstat[g][1] = (float)(1.0 / sqrt(var + (double)eps));
      }
    }
    __syncthreads();
  } else if (acc0 != nullptr) {
    // statistics were accumulated per group by the producing igemm epilogue(s) (single source only)
    if ((int)threadIdx.x < groups) {
      const int g = threadIdx.x;
      const double* src = acc0 + ((size_t)n * groups + g) * 2;
      const double a = __ldcg(src), b = __ldcg(src + 1);
      const double cnt = (double)hw * cpg;
      const double mean = a / cnt;
      double var = b / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      gstat[g][0] = (float)mean;
      gstat[g][1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
  }
  if (v * 8 >= C) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int c = v * 8 + j, g = c / cpg;
    const float mean = acc0 != nullptr ? gstat[g][0] : stats[((long long)n * groups + g) * 2];
    const float rstd = acc0 != nullptr ? gstat[g][1] : stats[((long long)n * groups + g) * 2 + 1];
    const float ga = gamma[c];
    sc[j] = rstd * ga;
    sh[j] = beta[c] - mean * rstd * ga;
  }
  const bool first = v * 8 < s.c0;
  const void* base = first ? s.x0 : s.x1;
  const long long ld = first ? s.ld0 : s.ld1;
  const int ch = first ? v * 8 : v * 8 - s.c0;
  (void)px_per_block;
  constexpr int TILE = 8 * PY;
  for (long long tb = (long long)blockIdx.y * TILE; tb < hw; tb += (long long)gridDim.y * TILE) {
    if (tb + TILE <= hw && s.dtype == DT_BF16) {
      uint4 q[8];
      const __nv_bfloat16* bp = reinterpret_cast<const __nv_bfloat16*>(base);
#pragma unroll
      for (int u = 0; u < 8; u++)
        q[u] = __ldg(reinterpret_cast<const uint4*>(bp + ((long long)n * hw + tb + ty + u * PY) * ld + ch));
#pragma unroll
      for (int u = 0; u < 8; u++) {
        float x[8];
        unpack8(q[u], x);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float y = x[j] * sc[j] + sh[j];
          if (silu) y = silu_fast(y);
          x[j] = y;
        }
        store8_bf16(out, ((long long)n * hw + tb + ty + u * PY) * out_ld + v * 8, x);
      }
    } else {
      const long long pend = tb + TILE < hw ? tb + TILE : hw;
      for (long long p = tb + ty; p < pend; p += PY) {
        float x[8];
        load8(base, s.dtype, ((long long)n * hw + p) * ld + ch, x);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float y = x[j] * sc[j] + sh[j];
          if (silu) y = silu_fast(y);
          x[j] = y;
        }
        store8_bf16(out, ((long long)n * hw + p) * out_ld + v * 8, x);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm, bf16 in / bf16 out, C <= 1280.  LPR lanes share a row and each holds up to five 16-byte vectors of it, so a
// warp works on 32 / LPR rows at once: LPR = 8 for C <= 320 (the UNet's 9216 x 320 token matrix: every lane busy, where
// one warp per row left 24 of 32 lanes idle in the second round), 16 for C <= 640, 32 above.  Warps walk the rows
// grid-stride, so a big matrix is a few resident waves instead of 100 k eight-row blocks.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* x, long long ld, long long rows, int c,
                                                        const float* gamma, const float* beta, float eps,
                                                        __nv_bfloat16* out, long long out_ld) {
  pdl_wait();
  constexpr int RPW = 32 / LPR;                                      // rows per warp and pass
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const int nvec = c >> 3;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * 8;
  const float inv_c = 1.f / (float)c;
  for (long long row = warp0 * RPW + rsel; row - rsel < rows; row += nwarps * RPW) {
    const bool rok = row < rows;
    uint4 q[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const int vi = sub + i * LPR;
      q[i] = make_uint4(0, 0, 0, 0);
      if (rok && vi < nvec) q[i] = __ldg(reinterpret_cast<const uint4*>(x + row * ld) + vi);
    }
    float v[5][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      unpack8(q[i], v[i]);
#pragma unroll
      for (int j = 0; j < 8; j++) sum += v[i][j];                    // vectors beyond C are zeros
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      if (sub + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const int vi = sub + i * LPR;
      if (rok && vi < nvec) {
        const float4* g4 = reinterpret_cast<const float4*>(gamma + vi * 8);
        const float4* b4 = reinterpret_cast<const float4*>(beta + vi * 8);
        const float4 g0 = __ldg(g4), g1 = __ldg(g4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) y[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
        store8_bf16(out, row * out_ld + vi * 8, y);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Row softmax: fp32 scores -> bf16 probabilities (one warp per row).  Columns >= valid get 0.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* s, long long ld, long long rows, int cols, int valid,
                                                           const int* valid_per_batch, int rows_per_batch, float scale,
                                                           __nv_bfloat16* out, long long out_ld) {
  pdl_wait();
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

static int gn_target_blocks() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ONEDC_GN_BLOCKS");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static void gn_geometry(int C, long long hw, int* vx, int* slabs, int* px, int* chunks) {
  const int nvec = C / 8;
  *vx = nvec <= 16 ? 16 : 32;
  *slabs = (nvec + *vx - 1) / *vx;
  *px = gn_chunk_pixels(hw, *slabs, C);
  if (gn_target_blocks() > 0) {
    long long t = gn_target_blocks() / *slabs;
    if (t < 1) t = 1;
    long long p = (hw + t - 1) / t;
    p = (p + 15) / 16 * 16;
    *px = (int)(p < 16 ? 16 : p);
  }
  *chunks = (int)((hw + *px - 1) / *px);
  const int tile = 8 * (256 / *vx);
  const long long ntiles = (hw + tile - 1) / tile;
  if (*chunks > ntiles) *chunks = (int)ntiles;
  if (*chunks < 1) *chunks = 1;
}

extern "C" int64_t onedc_groupnorm_ws_floats(int32_t n_img, int64_t hw, int32_t c_total) {
  int vx, slabs, px, chunks;
  gn_geometry(c_total, hw, &vx, &slabs, &px, &chunks);
  return (int64_t)n_img * slabs * chunks * 64 * 2;       // block partials [n][blocks][groups<=64][2]
}

extern "C" int onedc_groupnorm_stats(const void* x0, int32_t c0, int64_t ld0, const void* x1, int32_t c1, int64_t ld1,
                                     int32_t in_dtype, int32_t n_img, int64_t hw, int32_t groups, float eps,
                                     float* partial, float* stats, uint32_t* counters, const int32_t* valid_px, void* stream) {
  const int C = c0 + c1;
  ONEDC_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && C % groups == 0 && ld0 % 8 == 0 && ld1 % 8 == 0 && groups <= 64,
              "groupnorm: bad channels");
  ONEDC_CHECK(counters != nullptr && partial != nullptr, "groupnorm: scratch is required");
  GnSrc s{x0, x1, c0, c1, ld0, ld1, in_dtype};
  int vx, slabs, px, chunks;
  gn_geometry(C, hw, &vx, &slabs, &px, &chunks);
  ONEDC_CHECK(chunks <= 65535 && n_img <= 65535, "groupnorm: grid too large");
  dim3 grid(slabs, chunks, n_img);
  if (vx == 16)
    ONEDC_CUDA(launch_k(gn_stats_kernel<16>, grid, 256, 0, (cudaStream_t)stream, s, hw, px, groups, eps, partial, stats, counters, valid_px));
  else
    ONEDC_CUDA(launch_k(gn_stats_kernel<32>, grid, 256, 0, (cudaStream_t)stream, s, hw, px, groups, eps, partial, stats, counters, valid_px));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_groupnorm_apply(const void* x0, int32_t c0, int64_t ld0, const void* x1, int32_t c1, int64_t ld1,
                                     int32_t in_dtype, int32_t n_img, int64_t hw, int32_t groups, const float* stats,
                                     const double* acc0, const double* acc1, int32_t acc_per_channel, float eps,
                                     const float* gamma, const float* beta, int32_t silu, void* out, int64_t out_ld,
                                     void* stream) {
  ONEDC_CHECK(stats != nullptr || (acc0 != nullptr && (c1 == 0 || (acc_per_channel && acc1 != nullptr))),
              "groupnorm_apply: no statistics");
  const int C = c0 + c1;
  ONEDC_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && C % groups == 0 && out_ld % 8 == 0, "groupnorm: bad channels");
  GnSrc s{x0, x1, c0, c1, ld0, ld1, in_dtype};
  const int nvec = C / 8;
  int vx_, slabs_, px, chunks;
  gn_geometry(C, hw, &vx_, &slabs_, &px, &chunks);
  ONEDC_CHECK(chunks <= 65535 && n_img <= 65535, "groupnorm: grid too large");
  if (nvec <= 16) {
    dim3 grid((nvec + 15) / 16, chunks, n_img);
    ONEDC_CUDA(launch_k(gn_apply_kernel<16>, grid, 256, 0, (cudaStream_t)stream, s, hw, groups, stats, acc0, acc1, acc_per_channel, eps, gamma, beta, silu, out, out_ld, px));
  } else {
    dim3 grid((nvec + 31) / 32, chunks, n_img);
    ONEDC_CUDA(launch_k(gn_apply_kernel<32>, grid, 256, 0, (cudaStream_t)stream, s, hw, groups, stats, acc0, acc1, acc_per_channel, eps, gamma, beta, silu, out, out_ld, px));
  }
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_layernorm(const void* x, int64_t ld, int64_t rows, int32_t c, const float* gamma, const float* beta,
                               float eps, void* out, int64_t out_ld, void* stream) {
  ONEDC_CHECK(c % 8 == 0 && c <= 1280 && ld % 8 == 0 && out_ld % 8 == 0, "layernorm: C must be a multiple of 8, <= 1280");
  ONEDC_CHECK(reinterpret_cast<uintptr_t>(gamma) % 16 == 0 && reinterpret_cast<uintptr_t>(beta) % 16 == 0, "layernorm: gamma / beta must be 16-byte aligned");
  const int lpr = c <= 320 ? 8 : (c <= 640 ? 16 : 32);
  long long blocks = (rows + 8 * (32 / lpr) - 1) / (8 * (32 / lpr));
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;       // grid-stride beyond a few waves
  auto kern = lpr == 8 ? layernorm_kernel<8> : (lpr == 16 ? layernorm_kernel<16> : layernorm_kernel<32>);
  ONEDC_CUDA(launch_k(kern, (int)blocks, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, ld, rows, c, gamma, beta, eps,
                                                 (__nv_bfloat16*)out, out_ld));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_softmax_rows(const float* scores, int64_t ld, int64_t rows, int32_t cols, int32_t valid, float scale,
                                  void* out, int64_t out_ld, void* stream) {
  const int blocks = (int)((rows + 7) / 8);
  ONEDC_CUDA(launch_k(softmax_rows_kernel, blocks, 256, 0, (cudaStream_t)stream, scores, ld, rows, cols, valid, nullptr, 1, scale,
                                                                (__nv_bfloat16*)out, out_ld));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_softmax_rows_batched(const float* scores, int64_t ld, int64_t rows, int32_t cols,
                                          const int32_t* valid_per_batch, int32_t rows_per_batch, float scale, void* out,
                                          int64_t out_ld, void* stream) {
  const int blocks = (int)((rows + 7) / 8);
  ONEDC_CUDA(launch_k(softmax_rows_kernel, blocks, 256, 0, (cudaStream_t)stream, scores, ld, rows, cols, cols, valid_per_batch,
                                                                rows_per_batch, scale, (__nv_bfloat16*)out, out_ld));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
