// Implicit-GEMM convolution / GEMM for sm_100a.
//
//   out[pixel, n] = epilogue( sum_{tap, c} A[pixel shifted by tap, c] * W[tap][n][c] )
//
// * A is NHWC bf16.  One CTA tile = 128 output pixels arranged as a TH x TW spatial box, so that for
//   every filter tap the A operand of the tile is ONE TMA box {64 ch, TW, 1, TH, 1} of a 5-D tensor map
//   shifted by the tap offset; TMA's out-of-bounds zero fill IS the convolution's zero padding (and the
//   K / M / N tails).  No im2col buffer exists anywhere.  Stride-2 convs use the parity view
//   (c' = px*S + c, x/2, py, y/2, n) of the same tensor, a plain GEMM is the degenerate H = 1, 1-tap case,
//   a batched GEMM takes one weight matrix per image, and a channel concat is a second tensor map
//   visited by the same K loop.
// * MMA: tcgen05.mma cta_group::1 kind::f16, M = 128, N = BN (16..256), K = 16 per instruction, operands in
//   128B-swizzled shared memory written by TMA, fp32 accumulators in TMEM (two buffers of 256 columns so the
//   epilogue of tile i overlaps the main loop of tile i+1).
// * Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..9 =
//   epilogue (tcgen05.ld -> bias/activation/pair-ops/residual -> global, fused GroupNorm statistics).  Persistent
//   CTAs, static round-robin tile schedule.  Producer and issuer loops run on the whole warp with uniform state
//   (see the comment at the role dispatch): single-lane scalar code was the first bound of this kernel.
// * Three tilings, chosen per layer by the host (igemm_launch): the tap tile above; the column-copy tile (16 x 8
//   pixels, ONE A box per tap column shared by its three taps, separate A / B rings) for stride-1 multi-tap convs
//   that run several tiles per CTA; and its transposed form (weights as the M operand, 32 x 8 pixels as N = 256)
//   for cout <= 128.  What bounds the big layers is shared-memory bandwidth (MMA operand reads + TMA writes,
//   128 B/clk/SM), so the tilings differ in bytes moved per MMA clock: 192 / 170 / 149.  DESIGN.md section 8.
// * Split-K (template SPLITK) for layers with too few tiles to fill the GPU: fp32 partial tiles parked in a
//   coalesced layout, every CTA fetches its row slice of all splits with bulk copies and finishes it.
//
// The SIMT kernel at the bottom evaluates the same descriptor with scalar loops; it exists to check the
// tensor-core kernel on the GPU (tests, impl = 1) and is never used by the decode path.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/onedc_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace onedc {

constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 128;       // 128 rows x 64 bf16
constexpr int kSmemBudget = 200 * 1024;  // operand ring
constexpr int kEpiWarps = 8;            // two warps per TMEM lane quarter, splitting the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct IgemmParams {
  // geometry
  int n_img, H, W;        // output dims
  int TH, TW, tiles_y, tiles_x;
  int cout, BN, n_tiles, m_tiles;
  int taps, kchunks[2], c1_off;
  int tap_dc[9], tap_dx[9], tap_dp[9], tap_dy[9];
  int w_batched;
  int stages, stage_bytes;
  // column-copy mode (stride-1 multi-tap convs, no split-K): the tile is a 16 x 8 pixel box; per 64-channel chunk the
  // A operand is loaded ONCE per distinct tap column offset dx as a (16 + dy span) x 8 box (cm_rows x 1024 bytes),
  // and every tap of that column reads it through a descriptor that starts (dy - min dy) * 1024 bytes into the box.
  // A 3x3 conv moves 3 x 18/16 = 3.4 A tiles per chunk instead of 9: the L2->SM ingest (~70 B/clk/SM) is what bounds
  // this kernel, not the tensor pipe.  A and B (weights) travel through separate rings.
  int colmode, cm_groups, cm_dx[3], cm_nt[3], cm_tap[3][3], cm_row[3][3], cm_miny, cm_rows;
  int a_slots, a_slot_bytes, b_slots, b_slot_bytes;
  // transposed column-copy mode (cout <= 128): the weights are the M operand (128 output channels) and a 32 x 8 pixel
  // box the N = 256 operand, D^T[cout][pixel].  With N = 128 every MMA reads 8 KB of operands per 64 clocks, which
  // alone saturates the 128 B/clk of shared memory; N = 256 reads 12 KB per 128 clocks.  Accumulator lanes are then
  // output channels and the epilogue stores one bf16 per lane (64 contiguous bytes per warp and pixel).
  int transposed;
  // residual on the tensor core (column-copy modes): res_chunks 64-channel chunks of the residual tensor (tensor map 1)
  // follow the K loop as one more column group -- the centre pixel, one "tap" -- whose weights are the identity stored
  // behind the taps (onedc_igemm_desc.w_identity_tap).  The epilogue then has no residual to load.
  int res_chunks;
  const uint8_t* pf_ptr;  // optional L2 prefetch hint (the next layer's weights)
  long long pf_bytes;
  // optional per-CTA role timing (onedc_igemm_set_debug): 16 clock counters per CTA, see tools/igemm_roles.py
  long long* dbg;
  // epilogue
  const float* bias;
  int epi_mode, act;
  float slope;
  const void* res;
  int res_dtype;
  long long res_ld;
  void* out;
  int out_dtype;
  long long out_ld;
  int out_col_off, store_mode, ps_c, quad;
  int ncols_out;          // cout (plain) or cout/2 (pair modes)
  int vec_ok;
  // split-K (small-M layers): fp32 partial tiles + per-tile arrival counters
  double* gn_acc;         // optional per-GROUP (sum, sum of squares) of the stored outputs, [img][gn_groups][2]
  int gn_ng, gn_groups;   // groups per 32-column chunk (32/8/4/2/1, i.e. 1/4/8/16/32 channels per group), groups per
                          // image; 1 channel per group = per-CHANNEL sums, regrouped by the consuming GroupNorm
  int splits;
  float* ws;
  int* counters;
  int counters_half;
  // raw views for the SIMT checker
  const __nv_bfloat16* a_ptr[2];
  int a_c[2];
  long long a_pix[2];
  int h_in, w_in, ksize, stride;
  const __nv_bfloat16* w_ptr;
  int ktot;
  long long w_row, w_z;
};

// ------------------------------------------------------------------------------------------------
// Epilogue for 16 consecutive output columns of one pixel (shared by both kernels).
//   a[16]: accumulator columns gcol .. gcol+15 (GEMM column space); b[16]: partner columns (pair modes)
//   gcol_a / gcol_b: GEMM columns of a[0] / b[0] (bias index); ocol: first output column
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue16(const IgemmParams& p, int img, int y, int x, int gcol_a, int gcol_b,
                                           int ocol, const float* a, const float* b) {
  int ncols = p.ncols_out - ocol;
  if (ncols <= 0) return;
  if (ncols > 16) ncols = 16;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    float va = a[j];
    if (p.bias != nullptr && j < ncols) va += __ldg(p.bias + gcol_a + j);
    if (p.epi_mode == EPI_PLAIN) {
      v[j] = act_apply(va, p.act, p.slope);
    } else {
      float vb = b[j];
      if (p.bias != nullptr && j < ncols) vb += __ldg(p.bias + gcol_b + j);
      if (p.epi_mode == EPI_PAIR_LRELU)
        v[j] = (va > 0.f ? va : 0.1f * va) + (vb > 0.f ? vb : 0.01f * vb);
      else
        v[j] = va * gelu_erf(vb);
    }
  }
  const long long pix = ((long long)img * p.H + y) * p.W + x;
  if (p.res != nullptr) {
    if (p.res_dtype == DT_BF16) {
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(p.res) + pix * p.res_ld + ocol;
      if (ncols == 16 && p.vec_ok) {
        const uint4* r4 = reinterpret_cast<const uint4*>(r);
        uint4 q0 = __ldg(r4), q1 = __ldg(r4 + 1);
        uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
          v[2 * j] += bf16lo(w[j]);
          v[2 * j + 1] += bf16hi(w[j]);
        }
      } else {
        for (int j = 0; j < ncols; j++) v[j] += __bfloat162float(r[j]);
      }
    } else {
      const float* r = reinterpret_cast<const float*>(p.res) + pix * p.res_ld + ocol;
      for (int j = 0; j < ncols; j++) v[j] += __ldg(r + j);
    }
  }
  if (p.store_mode == ST_TRANSPOSED) {
    const long long pin = (long long)y * p.W + x;
    for (int j = 0; j < ncols; j++) {
      long long o = ((long long)img * p.ncols_out + ocol + j) * p.out_ld + pin;
      if (p.out_dtype == DT_BF16)
        reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16(v[j]);
      else
        reinterpret_cast<float*>(p.out)[o] = v[j];
    }
    return;
  }
  long long opix = pix;
  int oc = ocol;
  if (p.store_mode == ST_PIXSHUF) {
    int q = ocol / p.ps_c;
    oc = ocol - q * p.ps_c;
    opix = ((long long)img * (2 * p.H) + (2 * y + (q >> 1))) * (2 * p.W) + (2 * x + (q & 1));
  } else if (p.store_mode == ST_QUAD) {
    opix = ((long long)img * (2 * p.H) + (2 * y + (p.quad >> 1))) * (2 * p.W) + (2 * x + (p.quad & 1));
  }
  const long long o = opix * p.out_ld + p.out_col_off + oc;
  if (p.out_dtype == DT_BF16) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + o;
    if (ncols == 16 && p.vec_ok) {
      uint4 q0, q1;
      q0.x = pack_bf16x2(v[0], v[1]);
      q0.y = pack_bf16x2(v[2], v[3]);
      q0.z = pack_bf16x2(v[4], v[5]);
      q0.w = pack_bf16x2(v[6], v[7]);
      q1.x = pack_bf16x2(v[8], v[9]);
      q1.y = pack_bf16x2(v[10], v[11]);
      q1.z = pack_bf16x2(v[12], v[13]);
      q1.w = pack_bf16x2(v[14], v[15]);
      reinterpret_cast<uint4*>(dst)[0] = q0;
      reinterpret_cast<uint4*>(dst)[1] = q1;
    } else {
      for (int j = 0; j < ncols; j++) dst[j] = __float2bfloat16(v[j]);
    }
  } else {
    float* dst = reinterpret_cast<float*>(p.out) + o;
    if (ncols == 16 && p.vec_ok) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      for (int j = 0; j < ncols; j++) dst[j] = v[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
struct TileCoord {
  int img, y0, x0, n_tile;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
  TileCoord t;
  int m_tile = tile / p.n_tiles;
  t.n_tile = tile - m_tile * p.n_tiles;
  int per_img = p.tiles_y * p.tiles_x;
  t.img = m_tile / per_img;
  int r = m_tile - t.img * per_img;
  int ty = r / p.tiles_x;
  t.y0 = ty * p.TH;
  t.x0 = (r - ty * p.tiles_x) * p.TW;
  return t;
}

// Reduces NG (= 8, 4, 2 or 1) per-thread values over the 32 lanes of a warp: halving butterfly while more than one
// value is left, plain xor all-reduce afterwards.  On return every lane holds in v[0] the warp total of the value
// with index (lane >> log2(32 / NG)), i.e. 32/NG consecutive lanes share one value.
template <int NG>
__device__ __forceinline__ float warp_group_sum(float* v, int lane) {
  int n = NG;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const bool upper = (lane & off) != 0;
      const int h = n >> 1;
#pragma unroll
      for (int i = 0; i < NG / 2; i++) {
        if (i < h) {
          const float mine = upper ? v[i + h] : v[i];
          const float send = upper ? v[i] : v[i + h];
          v[i] = mine + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return v[0];
}

// per-thread sums of `cpg` consecutive values (cpg = 32 / NG) of a[32] and of their squares, then warp reduction
template <int NG>
__device__ __forceinline__ void chunk_group_stats(const float* a, bool valid, int lane, float* s_out, float* q_out) {
  constexpr int CPG = 32 / NG;
  float gs[NG], gq[NG];
#pragma unroll
  for (int j = 0; j < NG; j++) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; i++) {
      const float x = valid ? a[j * CPG + i] : 0.f;
      s1 += x;
      s2 += x * x;
    }
    gs[j] = s1;
    gq[j] = s2;
  }
  *s_out = warp_group_sum<NG>(gs, lane);
  *q_out = warp_group_sum<NG>(gq, lane);
}

// Epilogue switches of a launch packed into one register (epi_flags): tested per 32-column chunk.  Read from the parameter
// block instead, every test was a dependent LDCU -> UISETP -> BRA.U chain of ~70 clocks, seven per chunk: a third of the
// 9000 clocks that the epilogue of a 128 x 256 tile took (the small-K layers are bound by exactly that).
enum : uint32_t { EF_LRELU = 1, EF_SILU = 2, EF_GELU = 4, EF_RES = 8, EF_RES_BF16 = 16, EF_PIXSHUF = 32, EF_QUAD = 64, EF_OUT_BF16 = 128,
                  EF_PAIR_LRELU = 256 };
__device__ __forceinline__ uint32_t epi_flags(const IgemmParams& p) {
  return (p.act == ACT_LRELU ? EF_LRELU : 0u) | (p.act == ACT_SILU ? EF_SILU : 0u) | (p.act == ACT_GELU ? EF_GELU : 0u) |
         (p.res != nullptr ? EF_RES : 0u) | (p.res_dtype == DT_BF16 ? EF_RES_BF16 : 0u) |
         (p.store_mode == ST_PIXSHUF ? EF_PIXSHUF : 0u) | (p.store_mode == ST_QUAD ? EF_QUAD : 0u) |
         (p.out_dtype == DT_BF16 ? EF_OUT_BF16 : 0u) | (p.epi_mode == EPI_PAIR_LRELU ? EF_PAIR_LRELU : 0u);
}

// Finishes 32 output columns of one pixel: v[] already holds accumulator (+ nothing else) values.
//   a[32]: accumulators of GEMM columns (n0 + c ..), b[32]: partner columns (pair modes only)
//   bias_a / bias_b: shared-memory bias slices aligned with a / b
template <bool PAIR>
__device__ __forceinline__ void epi_finish32(const IgemmParams& p, float* a, const float* b, const float* bias_a,
                                             const float* bias_b, bool valid, int img, int y, int x, long long pix,
                                             int ocol, const uint32_t fl, uint4* stg = nullptr, int lane = 0) {
  float* v = a;
  if (!PAIR) {
    const float4* ba = reinterpret_cast<const float4*>(bias_a);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 a4 = ba[i];
      v[4 * i + 0] += a4.x; v[4 * i + 1] += a4.y; v[4 * i + 2] += a4.z; v[4 * i + 3] += a4.w;
    }
    if (fl & EF_LRELU) {
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = v[i] > 0.f ? v[i] : v[i] * p.slope;
    } else if (fl & EF_SILU) {
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = __fdividef(v[i], 1.f + __expf(-v[i]));
    } else if (fl & EF_GELU) {
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = gelu_erf(v[i]);
    }
  } else {
    const float4* ba = reinterpret_cast<const float4*>(bias_a);
    const float4* bb = reinterpret_cast<const float4*>(bias_b);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 a4 = ba[i], b4 = bb[i];
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float va = a[4 * i + j] + av[j];
        const float vb = b[4 * i + j] + bv[j];
        if (fl & EF_PAIR_LRELU)
          v[4 * i + j] = (va > 0.f ? va : 0.1f * va) + (vb > 0.f ? vb : 0.01f * vb);
        else
          v[4 * i + j] = va * gelu_erf(vb);
      }
    }
  }
  if (!valid && stg == nullptr) return;
  // (fetching the residual rows the same way -- four lanes per row through the tile -- measured slower: the loads' latency
  // then sits in front of two more synchronisations; kept row per lane)
  if (valid && (fl & EF_RES)) {
    if (fl & EF_RES_BF16) {
      const uint4* r4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res) + pix * p.res_ld + ocol);
      uint4 qv[4];
#pragma unroll
      for (int i = 0; i < 4; i++) qv[i] = __ldg(r4 + i);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        v[8 * i + 0] += bf16lo(qv[i].x); v[8 * i + 1] += bf16hi(qv[i].x);
        v[8 * i + 2] += bf16lo(qv[i].y); v[8 * i + 3] += bf16hi(qv[i].y);
        v[8 * i + 4] += bf16lo(qv[i].z); v[8 * i + 5] += bf16hi(qv[i].z);
        v[8 * i + 6] += bf16lo(qv[i].w); v[8 * i + 7] += bf16hi(qv[i].w);
      }
    } else {
      const float4* r4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + pix * p.res_ld + ocol);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float4 f = __ldg(r4 + i);
        v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
      }
    }
  }
  long long opix = pix;
  int oc = ocol;
  if (fl & EF_PIXSHUF) {
    const int qd = ocol / p.ps_c;
    oc = ocol - qd * p.ps_c;
    opix = ((long long)img * (2 * p.H) + (2 * y + (qd >> 1))) * (2 * p.W) + (2 * x + (qd & 1));
  } else if (fl & EF_QUAD) {
    opix = ((long long)img * (2 * p.H) + (2 * y + (p.quad >> 1))) * (2 * p.W) + (2 * x + (p.quad & 1));
  }
  const long long o = opix * p.out_ld + p.out_col_off + oc;
  if (fl & EF_OUT_BF16) {
    uint4 w4[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      w4[i].x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      w4[i].y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      w4[i].z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      w4[i].w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    }
    if (stg != nullptr) {
      // Row-per-lane stores touch 32 different 128-byte lines per instruction (16 bytes each) and the store path, not the
      // tensor pipe, bounded every layer with a short K loop (9000 clocks of epilogue per 128 x 256 tile).  The warp's
      // 32 rows x 64 bytes go through a warp-private shared-memory tile (row pitch 80 bytes: conflict-free) so that
      // four lanes write one row's 64 contiguous bytes: 8 lines per instruction instead of 32.
#pragma unroll
      for (int i = 0; i < 4; i++) stg[lane * 5 + i] = w4[i];
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = 8 * i + (lane >> 2);
        const uint4 wv = stg[r * 5 + (lane & 3)];
        const long long orow = __shfl_sync(0xffffffffu, o, r);
        const int vrow = __shfl_sync(0xffffffffu, valid ? 1 : 0, r);
        if (vrow) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow + (lane & 3) * 8) = wv;
      }
      __syncwarp();
    } else {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
#pragma unroll
      for (int i = 0; i < 4; i++) dst[i] = w4[i];
    }
  } else if (valid) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
    for (int i = 0; i < 8; i++) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

// mbarrier wait on a shared-memory ADDRESS; adds the cycles spent waiting to *acc when role timing is compiled in
template <bool TIMED>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long* acc) {
  if (!TIMED) {
    mbar_wait_a(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait_a(bar, parity);
  *acc += clock64() - t0;
}

// Transposed tiles: every lane owns one output channel.  Adds its (sum, sum of squares) to the fp64 accumulators:
// groups of cpg = cout / gn_groups consecutive channels (cpg a power of two <= 32) are folded by a butterfly first.
__device__ __forceinline__ void flush_channel_stats(const IgemmParams& p, int img, int ch, bool ch_ok, float s1, float s2) {
  const int cpg = p.cout / p.gn_groups;
  if (!ch_ok) s1 = s2 = 0.f;
  for (int o = 1; o < cpg; o <<= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (ch_ok && (ch & (cpg - 1)) == 0) {
    double* dst = p.gn_acc + ((size_t)img * p.gn_groups + ch / cpg) * 2;
    atomicAdd(dst, (double)s1);
    atomicAdd(dst + 1, (double)s2);
  }
}

// k-iteration index -> (tap, source, 64-channel chunk)
struct KIter {
  int tap, src, kc;
};
__device__ __forceinline__ KIter decode_kiter(const IgemmParams& p, int ki) {
  const int per_tap = p.kchunks[0] + p.kchunks[1];
  KIter k;
  k.tap = ki / per_tap;
  const int r = ki - k.tap * per_tap;
  k.src = r >= p.kchunks[0] ? 1 : 0;
  k.kc = k.src ? r - p.kchunks[0] : r;
  return k;
}

template <bool SPLITK, bool TIMED>
__global__ void __launch_bounds__(kThreads, 1)
igemm_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t a_full[4];       // column-copy mode: the A ring (full_bar / empty_bar = B ring)
  __shared__ __align__(8) uint64_t a_empty[4];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[2][256];
  __shared__ __align__(16) uint4 store_stage[kEpiWarps][32 * 5];   // epilogue warps' private 32 x 64-byte transpose tiles (pitch 80 B)

  // broadcast from lane 0 so that the compiler KNOWS the warp index is warp-uniform (role branches stay uniform)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  // barrier init spread over the lanes of warp 0 (28 serial mbarrier.init cost ~0.3 us at the head of every launch)
  if (threadIdx.x < kMaxStages) {
    mbar_init(&full_bar[threadIdx.x], 1);
    mbar_init(&empty_bar[threadIdx.x], 1);
  } else if (threadIdx.x < kMaxStages + 4) {
    mbar_init(&a_full[threadIdx.x - kMaxStages], 1);
    mbar_init(&a_empty[threadIdx.x - kMaxStages], 1);
  } else if (threadIdx.x < kMaxStages + 6) {
    mbar_init(&tmem_full[threadIdx.x - kMaxStages - 4], 1);
    mbar_init(&tmem_empty[threadIdx.x - kMaxStages - 4], kEpiWarps);   // one arrive per epilogue warp
  }
  if (threadIdx.x < kMaxStages + 6) fence_mbar_init();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b);
    if (p.kchunks[1] > 0) tma_prefetch_desc(&map_a1);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();           // everything above overlapped the previous kernel's tail

  if (p.pf_bytes > 0 && threadIdx.x == 64) {
    // L2 prefetch of this CTA's slice of the hinted range (the next layer's weights), in 32 KB requests
    const long long per = ((p.pf_bytes + gridDim.x - 1) / gridDim.x + 127) & ~127ll;
    const long long lo = (long long)blockIdx.x * per;
    const long long hi = lo + per < p.pf_bytes ? lo + per : p.pf_bytes;
    for (long long off = lo; off < hi; off += 32768) {
      const uint32_t sz = (uint32_t)(hi - off < 32768 ? hi - off : 32768);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.pf_ptr + off), "r"(sz) : "memory");
    }
  }

  // work item = (tile, k-split); splits of one tile are adjacent items, i.e. run on different CTAs
  const int nsplit = SPLITK ? p.splits : 1;
  const int total_items = p.m_tiles * p.n_tiles * nsplit;
  const int kiters = p.taps * (p.kchunks[0] + p.kchunks[1]);
  const uint32_t a_bytes = (uint32_t)(p.TH * p.TW) * 128u;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;

  // The producer and the MMA issuer are single-lane jobs, but their loops are executed by the WHOLE warp with only
  // the TMA / MMA / commit instructions under `if (leader)`, and every ring address (slot, barrier, descriptor) is a
  // register that is bumped by a constant per step.  The loop state then lives in uniform registers and ptxas emits
  // bare UTMALDG / UTCHMMA: ~35 scalar instructions per k-step.  (A loop entered under `if (lane == 0)` with
  // barrier[stage] indexing cost ~100 clocks of scalar latency per MMA, which bounded every layer with N <= 128.)
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t afull0 = smem_u32(&a_full[0]), aempty0 = smem_u32(&a_empty[0]);
  const uint32_t tfull0 = smem_u32(&tmem_full[0]), tempty0 = smem_u32(&tmem_empty[0]);
  const uint32_t smem0 = smem_u32(smem);
  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t leader = elect_one();
    long long w_a = 0, w_b = 0;
    const long long t_start = TIMED ? clock64() : 0;
    const int nch = p.kchunks[0] + p.kchunks[1], kch0 = p.kchunks[0], c1_off = p.c1_off;
    if (!SPLITK && p.colmode) {
      const int a_slots = p.a_slots, b_slots = p.b_slots, groups = p.cm_groups;
      const uint32_t a_sz = (uint32_t)p.a_slot_bytes, b_sz = (uint32_t)p.b_slot_bytes;
      const uint32_t smem_b0 = smem0 + (uint32_t)a_slots * a_sz;
      int sa = 0, sb = 0;
      uint32_t pa = 1, pb = 1;                       // parity to wait for on the EMPTY barriers
      uint32_t a_dst = smem0, b_dst = smem_b0, a_fb = afull0, a_eb = aempty0, b_fb = full0, b_eb = empty0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const TileCoord t = decode_tile(p, item);
        const int n0 = t.n_tile * p.BN;
        const int ay = t.y0 + p.cm_miny;
        for (int c = 0; c < nch + p.res_chunks; c++) {
          const bool isres = c >= nch;                     // residual chunk: tensor map 1, centre column, identity "tap"
          const int src = (isres || c >= kch0) ? 1 : 0;
          const int kc = isres ? c - nch : (src ? c - kch0 : c);
          const CUtensorMap* ma = src ? &map_a1 : &map_a0;
          const int kb = isres ? kc * 64 : (src ? c1_off : 0) + kc * 64;
          const int ng = isres ? 1 : groups;
          for (int g = 0; g < ng; g++) {
            const int nt = isres ? 1 : p.cm_nt[g], ax = t.x0 + (isres ? 0 : p.cm_dx[g]);
            mbar_wait_t<TIMED>(a_eb, pa, &w_a);
            if (leader) {
              mbar_expect_tx_a(a_fb, a_sz);
              tma_load_5d_a(a_dst, ma, a_fb, kc * 64, ax, 0, ay, t.img);
            }
            __syncwarp();
            a_dst += a_sz; a_fb += 8; a_eb += 8;
            if (++sa == a_slots) {
              sa = 0; pa ^= 1; a_dst = smem0; a_fb = afull0; a_eb = aempty0;
            }
            for (int j = 0; j < nt; j++) {
              const int tap = isres ? p.taps : p.cm_tap[g][j];
              mbar_wait_t<TIMED>(b_eb, pb, &w_b);
              if (leader) {
                mbar_expect_tx_a(b_fb, b_bytes);
                tma_load_3d_a(b_dst, &map_b, b_fb, kb, n0, tap);
              }
              __syncwarp();
              b_dst += b_sz; b_fb += 8; b_eb += 8;
              if (++sb == b_slots) {
                sb = 0; pb ^= 1; b_dst = smem_b0; b_fb = full0; b_eb = empty0;
              }
            }
          }
        }
      }
    } else {
      const int stages = p.stages;
      const uint32_t st_sz = (uint32_t)p.stage_bytes;
      int stage = 0;
      uint32_t pe = 1;
      uint32_t dst = smem0, fb = full0, eb = empty0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int tile = item / nsplit, split = item - tile * nsplit;
        const TileCoord t = decode_tile(p, tile);
        const int n0 = t.n_tile * p.BN;
        const int k0 = (int)((long long)kiters * split / nsplit), k1 = (int)((long long)kiters * (split + 1) / nsplit);
        // (tap, chunk) counters advance incrementally: no division in the loop
        int tap = k0 / nch, c = k0 - tap * nch;
        int dc = p.tap_dc[tap], ax = t.x0 + p.tap_dx[tap], dp = p.tap_dp[tap], ay = t.y0 + p.tap_dy[tap];
        int bz = p.w_batched ? t.img : tap;
        for (int ki = k0; ki < k1; ki++) {
          const int src = c >= kch0 ? 1 : 0;
          const int kc = src ? c - kch0 : c;
          const CUtensorMap* ma = src ? &map_a1 : &map_a0;
          mbar_wait_t<TIMED>(eb, pe, &w_b);
          if (leader) {
            mbar_expect_tx_a(fb, a_bytes + b_bytes);
            tma_load_5d_a(dst, ma, fb, dc + kc * 64, ax, dp, ay, t.img);
            tma_load_3d_a(dst + kABytes, &map_b, fb, (src ? c1_off : 0) + kc * 64, n0, bz);
          }
          __syncwarp();
          dst += st_sz; fb += 8; eb += 8;
          if (++stage == stages) {
            stage = 0; pe ^= 1; dst = smem0; fb = full0; eb = empty0;
          }
          if (++c == nch && ki + 1 < k1) {
            c = 0;
            ++tap;
            dc = p.tap_dc[tap];
            ax = t.x0 + p.tap_dx[tap];
            dp = p.tap_dp[tap];
            ay = t.y0 + p.tap_dy[tap];
            if (!p.w_batched) bz = tap;
          }
        }
      }
    }
    if (TIMED && leader) {
      long long* o = p.dbg + (size_t)blockIdx.x * 16;
      o[0] = clock64() - t_start;
      o[1] = w_a;
      o[2] = w_b;
    }
    if (SPLITK) {                 // the two cluster barriers of the split-K epilogue need every thread of the cluster
      __syncwarp();
      cluster_arrive();
      cluster_wait();
      cluster_arrive();
      cluster_wait();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t leader = elect_one();
    long long w_a = 0, w_b = 0, w_t = 0;
    const long long t_start = TIMED ? clock64() : 0;
    const uint32_t idesc = umma_idesc_bf16(128, p.transposed ? 256 : p.BN, 0, 0);
    // descriptor = constant high part | (shared address >> 4): advancing a slot / a row offset / 16 K elements is an add
    const uint64_t desc_hi = umma_smem_desc(0, 16, 1024);
    const uint32_t smem_enc = (smem0 & 0x3FFFF) >> 4;
    int it = 0;
    if (!SPLITK && p.colmode) {
      const int nch = p.kchunks[0] + p.kchunks[1], a_slots = p.a_slots, b_slots = p.b_slots, groups = p.cm_groups;
      const bool transposed = p.transposed != 0;
      const uint32_t a_enc = (uint32_t)p.a_slot_bytes >> 4, b_enc = (uint32_t)p.b_slot_bytes >> 4;
      const uint32_t smem_b_enc = smem_enc + (uint32_t)a_slots * a_enc;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      uint32_t a_lo = smem_enc, b_lo = smem_b_enc, a_fb = afull0, a_eb = aempty0, b_fb = full0, b_eb = empty0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, it++) {
        const int acc = it & 1;
        mbar_wait_t<TIMED>(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1, &w_t);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0;
        for (int c = 0; c < nch + p.res_chunks; c++) {
          const bool isres = c >= nch;
          const int ng = isres ? 1 : groups;
          for (int g = 0; g < ng; g++) {
            const int nt = isres ? 1 : p.cm_nt[g];
            mbar_wait_t<TIMED>(a_fb, pa, &w_a);
            for (int j = 0; j < nt; j++) {
              const uint32_t row_enc = (uint32_t)(isres ? -p.cm_miny : p.cm_row[g][j]) * 64u;   // rows of 8 pixels x 128 B = 1024 B
              mbar_wait_t<TIMED>(b_fb, pb, &w_b);
              tc_fence_after();
              if (leader) {
                const uint64_t dpix = desc_hi | (uint64_t)(a_lo + row_enc);
                const uint64_t dwgt = desc_hi | (uint64_t)b_lo;
                const uint64_t da = transposed ? dwgt : dpix;        // M operand
                const uint64_t db = transposed ? dpix : dwgt;        // N operand
                umma_bf16(d_tmem, da, db, idesc, accumulate);
                umma_bf16(d_tmem, da + 2, db + 2, idesc, 1);
                umma_bf16(d_tmem, da + 4, db + 4, idesc, 1);
                umma_bf16(d_tmem, da + 6, db + 6, idesc, 1);
                umma_commit_a(b_eb);
                if (j == nt - 1) umma_commit_a(a_eb);         // all taps of this column have read the box
              }
              __syncwarp();
              accumulate = 1;
              b_lo += b_enc; b_fb += 8; b_eb += 8;
              if (++sb == b_slots) {
                sb = 0; pb ^= 1; b_lo = smem_b_enc; b_fb = full0; b_eb = empty0;
              }
            }
            a_lo += a_enc; a_fb += 8; a_eb += 8;
            if (++sa == a_slots) {
              sa = 0; pa ^= 1; a_lo = smem_enc; a_fb = afull0; a_eb = aempty0;
            }
          }
        }
        if (leader) umma_commit_a(tfull0 + acc * 8);
        __syncwarp();
      }
    } else {
      const int stages = p.stages;
      const uint32_t stage_enc = (uint32_t)p.stage_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t a_lo = smem_enc, fb = full0, eb = empty0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, it++) {
        const int split = item % nsplit;
        const int k0 = (int)((long long)kiters * split / nsplit), k1 = (int)((long long)kiters * (split + 1) / nsplit);
        const int acc = it & 1;
        mbar_wait_t<TIMED>(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1, &w_t);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0;
        for (int ki = k0; ki < k1; ki++) {
          mbar_wait_t<TIMED>(fb, phase, &w_b);
          tc_fence_after();
          if (leader) {
            const uint64_t da = desc_hi | (uint64_t)a_lo;
            const uint64_t db = da + (kABytes >> 4);
            // +32 bytes (16 bf16) along K inside the 128B swizzle atom == +2 in the (addr >> 4) field
            umma_bf16(d_tmem, da, db, idesc, accumulate);
            umma_bf16(d_tmem, da + 2, db + 2, idesc, 1);
            umma_bf16(d_tmem, da + 4, db + 4, idesc, 1);
            umma_bf16(d_tmem, da + 6, db + 6, idesc, 1);
            umma_commit_a(eb);                         // frees the smem slot once these MMAs have read it
          }
          __syncwarp();
          accumulate = 1;
          a_lo += stage_enc; fb += 8; eb += 8;
          if (++stage == stages) {
            stage = 0; phase ^= 1; a_lo = smem_enc; fb = full0; eb = empty0;
          }
        }
        if (leader) umma_commit_a(tfull0 + acc * 8);   // accumulator complete -> epilogue
        __syncwarp();
      }
    }
    if (TIMED && leader) {
      long long* o = p.dbg + (size_t)blockIdx.x * 16;
      o[4] = clock64() - t_start;
      o[5] = w_a;
      o[6] = w_b;
      o[7] = w_t;
    }
    if (leader) pdl_trigger();    // this CTA's MMAs are all issued: the next kernel may start under our epilogue
    if (SPLITK) {
      __syncwarp();
      cluster_arrive();
      cluster_wait();
      cluster_arrive();
      cluster_wait();
    }
  } else {
    // ===================== epilogue warps =====================
    // warp w may only touch TMEM lanes 32*(w%4)..+31; the two warps of a quarter interleave 32-column chunks.
    const int q = warp & 3;
    const int member = (warp - 2) >> 2;
    const int etid = threadIdx.x - 64;                 // 0..255 among epilogue threads
    const int row = q * 32 + lane;
    const int ry = row / p.TW, rx = row - ry * p.TW;
    const bool pair = p.epi_mode != EPI_PLAIN;
    const int half = p.BN >> 1;
    const int out_cols_tile = pair ? half : p.BN;      // output columns produced per tile
    float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};   // fused GroupNorm statistics
    const uint32_t eflags = epi_flags(p);
    uint4* stg_w = (eflags & EF_OUT_BF16) ? store_stage[warp - 2] : nullptr;
    const long long t_epi0 = TIMED ? clock64() : 0;
    int st_img = -1, st_nt = -1;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, it++) {
      const int tile = item / nsplit, split = item - tile * nsplit;
      const int acc = it & 1;
      TileCoord t = decode_tile(p, tile);
      const int y = t.y0 + ry, x = t.x0 + rx;
      const bool valid = (ry < p.TH) && (y < p.H) && (x < p.W);
      const int n0 = t.n_tile * p.BN;
      // stage this tile's bias slice (GEMM column order) in shared memory
      if (etid < p.BN) bias_s[acc][etid] = (p.bias != nullptr && n0 + etid < p.cout) ? __ldg(p.bias + n0 + etid) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (TIMED && threadIdx.x == 64) {
        const long long t0 = clock64();
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
        p.dbg[(size_t)blockIdx.x * 16 + 9] += clock64() - t0;
      } else {
        mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      }
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      const int o0 = t.n_tile * out_cols_tile;         // first output column of this tile
      const long long pix = ((long long)t.img * p.H + y) * p.W + x;
      const bool fast = p.vec_ok && (o0 + out_cols_tile <= p.ncols_out) && (out_cols_tile % 32 == 0) &&
                        (p.store_mode == ST_NORMAL || p.store_mode == ST_QUAD || (p.store_mode == ST_PIXSHUF && p.ps_c % 32 == 0));
      if (!SPLITK && p.transposed) {
        // ---- transposed tile: lane = output channel, TMEM column n = pixel (y0 + n / 8, x0 + n % 8)
        const int ch = q * 32 + lane;
        const bool ch_ok = ch < p.cout;
        const float bias_v = bias_s[acc][ch];
        if (p.gn_acc != nullptr && t.img != st_img) {
          if (st_img >= 0) flush_channel_stats(p, st_img, ch, ch_ok, csum[0], csq[0]);
          csum[0] = csq[0] = 0.f;
          st_img = t.img;
        }
        float s1 = 0.f, s2 = 0.f;
        const long long pix_img = (long long)t.img * p.H * p.W;
        const int nx = p.W - t.x0 < 8 ? p.W - t.x0 : 8;                    // valid pixels per box row (warp-uniform)
        // host guarantees: bf16 output, bf16 (or no) residual, activation none / LeakyReLU (max(v, slope v), slope 1 = none):
        // the per-pixel body must stay tiny, it is unrolled 32 times
        const float slope_eff = p.act == ACT_LRELU ? p.slope : 1.f;
        const bool full_tile = ch_ok && t.y0 + 32 <= p.H && t.x0 + 8 <= p.W;
        const __nv_bfloat16* resp = reinterpret_cast<const __nv_bfloat16*>(p.res);
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out) + p.out_col_off + ch;
        for (int c = member * 32; c < 256; c += 64) {
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          // 32 pixels = 4 box rows of 8: issue all the residual loads of the chunk before the TMEM data is needed
          float rv[32];
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int yy = t.y0 + (c >> 3) + g;
            const __nv_bfloat16* rrow = resp + (pix_img + (long long)yy * p.W + t.x0) * p.res_ld + ch;
            const bool rok = resp != nullptr && ch_ok && yy < p.H;
#pragma unroll
            for (int jx = 0; jx < 8; jx++) rv[g * 8 + jx] = (rok && jx < nx) ? __bfloat162float(rrow[(long long)jx * p.res_ld]) : 0.f;
          }
          tmem_ld_wait();
          if (full_tile) {
            // interior tile: no per-pixel tests, pointers advance by constants (10 instructions per pixel)
            __nv_bfloat16* op = outp + (pix_img + (long long)(t.y0 + (c >> 3)) * p.W + t.x0) * p.out_ld;
            const long long row_skip = (long long)(p.W - 8) * p.out_ld;
#pragma unroll
            for (int g = 0; g < 4; g++) {
#pragma unroll
              for (int jx = 0; jx < 8; jx++) {
                float v = __uint_as_float(r[g * 8 + jx]) + bias_v;
                v = fmaxf(v, v * slope_eff) + rv[g * 8 + jx];
                *op = __float2bfloat16(v);
                op += p.out_ld;
                s1 += v;
                s2 += v * v;
              }
              op += row_skip;
            }
          } else {
#pragma unroll
            for (int g = 0; g < 4; g++) {
              const int yy = t.y0 + (c >> 3) + g;
              __nv_bfloat16* orow = outp + (pix_img + (long long)yy * p.W + t.x0) * p.out_ld;
              const bool ook = ch_ok && yy < p.H;
#pragma unroll
              for (int jx = 0; jx < 8; jx++) {
                if (ook && jx < nx) {
                  float v = __uint_as_float(r[g * 8 + jx]) + bias_v;
                  v = fmaxf(v, v * slope_eff) + rv[g * 8 + jx];
                  orow[(long long)jx * p.out_ld] = __float2bfloat16(v);
                  s1 += v;
                  s2 += v * v;
                }
              }
            }
          }
        }
        csum[0] += s1;
        csq[0] += s2;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        continue;
      }
      if (SPLITK) {
        long long tp = TIMED ? clock64() : 0;
        auto phase_mark = [&](int slot) {
          if (TIMED && threadIdx.x == 64) {
            const long long now = clock64();
            p.dbg[(size_t)blockIdx.x * 16 + slot] += now - tp;
            tp = now;
          }
        };
        // ---- split-K: the `splits` CTAs of a tile are one thread-block cluster (cluster rank = split).  Every CTA sends the
        //      rows of its fp32 partial accumulator straight into the shared memory of the CTA that owns them
        //      (st.shared::cluster into the idle operand ring), the cluster synchronises, and every CTA reduces its row
        //      slice in split order (deterministic) and runs the real epilogue on it.  (Round 1 parked the partials in an L2
        //      workspace behind arrival counters: park 2.1-3.8 k + fence 1.6-2.1 k + wait 2.6-3.1 k + bulk fetch 5.5-7.9 k
        //      clocks per CTA, more than the weight-bound main loop.)  The host only enables this on the vector-store layout.
        // Receive layout in the owner, in float4 units: [sending split][column quad][row of the slice]: a warp's 32 rows
        // store runs of consecutive 16-byte words.
        const int bn4 = p.BN >> 2;
        const int my_s = ((row + 1) * p.splits + 127) / 128 - 1;                   // owner of this thread's row
        const int my_r0 = 128 * my_s / p.splits, my_n = 128 * (my_s + 1) / p.splits - my_r0;
        // 1) every CTA of the cluster is past its main loop (this CTA: tmem_full above): all operand rings are idle
        cluster_arrive();
        cluster_wait();
        phase_mark(12);
        const uint32_t dst0 = mapa_shared(smem0, (uint32_t)my_s) +
                              16u * (uint32_t)(split * my_n * bn4 + (row - my_r0));
        for (int c = member * 32; c < p.BN; c += 64) {
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          if (!valid) continue;                       // padding rows of an edge tile are never read back
#pragma unroll
          for (int i = 0; i < 8; i++)
            st_cluster_f4(dst0 + 16u * (uint32_t)(((c >> 2) + i) * my_n), __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                          __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);  // TMEM buffer is free again
        phase_mark(11);                                // send
        // 2) all partial rows have landed (release / acquire at cluster scope)
        cluster_arrive();
        cluster_wait();
        phase_mark(13);
        const int r0 = 128 * split / p.splits, r1 = 128 * (split + 1) / p.splits;
        // B1: sum this CTA's row slice over the splits in split order (deterministic) into `red`
        const int nrows = r1 - r0;
        const int slice_f4 = nrows * bn4;
        const uint32_t slice_bytes = (uint32_t)slice_f4 * 16u;
        const float4* stage4 = reinterpret_cast<const float4*>(smem);
        float* red = reinterpret_cast<float*>(smem + (size_t)p.splits * slice_bytes);
        // `red` rows are padded by one float4: threads that walk ROWS then hit 8 different bank groups (the unpadded row
        // stride of BN floats put a whole warp on four banks: the reduce and the epilogue below cost 5..15 k clocks)
        const int red_ld = p.BN + 4;
        {
          // thread = (row rr, column quads c4, c4 + cstep, ...): one division per thread, stage reads of a warp contiguous
          const int rr = etid % nrows, cstep = 256 / nrows;
          const int rw = r0 + rr, yy = t.y0 + rw / p.TW, xx = t.x0 + rw % p.TW;
          const bool rvalid = !(rw >= p.TH * p.TW || yy >= p.H || xx >= p.W);     // padding row: nothing sent, nothing stored
          for (int c4 = etid / nrows; c4 < bn4 && cstep > 0; c4 += cstep) {
            if (!rvalid || etid >= cstep * nrows) break;
            float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int u = 0; u < p.splits; u++) {
              const float4 f = stage4[(size_t)u * slice_f4 + c4 * nrows + rr];
              acc4.x += f.x; acc4.y += f.y; acc4.z += f.z; acc4.w += f.w;
            }
            *reinterpret_cast<float4*>(red + (size_t)rr * red_ld + 4 * c4) = acc4;
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        phase_mark(14);                                // B1: reduce over splits
        // B2: (row, 32-column chunk) tasks run the real epilogue from shared memory; consecutive threads take consecutive rows
        const int nch = out_cols_tile >> 5;
        const int tasks = nrows * nch;
        for (int task = etid; task < tasks; task += 256) {
          const int ch_i = task / nrows, rl = task - ch_i * nrows, c = ch_i << 5;
          const int rr = r0 + rl;
          const int yy = t.y0 + rr / p.TW, xx = t.x0 + rr % p.TW;
          const bool vld = (rr < p.TH * p.TW) && (yy < p.H) && (xx < p.W);
          float a[32], b[32];
          const float4* ra = reinterpret_cast<const float4*>(red + (size_t)rl * red_ld + c);
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 f = ra[i];
            a[4 * i] = f.x; a[4 * i + 1] = f.y; a[4 * i + 2] = f.z; a[4 * i + 3] = f.w;
          }
          if (pair) {
            const float4* rb = reinterpret_cast<const float4*>(red + (size_t)rl * red_ld + half + c);
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const float4 f = rb[i];
              b[4 * i] = f.x; b[4 * i + 1] = f.y; b[4 * i + 2] = f.z; b[4 * i + 3] = f.w;
            }
          }
          const long long pix2 = ((long long)t.img * p.H + yy) * p.W + xx;
          if (pair)
            epi_finish32<true>(p, a, b, &bias_s[acc][c], &bias_s[acc][half + c], vld, t.img, yy, xx, pix2, o0 + c, eflags);
          else
            epi_finish32<false>(p, a, a, &bias_s[acc][c], &bias_s[acc][c], vld, t.img, yy, xx, pix2, o0 + c, eflags);
          if (p.gn_acc != nullptr) {
            // fused GroupNorm statistics: park the finished values (zeros for padding rows) where the accumulators were
            float4* wr = reinterpret_cast<float4*>(red + (size_t)rl * red_ld + c);
#pragma unroll
            for (int i = 0; i < 8; i++)
              wr[i] = vld ? make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (p.gn_acc != nullptr) {
          // one thread per output column sums this CTA's row slice in a fixed order, then one fp64 atomic pair
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int cpg = p.cout / p.gn_groups;
          for (int col = etid; col < out_cols_tile; col += 256) {
            float s1 = 0.f, s2 = 0.f;                   // <= 43 rows: fp32 is plenty, the cross-CTA sum is fp64
            for (int rr = 0; rr < nrows; rr++) {
              const float xv = red[(size_t)rr * red_ld + col];
              s1 += xv;
              s2 += xv * xv;
            }
            double* dst = p.gn_acc + ((size_t)t.img * p.gn_groups + (o0 + col) / cpg) * 2;
            atomicAdd(dst, (double)s1);
            atomicAdd(dst + 1, (double)s2);
          }
        }
        phase_mark(15);                                // B2: epilogue + statistics
        continue;
      }
      if (fast) {
        const bool want_stats = !pair && p.gn_acc != nullptr;
        if (want_stats && (t.img != st_img || t.n_tile != st_nt)) {
          if (st_img >= 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const int cl = member * 32 + k * 64;
              if (cl < out_cols_tile && (lane & (32 / p.gn_ng - 1)) == 0) {
                const int g = ((st_nt * out_cols_tile + cl) >> 5) * p.gn_ng + lane / (32 / p.gn_ng);
                double* dst = p.gn_acc + ((size_t)st_img * p.gn_groups + g) * 2;
                atomicAdd(dst, (double)csum[k]);
                atomicAdd(dst + 1, (double)csq[k]);
              }
              csum[k] = csq[k] = 0.f;
            }
          }
          st_img = t.img;
          st_nt = t.n_tile;
        }
        // (Issuing the tcgen05.ld of the next chunk before the current one is finished -- a tensor-memory load takes ~700
        // clocks under a running main loop, a third of the epilogue of the small-K layers -- was tried in round 2: the second
        // register set makes ptxas reschedule the whole kernel and EVERY layer got 10..50 % slower, also the tilings that do
        // not run this loop.)
        for (int c = member * 32; c < out_cols_tile; c += 64) {
          float a[32], b[32];
          {
            uint32_t r[32];
            tmem_ld32(taddr + c, r);
            if (pair) {
              uint32_t r2[32];
              tmem_ld32(taddr + half + c, r2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i++) b[i] = __uint_as_float(r2[i]);
            } else {
              tmem_ld_wait();
            }
#pragma unroll
            for (int i = 0; i < 32; i++) a[i] = __uint_as_float(r[i]);
          }
          if (pair)
            epi_finish32<true>(p, a, b, &bias_s[acc][c], &bias_s[acc][half + c], valid, t.img, y, x, pix, o0 + c, eflags, stg_w, lane);
          else
            epi_finish32<false>(p, a, a, &bias_s[acc][c], &bias_s[acc][c], valid, t.img, y, x, pix, o0 + c, eflags, stg_w, lane);
          if (want_stats) {
            // GroupNorm statistics of what was just stored: per-group sums over this warp's 32 rows, kept in
            // registers across the CTA's tiles, flushed with one fp64 atomic per group when the tile column changes
            float s1, s2;
            if (p.gn_ng == 32) chunk_group_stats<32>(a, valid, lane, &s1, &s2);      // per channel
            else if (p.gn_ng == 8) chunk_group_stats<8>(a, valid, lane, &s1, &s2);
            else if (p.gn_ng == 4) chunk_group_stats<4>(a, valid, lane, &s1, &s2);
            else if (p.gn_ng == 2) chunk_group_stats<2>(a, valid, lane, &s1, &s2);
            else chunk_group_stats<1>(a, valid, lane, &s1, &s2);
            const int k = (c - member * 32) >> 6;
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
              if (kk == k) {
                csum[kk] += s1;
                csq[kk] += s2;
              }
          }
        }
      } else if (!pair) {
        // generic path (partial N tiles, transposed / unaligned stores): 16 columns at a time
        for (int c = member * 16; c < p.BN; c += 32) {
          uint32_t r[16];
          tmem_ld16(taddr + c, r);
          tmem_ld_wait();
          if (valid) {
            float a[16];
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = __uint_as_float(r[j]);
            epilogue16(p, t.img, y, x, n0 + c, 0, n0 + c, a, a);
          }
        }
      } else {
        for (int c = member * 16; c < half; c += 32) {
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr + c, r0);
          tmem_ld16(taddr + half + c, r1);
          tmem_ld_wait();
          if (valid) {
            float a[16], b[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
              a[j] = __uint_as_float(r0[j]);
              b[j] = __uint_as_float(r1[j]);
            }
            epilogue16(p, t.img, y, x, n0 + c, n0 + half + c, t.n_tile * half + c, a, b);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (TIMED && threadIdx.x == 64) p.dbg[(size_t)blockIdx.x * 16 + 8] = clock64() - t_epi0;
    if (!SPLITK && p.transposed) {
      if (p.gn_acc != nullptr && st_img >= 0) flush_channel_stats(p, st_img, q * 32 + lane, q * 32 + lane < p.cout, csum[0], csq[0]);
    } else if (p.gn_acc != nullptr && st_img >= 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int cl = member * 32 + k * 64;
        if (cl < out_cols_tile && (lane & (32 / p.gn_ng - 1)) == 0) {
          const int g = ((st_nt * out_cols_tile + cl) >> 5) * p.gn_ng + lane / (32 / p.gn_ng);
          double* dst = p.gn_acc + ((size_t)st_img * p.gn_groups + g) * 2;
          atomicAdd(dst, (double)csum[k]);
          atomicAdd(dst + 1, (double)csq[k]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT checking kernel: one thread = one pixel x 16 output columns.  Debug / test use only.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void simt_dot16(const IgemmParams& p, int img, int y, int x, int gcol, float* acc) {
#pragma unroll
  for (int j = 0; j < 16; j++) acc[j] = 0.f;
  const int taps = p.taps;
  for (int tap = 0; tap < taps; tap++) {
    int iy, ix;
    if (p.stride == 2) {
      iy = y * 2 + tap / 3 - 1;
      ix = x * 2 + tap % 3 - 1;
    } else {
      iy = y + p.tap_dy[tap];
      ix = x + p.tap_dx[tap];
    }
    if (iy < 0 || iy >= p.h_in || ix < 0 || ix >= p.w_in) continue;
    const long long ipix = ((long long)img * p.h_in + iy) * p.w_in + ix;
    const __nv_bfloat16* wz = p.w_ptr + (long long)(p.w_batched ? img : tap) * p.w_z;
    int kbase = 0;
    for (int src = 0; src < 2; src++) {
      const int C = p.a_c[src];
      if (C == 0) continue;
      const __nv_bfloat16* ap = p.a_ptr[src] + ipix * p.a_pix[src];
      for (int c = 0; c < C; c++) {
        const float av = __bfloat162float(ap[c]);
#pragma unroll
        for (int j = 0; j < 16; j++) {
          if (gcol + j < p.cout) acc[j] += av * __bfloat162float(wz[(long long)(gcol + j) * p.w_row + kbase + c]);
        }
      }
      kbase += C;
    }
  }
}

__global__ void igemm_simt_kernel(const __grid_constant__ IgemmParams p) {
  pdl_wait();
  const int groups = (p.ncols_out + 15) / 16;
  const long long total = (long long)p.n_img * p.H * p.W * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long pix = i / groups;
    const int x = (int)(pix % p.W);
    pix /= p.W;
    const int y = (int)(pix % p.H);
    const int img = (int)(pix / p.H);
    const int ocol = g * 16;
    float a[16], b[16];
    if (p.epi_mode == EPI_PLAIN) {
      simt_dot16(p, img, y, x, ocol, a);
      epilogue16(p, img, y, x, ocol, 0, ocol, a, a);
    } else {
      const int half = p.BN >> 1;
      const int nt = ocol / half, c = ocol - nt * half;
      const int ga = nt * p.BN + c, gb = ga + half;
      simt_dot16(p, img, y, x, ga, a);
      simt_dot16(p, img, y, x, gb, b);
      epilogue16(p, img, y, x, ga, gb, ocol, a, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &f, 12000, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  ONEDC_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[4];
  cuuint32_t b[5];
  for (int i = 0; i < rank; i++) {
    d[i] = dims[i];
    b[i] = box[i];
  }
  for (int i = 0; i + 1 < rank; i++) s[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, s, b, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] strides [%llu %llu %llu %llu] "
              "box [%u %u %u %u %u] ptr %p",
              (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
              (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0),
              (unsigned long long)(rank > 4 ? d[4] : 0), (unsigned long long)s[0],
              (unsigned long long)(rank > 2 ? s[1] : 0), (unsigned long long)(rank > 3 ? s[2] : 0),
              (unsigned long long)(rank > 4 ? s[3] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0,
              rank > 3 ? b[3] : 0, rank > 4 ? b[4] : 0, ptr);
    return -3;
  }
  return 0;
}

static int pick_bn(int cout, bool pair) {
  (void)pair;
  // the vector epilogue works on 32-column chunks: prefer N tiles that are multiples of 32
  if (cout <= 256) return cout % 32 == 0 ? cout : ((cout + 15) / 16) * 16;
  int best = 256, best_waste = 1 << 30;
  for (int bn = 256; bn >= 96; bn -= 32) {
    int nt = (cout + bn - 1) / bn;
    int waste = nt * bn - cout;
    if (waste < best_waste) {
      best_waste = waste;
      best = bn;
    }
  }
  return best;
}

static void pick_tile(int H, int W, int* th, int* tw) {
  if (H == 1) {
    *th = 1;
    *tw = 128;
    return;
  }
  // fewest tiles wins; ties go to the most square box (best halo reuse in L2 for 3x3 taps)
  const int cand[7] = {16, 8, 32, 4, 64, 2, 128};
  long long best = -1;
  for (int i = 0; i < 7; i++) {
    const int t = cand[i], h = 128 / t;
    const long long tiles = (long long)((H + h - 1) / h) * ((W + t - 1) / t);
    if (best < 0 || tiles < best) {
      best = tiles;
      *th = h;
      *tw = t;
    }
  }
}

static long long* g_igemm_dbg = nullptr;

// Shortest K loop (in 64-channel x tap iterations) that is split.  A cluster launch plus the exchange costs ~8 us; measured
// on B200 (round 2, L2-flushed single launches, split vs plain): 48x48 512->512 3x3 (72 iterations) 24.6 vs 18.4 us and
// 576 x 5120->1280 (80) 22.5 vs 20.5 us -- plain wins; 24x24 1280->1280 3x3 (180) 32.8 vs 34.8, 12x12 1280->1280 3x3 (180)
// 20.5 vs 34.8, 24x24 2560->1280 3x3 (360) 52.2 vs 59.4 -- split wins.
static int split_min_kiters() {
  static const char* e = getenv("ONEDC_SPLITK_MIN_KITERS");
  return e != nullptr ? atoi(e) : 100;
}

// clusters of `size` split-K CTAs (one CTA per SM: 200 KB of shared memory) that the device can hold at once
static int max_active_clusters(int size) {
  static int cache[17] = {0};
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (size < 1 || size > 16) return 0;
  if (cache[size] == 0) {
    cudaFuncSetAttribute(igemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(size * 64));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBudget + 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, igemm_tc_kernel<true, false>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = sm_count() / (size < 4 ? size : (size <= 8 ? 9 : 18));      // conservative guess: 16-SM GPCs
    }
    cache[size] = n > 0 ? n : -1;
  }
  return cache[size] > 0 ? cache[size] : 0;
}

static int igemm_launch(onedc_igemm_desc* d, cudaStream_t stream) {
  ONEDC_CHECK(d->ksize == 1 || d->ksize == 3, "igemm: ksize must be 1 or 3");
  ONEDC_CHECK(d->stride == 1 || (d->stride == 2 && d->ksize == 3), "igemm: stride 2 needs ksize 3");
  ONEDC_CHECK(d->a_c[0] > 0 && d->a_c[0] % 8 == 0 && d->a_c[1] % 8 == 0, "igemm: channels must be multiples of 8");
  ONEDC_CHECK(d->a_pix_stride[0] % 8 == 0 && (d->a_c[1] == 0 || d->a_pix_stride[1] % 8 == 0),
              "igemm: pixel strides must be multiples of 8 elements");
  ONEDC_CHECK(d->ktot >= d->a_c[0] + d->a_c[1] && d->w_row_stride % 8 == 0 && d->w_z_stride % 8 == 0,
              "igemm: bad weight strides");
  const bool pair = d->epi_mode != EPI_PLAIN;
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = d->n_img;
  p.h_in = d->h_in;
  p.w_in = d->w_in;
  p.ksize = d->ksize;
  p.stride = d->stride;
  if (d->stride == 2) {
    ONEDC_CHECK(d->h_in % 2 == 0 && d->w_in % 2 == 0 && d->a_c[0] % 64 == 0 && d->a_c[1] == 0,
                "igemm: stride-2 conv needs even dims, C %% 64 == 0 and a single source");
    p.H = d->h_in / 2;
    p.W = d->w_in / 2;
  } else {
    p.H = d->h_in;
    p.W = d->w_in;
  }
  p.cout = d->cout;
  p.BN = d->bn > 0 ? d->bn : pick_bn(d->cout, pair);
  ONEDC_CHECK(p.BN >= 16 && p.BN <= 256 && p.BN % (pair ? 32 : 16) == 0, "igemm: bad BN %d", p.BN);
  p.n_tiles = (d->cout + p.BN - 1) / p.BN;
  if (d->bn <= 0 && !pair && d->impl != 1) {
    // few tiles and a short K loop (split-K will not apply): halve the N tile while that still fits one wave,
    // so twice as many SMs share the MMA work (A tiles are re-read from L2, which is cheap at these sizes)
    int th0, tw0;
    const int Ho = d->stride == 2 ? d->h_in / 2 : d->h_in, Wo = d->stride == 2 ? d->w_in / 2 : d->w_in;
    pick_tile(Ho, Wo, &th0, &tw0);
    const int mt = d->n_img * ((Ho + th0 - 1) / th0) * ((Wo + tw0 - 1) / tw0);
    const int taps0 = d->ntaps > 0 ? d->ntaps : d->ksize * d->ksize;
    const int kit = taps0 * ((d->a_c[0] + 63) / 64 + (d->a_c[1] + 63) / 64);
    const bool will_split = !d->deterministic && d->splitk_ws != nullptr && mt * p.n_tiles * 2 <= sm_count() && kit >= split_min_kiters() &&
                            sm_count() / (mt * p.n_tiles) >= 4;
    while (!will_split && p.BN % 64 == 0 && p.BN >= 128 && d->cout % (p.BN / 2) == 0 &&
           mt * ((d->cout + p.BN / 2 - 1) / (p.BN / 2)) <= sm_count()) {
      p.BN /= 2;
      p.n_tiles = (d->cout + p.BN - 1) / p.BN;
    }
    // split-K runs at most 8 splits per tile (one cluster): with very few tiles that leaves SMs idle, so halve the N tile
    // first (same weight traffic, same parked bytes, twice the CTAs)
    if (will_split && p.BN == 256 && d->cout % 128 == 0 && mt * p.n_tiles * 8 * 2 <= sm_count() + 20) {
      p.BN = 128;
      p.n_tiles = (d->cout + p.BN - 1) / p.BN;
    }
  }
  if (pair) ONEDC_CHECK(d->cout % p.BN == 0, "igemm: pair epilogues need cout %% BN == 0 (cout %d BN %d)", d->cout, p.BN);
  pick_tile(p.H, p.W, &p.TH, &p.TW);
  p.tiles_y = (p.H + p.TH - 1) / p.TH;
  p.tiles_x = (p.W + p.TW - 1) / p.TW;
  p.m_tiles = p.n_img * p.tiles_y * p.tiles_x;
  p.taps = d->ntaps > 0 ? d->ntaps : d->ksize * d->ksize;
  ONEDC_CHECK(p.taps <= 9 && !(d->ntaps > 0 && d->stride != 1), "igemm: custom taps need stride 1 and at most 9 taps");
  p.quad = d->quad;
  p.kchunks[0] = (d->a_c[0] + 63) / 64;
  p.kchunks[1] = (d->a_c[1] + 63) / 64;
  p.c1_off = d->a_c[0];
  p.w_batched = d->w_batched;
  ONEDC_CHECK(!(d->w_batched && p.taps != 1), "igemm: batched weights need a 1x1 kernel");
  for (int t = 0; t < p.taps; t++) {
    int ky = t / 3, kx = t % 3;
    if (d->ntaps > 0) {
      p.tap_dc[t] = 0;
      p.tap_dp[t] = 0;
      p.tap_dx[t] = d->tap_dx[t];
      p.tap_dy[t] = d->tap_dy[t];
    } else if (d->ksize == 1) {
      p.tap_dc[t] = p.tap_dx[t] = p.tap_dp[t] = p.tap_dy[t] = 0;
    } else if (d->stride == 1) {
      p.tap_dc[t] = 0;
      p.tap_dx[t] = kx - 1;
      p.tap_dp[t] = 0;
      p.tap_dy[t] = ky - 1;
    } else {
      const int px = (kx == 1) ? 0 : 1, py = (ky == 1) ? 0 : 1;
      p.tap_dc[t] = px * (int)d->a_pix_stride[0];
      p.tap_dx[t] = (kx == 0) ? -1 : 0;
      p.tap_dp[t] = py;
      p.tap_dy[t] = (ky == 0) ? -1 : 0;
    }
  }
  p.stage_bytes = kABytes + ((p.BN * 128 + 1023) / 1024) * 1024;
  p.stages = kSmemBudget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.bias = d->bias;
  p.epi_mode = d->epi_mode;
  p.act = d->act;
  p.slope = d->slope;
  p.res = d->res;
  p.res_dtype = d->res_dtype;
  p.res_ld = d->res_ld;
  p.out = d->out;
  p.out_dtype = d->out_dtype;
  p.out_ld = d->out_ld;
  p.out_col_off = d->out_col_off;
  p.store_mode = d->store_mode;
  p.ps_c = d->ps_c;
  p.ncols_out = pair ? d->cout / 2 : d->cout;
  if (d->store_mode == ST_PIXSHUF)
    ONEDC_CHECK(d->ps_c > 0 && d->ps_c % 16 == 0 && p.ncols_out == 4 * d->ps_c && d->res == nullptr,
                "igemm: pixel-shuffle store needs ncols == 4*ps_c, ps_c %% 16 == 0, no residual");
  {
    const int va = (d->out_dtype == DT_BF16) ? 8 : 4;
    bool ok = (d->out_ld % va == 0) && (d->out_col_off % va == 0) &&
              (reinterpret_cast<uintptr_t>(d->out) % 16 == 0) && d->store_mode != ST_TRANSPOSED;
    if (d->res != nullptr)
      ok = ok && (d->res_ld % 8 == 0) && (reinterpret_cast<uintptr_t>(d->res) % 16 == 0);
    p.vec_ok = ok ? 1 : 0;
  }
  p.a_ptr[0] = reinterpret_cast<const __nv_bfloat16*>(d->a_ptr[0]);
  p.a_ptr[1] = reinterpret_cast<const __nv_bfloat16*>(d->a_ptr[1]);
  p.a_c[0] = d->a_c[0];
  p.a_c[1] = d->a_c[1];
  p.a_pix[0] = d->a_pix_stride[0];
  p.a_pix[1] = d->a_pix_stride[1];
  p.w_ptr = reinterpret_cast<const __nv_bfloat16*>(d->w_ptr);
  p.ktot = d->ktot;
  p.w_row = d->w_row_stride;
  p.w_z = d->w_z_stride;

  // ---- split-K for layers with too few tiles to fill the GPU (UNet 24x24 / 12x12 levels): only on the
  //      vector-store fast path, with caller-provided scratch; each split keeps >= 8 k-iterations
  p.splits = 1;
  p.ws = reinterpret_cast<float*>(d->splitk_ws);
  p.counters = reinterpret_cast<int*>(d->splitk_counters);
  p.counters_half = d->splitk_max_tiles / 2;
  {
    const int tiles = p.m_tiles * p.n_tiles;
    const int kiters = p.taps * (p.kchunks[0] + p.kchunks[1]);
    const int octile = pair ? p.BN / 2 : p.BN;
    const bool fast_all = p.vec_ok && (d->cout % p.BN == 0) && (octile % 32 == 0) &&
                          (d->store_mode == ST_NORMAL || d->store_mode == ST_QUAD || (d->store_mode == ST_PIXSHUF && d->ps_c % 32 == 0));
    if (d->impl != 1 && !d->deterministic && p.ws != nullptr && p.counters != nullptr && fast_all && tiles * 2 <= sm_count() &&
        kiters >= split_min_kiters() && tiles <= d->splitk_max_tiles / 2) {
      int s = sm_count() / tiles;
      if (s > kiters / 8) s = kiters / 8;
      // The splits of a tile wait for each other (arrival counter), so they are launched as one thread-block cluster:
      // co-residency is then guaranteed by the hardware.  Without it two split-K kernels of different streams (the
      // pipelined decoder runs three) could each occupy part of the GPU and spin for CTAs that never get an SM.
      // 8 is the portable cluster limit.  A cluster lives inside one GPC (16..20 SMs, one CTA per SM here), so not every
      // size packs the GPU: the largest split count whose clusters are all resident at once wins (two waves of 7-CTA
      // clusters cost the 12 x 12 layers +45 %); if none fits, the one with the most resident CTAs.
      if (s > 8) s = 8;
      while (s > 1 && (long long)tiles * s * 128 * p.BN > d->splitk_ws_floats) s--;
      if (s >= 4) {
        int best = 0, best_ctas = 0;
        for (int c = s; c >= 3; c--) {
          const int cap = max_active_clusters(c);
          if (cap >= tiles) {
            best = c;
            break;
          }
          if (cap * c > best_ctas) {
            best_ctas = cap * c;
            best = c;
          }
        }
        if (best >= 3) p.splits = best;
      }
    }
  }

  // ---- column-copy mode: stride-1 multi-tap convs that do not split K (see IgemmParams)
  p.colmode = 0;
  {
    static const bool colmode_on = getenv("ONEDC_COLMODE") == nullptr || getenv("ONEDC_COLMODE")[0] != '0';
    if (colmode_on && d->impl != 1 && !d->deterministic && p.splits == 1 && d->stride == 1 && p.taps > 1 && !d->w_batched) {
      int miny = 99, maxy = -99, ng = 0;
      bool ok = true;
      memset(p.cm_nt, 0, sizeof(p.cm_nt));
      for (int t = 0; t < p.taps; t++) {
        miny = p.tap_dy[t] < miny ? p.tap_dy[t] : miny;
        maxy = p.tap_dy[t] > maxy ? p.tap_dy[t] : maxy;
      }
      for (int t = 0; t < p.taps && ok; t++) {
        int g = 0;
        while (g < ng && p.cm_dx[g] != p.tap_dx[t]) g++;
        if (g == ng) {
          if (ng == 3) {
            ok = false;
            break;
          }
          p.cm_dx[ng++] = p.tap_dx[t];
        }
        if (p.cm_nt[g] == 3) {
          ok = false;
          break;
        }
        p.cm_tap[g][p.cm_nt[g]] = t;
        p.cm_row[g][p.cm_nt[g]] = p.tap_dy[t] - miny;
        p.cm_nt[g]++;
      }
      if (ok && maxy - miny <= 2) {
        p.cm_groups = ng;
        p.cm_miny = miny;
        p.cm_rows = 16 + maxy - miny;
        p.a_slot_bytes = p.cm_rows * 1024;
        p.b_slot_bytes = ((p.BN * 128 + 1023) / 1024) * 1024;
        p.a_slots = p.BN <= 128 ? 4 : 3;
        p.b_slots = (kSmemBudget - p.a_slots * p.a_slot_bytes) / p.b_slot_bytes;
        if (p.b_slots > kMaxStages) p.b_slots = kMaxStages;
        // measured on B200: wins 4-6 % when a CTA runs several tiles back to back, loses on single-wave layers
        const int tiles16x8 = p.n_img * ((p.H + 15) / 16) * ((p.W + 7) / 8) * p.n_tiles;
        static const bool colmode_all = getenv("ONEDC_COLMODE") != nullptr && getenv("ONEDC_COLMODE")[0] == '2';
        if (p.b_slots >= 3 && (colmode_all || tiles16x8 > sm_count())) {
          p.colmode = 1;
          p.TH = 16;
          p.TW = 8;
          // transposed variant for narrow outputs: one 128-channel M tile, N = 32 x 8 pixels
          static const bool transp_on = getenv("ONEDC_TRANSPOSED") == nullptr || getenv("ONEDC_TRANSPOSED")[0] != '0';
          const int tiles32x8 = p.n_img * ((p.H + 31) / 32) * ((p.W + 7) / 8);
          static const bool res_mma_on = getenv("ONEDC_RES_MMA") == nullptr || getenv("ONEDC_RES_MMA")[0] != '0';
          const bool res_mma = res_mma_on && d->res != nullptr && d->w_identity_tap && d->res_dtype == DT_BF16 && d->act == ACT_NONE &&
                               d->a_c[1] == 0 && d->ktot == d->cout && d->res_ld % 8 == 0 &&
                               reinterpret_cast<uintptr_t>(d->res) % 16 == 0;
          const int cpg_t = d->gn_acc != nullptr && d->gn_groups > 0 ? d->cout / d->gn_groups : 1;
          if (transp_on && !pair && d->cout <= 128 && d->cout >= 64 && p.n_tiles == 1 && d->store_mode == ST_NORMAL &&
              d->out_dtype == DT_BF16 && (d->res == nullptr || res_mma || (colmode_all && d->res_dtype == DT_BF16)) &&
              // a residual read in the epilogue costs 2-byte loads per lane and pixel and makes the epilogue the bottleneck
              // (208 us against 141 us for 128 -> 128 at 768 x 768): it goes through the tensor core instead (res_mma)
              (d->act == ACT_NONE || (d->act == ACT_LRELU && d->slope > 0.f && d->slope <= 1.f)) &&
              (colmode_all || tiles32x8 > sm_count()) && (cpg_t & (cpg_t - 1)) == 0 && cpg_t <= 32 &&
              (d->gn_acc == nullptr || d->cout % d->gn_groups == 0)) {
            p.transposed = 1;
            if (res_mma) {
              p.res_chunks = (d->cout + 63) / 64;
              p.res = nullptr;                            // nothing left for the epilogue to add
            }
            p.TH = 32;
            p.BN = 128;                                   // weight rows per tile (TMA box; rows >= cout are zero-filled)
            p.cm_rows = 32 + maxy - miny;
            p.a_slot_bytes = p.cm_rows * 1024;
            p.b_slot_bytes = 128 * 128;
            p.a_slots = 3;
            p.b_slots = (kSmemBudget - p.a_slots * p.a_slot_bytes) / p.b_slot_bytes;
            if (p.b_slots > kMaxStages) p.b_slots = kMaxStages;
          }
          p.tiles_y = (p.H + p.TH - 1) / p.TH;
          p.tiles_x = (p.W + p.TW - 1) / p.TW;
          p.m_tiles = p.n_img * p.tiles_y * p.tiles_x;
        }
      }
    }
  }

  p.dbg = g_igemm_dbg;
  p.pf_ptr = nullptr;
  p.pf_bytes = 0;
  if (d->prefetch_ptr != nullptr && d->prefetch_bytes >= 16 && reinterpret_cast<uintptr_t>(d->prefetch_ptr) % 16 == 0) {
    p.pf_ptr = reinterpret_cast<const uint8_t*>(d->prefetch_ptr);
    p.pf_bytes = d->prefetch_bytes & ~15ll;
    if (p.pf_bytes > (64ll << 20)) p.pf_bytes = 64ll << 20;          // half of the L2 at most
  }
  p.gn_acc = nullptr;
  d->gn_fused_out = 0;
  if (d->gn_acc != nullptr && d->impl != 1 && !pair) {
    const bool fast_all = p.vec_ok && (d->cout % p.BN == 0) && (p.BN % 32 == 0) &&
                          (d->store_mode == ST_NORMAL || d->store_mode == ST_QUAD) && p.BN <= 256;
    const int cpg = d->gn_groups > 0 ? d->cout / d->gn_groups : 0;
    const bool cpg_ok = p.transposed ? (cpg >= 1 && cpg <= 32 && (cpg & (cpg - 1)) == 0)
                                     : (cpg == 1 || cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32);
    if ((fast_all || p.transposed) && d->gn_groups > 0 && d->cout % d->gn_groups == 0 && cpg_ok) {
      p.gn_acc = reinterpret_cast<double*>(d->gn_acc);
      p.gn_ng = cpg <= 32 ? 32 / cpg : 1;
      p.gn_groups = d->gn_groups;
      d->gn_fused_out = 1;
    }
  }

  if (d->impl == 1) {
    long long total = (long long)p.n_img * p.H * p.W * ((p.ncols_out + 15) / 16);
    int blocks = (int)((total + 127) / 128);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    ONEDC_CUDA(launch_k(igemm_simt_kernel, blocks, 128, 0, stream, p));
    ONEDC_CUDA(cudaGetLastError());
    return 0;
  }

  if (d->impl == 2) return 0;          // dry run (bench): decisions only

  // ---- tensor maps
  CUtensorMap ma[2], mb;
  memset(ma, 0, sizeof(ma));
  for (int s = 0; s < 2; s++) {
    const bool res_map = s == 1 && p.res_chunks > 0;      // the residual as a second activation source (stride 1, cout channels)
    if (d->a_c[s] == 0 && !res_map) {
      ma[s] = ma[0];
      continue;
    }
    const void* a_base = res_map ? d->res : d->a_ptr[s];
    const int a_ch = res_map ? d->cout : d->a_c[s];
    ONEDC_CHECK(reinterpret_cast<uintptr_t>(a_base) % 16 == 0, "igemm: A pointer must be 16-byte aligned");
    const uint64_t S = (uint64_t)(res_map ? d->res_ld : d->a_pix_stride[s]);
    uint64_t dims[5], str[4];
    uint32_t box[5] = {64, (uint32_t)p.TW, 1, (uint32_t)(p.colmode ? p.cm_rows : p.TH), 1};
    if (d->stride == 1) {
      dims[0] = (uint64_t)a_ch;
      dims[1] = (uint64_t)d->w_in;
      dims[2] = 1;
      dims[3] = (uint64_t)d->h_in;
      dims[4] = (uint64_t)d->n_img;
      str[0] = S * 2;
      str[1] = (uint64_t)d->w_in * S * 2;
      str[2] = (uint64_t)d->w_in * S * 2;
      str[3] = (uint64_t)d->h_in * d->w_in * S * 2;
    } else {
      dims[0] = S + (uint64_t)a_ch;
      dims[1] = (uint64_t)d->w_in / 2;
      dims[2] = 2;
      dims[3] = (uint64_t)d->h_in / 2;
      dims[4] = (uint64_t)d->n_img;
      str[0] = 2 * S * 2;
      str[1] = (uint64_t)d->w_in * S * 2;
      str[2] = 2 * (uint64_t)d->w_in * S * 2;
      str[3] = (uint64_t)d->h_in * d->w_in * S * 2;
    }
    int rc = make_tensor_map(&ma[s], a_base, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    ONEDC_CHECK(reinterpret_cast<uintptr_t>(d->w_ptr) % 16 == 0, "igemm: W pointer must be 16-byte aligned");
    const int nz = d->w_batched ? d->n_img : p.taps + (d->w_identity_tap ? 1 : 0);
    uint64_t dims[3] = {(uint64_t)d->ktot, (uint64_t)d->cout, (uint64_t)nz};
    uint64_t str[2] = {(uint64_t)d->w_row_stride * 2, (uint64_t)d->w_z_stride * 2};
    if (nz == 1) str[1] = str[0] * dims[1];
    uint32_t box[3] = {64, (uint32_t)p.BN, 1};
    int rc = make_tensor_map(&mb, d->w_ptr, 3, dims, str, box);
    if (rc) return rc;
  }
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  int grid = total_tiles < sm_count() ? total_tiles : sm_count();
  const size_t smem = (p.colmode ? (size_t)p.a_slots * p.a_slot_bytes + (size_t)p.b_slots * p.b_slot_bytes
                                 : (size_t)p.stages * p.stage_bytes) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    ONEDC_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget + 1024));
    ONEDC_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget + 1024));
    ONEDC_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget + 1024));
    ONEDC_CUDA(cudaFuncSetAttribute(igemm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBudget + 1024));
    attr_set = true;
  }
  if (p.splits > 1) ONEDC_CHECK(grid == total_tiles && grid % p.splits == 0, "igemm: split-K grid must hold whole clusters");
  if (p.splits > 1 && p.dbg != nullptr)
    ONEDC_CUDA(launch_cluster_k(igemm_tc_kernel<true, true>, grid, kThreads, smem, stream, (unsigned)p.splits, ma[0], ma[1], mb, p));
  else if (p.splits > 1)
    ONEDC_CUDA(launch_cluster_k(igemm_tc_kernel<true, false>, grid, kThreads, smem, stream, (unsigned)p.splits, ma[0], ma[1], mb, p));
  else if (p.dbg != nullptr)
    ONEDC_CUDA(launch_k(igemm_tc_kernel<false, true>, grid, kThreads, smem, stream, ma[0], ma[1], mb, p));
  else
    ONEDC_CUDA(launch_k(igemm_tc_kernel<false, false>, grid, kThreads, smem, stream, ma[0], ma[1], mb, p));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace onedc

extern "C" void onedc_igemm_set_debug(void* dev_counters) { onedc::g_igemm_dbg = reinterpret_cast<long long*>(dev_counters); }

extern "C" int onedc_igemm(onedc_igemm_desc* d, void* stream) {
  return onedc::igemm_launch(d, reinterpret_cast<cudaStream_t>(stream));
}
