// Flash attention for the UNet's self/cross attention on tcgen05 (sm_100a).
//
//   O[b, s, h*d : (h+1)*d] = softmax(Q K^T * scale) V      per (batch b, head h), non-causal
//
// One CTA = 128 queries of one (b, h).  Q/K/V are read straight out of the [b, s, heads*d] projection
// outputs through 4-D tensor maps {d, s, head, b}: the box is 64 channels wide, so for d = 40/80/160 the
// columns beyond d are out of bounds and TMA zero-fills them (no padded copies, no head split kernel).
//   S = Q K^T     : tcgen05.mma, A = Q (K-major, smem), B = K tile (K-major, smem)  -> TMEM (2 buffers of 64 columns)
//   softmax       : 4 warps, thread = query row, S row in registers; exp2 against a LAZY reference maximum;
//                   P (bf16) is written back into TMEM over the first 32 columns of the S buffer it came from
//   O += P V      : tcgen05.mma with the A operand (P) read from TMEM, B = V tile (MN-major, smem as loaded by TMA)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax + output.  QK^T of block j+1 is issued
// before P V of block j, so the tensor core works on the next scores while softmax runs; two CTAs per SM (TMEM 256
// columns each) overlap one CTA's softmax with the other's.
//
// What bounds this kernel is the softmax warps, not the tensor pipe (a 64-key block costs ~100 MMA clocks and 8192
// exponentials = 512 clocks of the SM's 16/clk MUFU).  Round 1 spent 2200 clocks per block (ncu: MUFU pipe 58 %, issue
// slots 47 %, the rest dependent-issue latency).  The loop now keeps off the critical path everything that is not
// max -> exp -> store:
//   * P never touches shared memory: no 16 KB of st.shared per block, no fence.proxy.async round trip (7 % of all
//     warp samples sat on that fence); tcgen05.st + wait::st instead, and the P V MMA takes A from TMEM.
//   * lazy rescale: the exponent reference m_ref moves only when the block maximum exceeds it by more than 8 (log2
//     units), so P <= 256 and O / l stay consistent without touching O; with the exact running maximum a warp rescaled
//     O in 80 of the 144 blocks of a 9216-key row (any of its 32 rows moving), each time waiting for the previous P V.
//   * packed fp32x2 arithmetic (fma.rn.f32x2 / add.f32x2) and 3-input max with four independent chains: ~225 instead of
//     ~356 warp instructions per block, no 64-long dependent chains.
// Launch plans (onedc_attention_set_plan): the key range may be split over 2..4 CTAs per query tile with an fp32 merge
// pass (attention_merge_kernel); measured slower for the UNet's shapes, kept selectable and tested.
//
// The SIMT kernel at the bottom is the on-GPU checker (impl = 1), never used by the decode path.
#include "../../include/onedc_b200.h"
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace onedc {

int make_tensor_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

constexpr int BKV = 64;        // keys per block
constexpr int kAttnThreads = 192;
constexpr float kLazyTau = 8.f;   // log2 units: P <= 2^8, far inside bf16 / fp32 range

struct AttnParams {
  int sq, skv, d, dk16, nchunk;      // dk16 = round_up(d,16), nchunk = ceil(d/64)
  int nblk;
  float scale_log2;
  __nv_bfloat16* out;
  long long o_ld;
  int heads;
  int tmem_cols;   // 256: two S buffers (2 x 64 columns) + O (dk16 <= 128 columns); else 512
  // key-range split (wave quantisation): CTA z = batch * kv_splits + split handles key blocks [split*bps, +bps) and,
  // when kv_splits > 1, leaves an unnormalised fp32 O and its (reference max, sum) for attention_merge_kernel
  int kv_splits, bps;
  float* ws_o;     // [kv_splits][batch][sq][heads*d] fp32
  float* ws_ml;    // [kv_splits][batch][heads][sq][2] fp32
};

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2) and 3-input max
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x (K/2) 32-bit columns, two bf16 per column (K-major)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full, pv_done;
  __shared__ uint32_t tmem_slot;

  // broadcast so that the compiler knows the warp index is warp-uniform (role branches stay uniform)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y;
  const int batch = blockIdx.z / p.kv_splits, split = blockIdx.z - batch * p.kv_splits;
  const int j0 = split * p.bps;                                        // first key block of this CTA
  const int nb = (p.nblk - j0 < p.bps) ? p.nblk - j0 : p.bps;         // its number of key blocks (>= 1, host-checked)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_bytes = p.nchunk * 16384;            // [chunk][128 rows][128 B]
  const int kv_bytes = p.nchunk * BKV * 128;       // [chunk][BKV rows][128 B]
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_bytes;                      // 2 stages
  uint8_t* sV = sK + 2 * kv_bytes;                 // 2 stages

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(&p_full, 4);       // one arrive per softmax warp
    mbar_init(&pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();
  const uint32_t tmem_o = tmem_base + 2 * BKV;      // S buffers at columns [0, 128), O after them

  // The TMA and MMA roles are single-lane jobs run by the WHOLE warp with only the TMA / MMA / commit instructions
  // under `if (leader)` and every address a register bumped by constants: the loop state stays in uniform registers
  // (as in igemm.cu).
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
  const uint32_t q_full_a = smem_u32(&q_full), k_full_a = smem_u32(&k_full[0]), k_empty_a = smem_u32(&k_empty[0]);
  const uint32_t v_full_a = smem_u32(&v_full[0]), v_empty_a = smem_u32(&v_empty[0]), s_full_a = smem_u32(&s_full[0]);
  const uint32_t p_full_a = smem_u32(&p_full), pv_done_a = smem_u32(&pv_done);
  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      mbar_expect_tx_a(q_full_a, (uint32_t)q_bytes);
      for (int c = 0; c < p.nchunk; c++) tma_load_4d_a(sQ_a + c * 16384, &map_q, q_full_a, c * 64, q0, head, batch);
    }
    __syncwarp();
    for (int j = 0; j < nb; j++) {
      const uint32_t st = j & 1, ph = (j >> 1) & 1;
      const int key0 = (j0 + j) * BKV;
      mbar_wait_a(k_empty_a + st * 8, ph ^ 1);
      if (leader) {
        mbar_expect_tx_a(k_full_a + st * 8, (uint32_t)kv_bytes);
        for (int c = 0; c < p.nchunk; c++)
          tma_load_4d_a(sK_a + st * kv_bytes + c * BKV * 128, &map_k, k_full_a + st * 8, c * 64, key0, head, batch);
      }
      __syncwarp();
      mbar_wait_a(v_empty_a + st * 8, ph ^ 1);
      if (leader) {
        mbar_expect_tx_a(v_full_a + st * 8, (uint32_t)kv_bytes);
        for (int c = 0; c < p.nchunk; c++)
          tma_load_4d_a(sV_a + st * kv_bytes + c * BKV * 128, &map_v, v_full_a + st * 8, c * 64, key0, head, batch);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc_qk = umma_idesc_bf16(128, BKV, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16(128, p.dk16, 0, 1);   // A = P from TMEM (K-major), B = V is MN-major
    // descriptor = constant high part | (shared address >> 4)
    const uint64_t dhi_k = umma_smem_desc(0, 16, 1024);              // K-major tiles: Q, K
    const uint64_t dhi_v = umma_smem_desc(0, BKV * 128, 1024);       // V: MN(d)-major, 64-wide d chunks BKV*128 B apart
    const uint32_t q_enc = (sQ_a & 0x3FFFF) >> 4, k_enc = (sK_a & 0x3FFFF) >> 4, v_enc = (sV_a & 0x3FFFF) >> 4;
    const uint32_t kv_enc = (uint32_t)kv_bytes >> 4;
    const int ksteps = p.dk16 / 16;
    mbar_wait_a(q_full_a, 0);
    for (int j = 0; j <= nb; j++) {
      if (j < nb) {
        // ---- S(j) = Q K(j)^T into S buffer j & 1.  That buffer's first 32 columns held P(j-2): P V(j-2) was issued
        //      before this MMA and the tensor pipe executes in issue order, so the overwrite is safe.
        const uint32_t st = j & 1;
        mbar_wait_a(k_full_a + st * 8, (j >> 1) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t kst = k_enc + st * kv_enc;
          for (int kk = 0; kk < ksteps; kk++) {
            const uint32_t c = kk >> 2, w = kk & 3;
            const uint64_t da = dhi_k | (uint64_t)(q_enc + c * (16384 >> 4) + w * 2);
            const uint64_t db = dhi_k | (uint64_t)(kst + c * (BKV * 128 >> 4) + w * 2);
            umma_bf16(tmem_base + st * BKV, da, db, idesc_qk, kk != 0);
          }
          umma_commit_a(s_full_a + st * 8);
          umma_commit_a(k_empty_a + st * 8);
        }
        __syncwarp();
      }
      if (j >= 1) {
        // ---- O += P(j-1) V(j-1)   (issued after Q K(j)^T so the tensor core works on the next scores meanwhile)
        const int jp = j - 1;
        const uint32_t st = jp & 1;
        mbar_wait_a(p_full_a, jp & 1);
        mbar_wait_a(v_full_a + st * 8, (jp >> 1) & 1);
        tc_fence_after();
        if (leader) {
          const uint32_t vst = v_enc + st * kv_enc;
          const uint32_t tmem_p = tmem_base + st * BKV;                // bf16 pairs: 8 columns per 16 keys
#pragma unroll
          for (int kk = 0; kk < BKV / 16; kk++) {
            const uint64_t db = dhi_v | (uint64_t)(vst + kk * (2048 >> 4));      // 8 kv rows per 1024-byte atom
            umma_bf16_ts(tmem_o, tmem_p + kk * 8, db, idesc_pv, (jp | kk) != 0);
          }
          umma_commit_a(pv_done_a);
          umma_commit_a(v_empty_a + st * 8);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- softmax / output warps -------------------------------
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.f;            // exponent reference (lazy maximum), running sum w.r.t. m_ref
    const uint64_t sc2 = f2_pack(p.scale_log2, p.scale_log2);
    for (int j = 0; j < nb; j++) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t t_s = tmem_base + lane_off + sb * BKV;
      uint32_t sr[BKV];
      tmem_ld32(t_s, sr);
      tmem_ld32(t_s + 32, sr + 32);
      tmem_ld_wait();
      const int nvalid = p.skv - (j0 + j) * BKV;     // columns >= nvalid are zero-filled padding keys
      if (nvalid < BKV) {                            // only the last block can be partial (uniform branch)
#pragma unroll
        for (int c = 0; c < BKV; c++)
          if (c >= nvalid) sr[c] = 0xff800000u;      // -inf
      }
      // block maximum: four independent chains of 3-input max
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < BKV; c += 8) {
#pragma unroll
        for (int u = 0; u < 4; u++)
          mx4[u] = fmax3(mx4[u], __uint_as_float(sr[c + 2 * u]), __uint_as_float(sr[c + 2 * u + 1]));
      }
      const float mxs = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;   // scale > 0
      // lazy reference: keep m_ref while the block stays within 2^tau of it
      const bool need = mxs > m_ref + kLazyTau;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mxs : m_ref;
        const float alpha = ex2_approx(m_ref - m_new);           // 1 where nothing moved, 0 on the first block
        if (j > 0) {
          // O must be final for block j-1 before it is rescaled (P V(j) cannot have been issued: it needs this warp's P)
          mbar_wait(&pv_done, (j - 1) & 1);
          tc_fence_after();
          for (int c = 0; c < p.dk16; c += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_off + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                "%14, %15, %16};" ::"r"(tmem_o + lane_off + c),
                "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]), "r"(o[8]),
                "r"(o[9]), "r"(o[10]), "r"(o[11]), "r"(o[12]), "r"(o[13]), "r"(o[14]), "r"(o[15])
                : "memory");
          }
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      // P = exp2(s * scale - m_ref): one FFMA2 + two EX2 + one pack + one FADD2 per pair
      const uint64_t nm2 = f2_pack(-m_ref, -m_ref);
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int c = 0; c < BKV; c += 2) {
        const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(sr[c]), __uint_as_float(sr[c + 1])), sc2, nm2);
        float x0, x1;
        f2_unpack(x2, x0, x1);
        const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
        pk[c >> 1] = pack_bf16x2(e0, e1);
        acc2[(c >> 1) & 3] = f2_add(acc2[(c >> 1) & 3], f2_pack(e0, e1));
      }
      // P over the first 32 columns of this S buffer (every S value of this row is in registers by now)
      tmem_st32(t_s, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(acc2[0], acc2[1]), a0, a1);
        f2_unpack(f2_add(acc2[2], acc2[3]), b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
    }
    // ------------------------------- epilogue: O / l -> global -------------------------------
    mbar_wait(&pv_done, (nb - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int s = q0 + row;
    __nv_bfloat16* dst = p.out + ((long long)batch * p.sq + s) * p.o_ld + head * p.d;
    if (p.kv_splits > 1) {
      // partial result of this key range: unnormalised O (fp32) + (reference max, sum); attention_merge_kernel finishes
      const long long sb_ = (long long)split * (gridDim.z / p.kv_splits) + batch;            // (split, batch) plane
      float* wo = p.ws_o + (sb_ * p.sq + s) * (p.heads * p.d) + head * p.d;
      if (s < p.sq) {
        float* ml = p.ws_ml + ((sb_ * p.heads + head) * p.sq + s) * 2;
        ml[0] = m_ref;
        ml[1] = l_run;
      }
      for (int c = 0; c < p.dk16; c += 16) {
        uint32_t o[16];
        tmem_ld16(tmem_o + lane_off + c, o);
        tmem_ld_wait();
        if (s < p.sq) {
#pragma unroll
          for (int g = 0; g < 4; g++)
            if (c + g * 4 < p.d)
              *reinterpret_cast<float4*>(wo + c + g * 4) = make_float4(__uint_as_float(o[g * 4]), __uint_as_float(o[g * 4 + 1]),
                                                                       __uint_as_float(o[g * 4 + 2]), __uint_as_float(o[g * 4 + 3]));
        }
      }
    } else
    for (int c = 0; c < p.dk16; c += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_o + lane_off + c, o);
      tmem_ld_wait();
      if (s < p.sq) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
          if (c + g * 8 < p.d) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + c + g * 8) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT checker: one thread per (batch, head, query); two passes over the keys.
__global__ void attention_simt_kernel(const __nv_bfloat16* q, long long q_ld, const __nv_bfloat16* k, const __nv_bfloat16* v,
                                      long long kv_ld, __nv_bfloat16* out, long long o_ld, int batch, int heads, int d,
                                      int sq, int skv, float scale) {
  pdl_wait();
  const long long total = (long long)batch * heads * sq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % sq);
    const int h = (int)((i / sq) % heads);
    const int b = (int)(i / ((long long)sq * heads));
    const __nv_bfloat16* qp = q + ((long long)b * sq + s) * q_ld + h * d;
    float mx = -INFINITY;
    for (int t = 0; t < skv; t++) {
      const __nv_bfloat16* kp = k + ((long long)b * skv + t) * kv_ld + h * d;
      float acc = 0.f;
      for (int c = 0; c < d; c++) acc += __bfloat162float(qp[c]) * __bfloat162float(kp[c]);
      mx = fmaxf(mx, acc * scale);
    }
    float o[256];
    for (int c = 0; c < d; c++) o[c] = 0.f;
    float l = 0.f;
    for (int t = 0; t < skv; t++) {
      const __nv_bfloat16* kp = k + ((long long)b * skv + t) * kv_ld + h * d;
      const __nv_bfloat16* vp = v + ((long long)b * skv + t) * kv_ld + h * d;
      float acc = 0.f;
      for (int c = 0; c < d; c++) acc += __bfloat162float(qp[c]) * __bfloat162float(kp[c]);
      const float e = expf(acc * scale - mx);
      l += e;
      for (int c = 0; c < d; c++) o[c] += e * __bfloat162float(vp[c]);
    }
    __nv_bfloat16* dst = out + ((long long)b * sq + s) * o_ld + h * d;
    for (int c = 0; c < d; c++) dst[c] = __float2bfloat16(o[c] / l);
  }
}

}  // namespace onedc

namespace onedc {
// out[b, s, h*d + c] = sum_i w_i O_i[c] / sum_i w_i l_i,  w_i = 2^(m_i - max_i m_i): thread = (b, s, head), whose d
// channels are contiguous in every operand (consecutive threads = consecutive heads = consecutive memory).
__global__ void __launch_bounds__(128) attention_merge_kernel(const float* ws_o, const float* ws_ml, __nv_bfloat16* out,
                                                              long long o_ld, int batch, int heads, int d, int sq, int ks) {
  pdl_wait();
  const long long total = (long long)batch * sq * heads;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int head = (int)(i % heads);
  const long long bs = i / heads;                       // b * sq + s
  const int s = (int)(bs % sq), b = (int)(bs / sq);
  float m[4], w[4];
  float mx = -INFINITY;
  for (int k = 0; k < ks; k++) {
    m[k] = ws_ml[((((long long)k * batch + b) * heads + head) * sq + s) * 2];
    mx = fmaxf(mx, m[k]);
  }
  float L = 0.f;
  for (int k = 0; k < ks; k++) {
    w[k] = exp2f(m[k] - mx);
    L += w[k] * ws_ml[((((long long)k * batch + b) * heads + head) * sq + s) * 2 + 1];
  }
  const float inv = 1.f / L;
  __nv_bfloat16* dst = out + bs * o_ld + head * d;
  const long long plane = (long long)batch * sq * heads * d;
  const float* src = ws_o + bs * ((long long)heads * d) + head * d;
  for (int c = 0; c < d; c += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ks; k++) {
      const float4 a = *reinterpret_cast<const float4*>(src + k * plane + c);
      const float4 bq = *reinterpret_cast<const float4*>(src + k * plane + c + 4);
      acc[0] += w[k] * a.x; acc[1] += w[k] * a.y; acc[2] += w[k] * a.z; acc[3] += w[k] * a.w;
      acc[4] += w[k] * bq.x; acc[5] += w[k] * bq.y; acc[6] += w[k] * bq.z; acc[7] += w[k] * bq.w;
    }
    uint4 v;
    v.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
    v.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    v.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
    v.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
    *reinterpret_cast<uint4*>(dst + c) = v;
  }
}

// Key-range splits.  Measured on B200 in round 1 (S = 9216, d = 40; 576 CTAs): no split 291 us, 2 / 3 / 4 splits 313 / 325 /
// 337 us -- a better-filled last wave does not pay for the merge pass.  Default: no split; the variant stays selectable.
// (Round 1 also had a one-S-buffer / three-CTAs-per-SM instance; it was never faster and is gone now that P lives in the
// S buffer: with one buffer Q K(j+1)^T would overwrite P(j).  `s_buffers` of onedc_attention_set_plan is ignored.)
static int g_force_ks = 0;
static void attention_plan(int batch, int heads, int head_dim, int sq, int skv, int* ks) {
  (void)batch; (void)heads; (void)sq; (void)head_dim;
  const int nblk = (skv + BKV - 1) / BKV;
  static const char* e_ks = getenv("ONEDC_ATTN_KVSPLIT");
  int want_ks = g_force_ks ? g_force_ks : (e_ks != nullptr ? e_ks[0] - '0' : 1);
  if (want_ks < 1 || want_ks > 4) want_ks = 1;
  while (want_ks > 1 && (long long)(want_ks - 1) * ((nblk + want_ks - 1) / want_ks) >= nblk) want_ks--;   // every split needs work
  *ks = want_ks;
}
}  // namespace onedc

using namespace onedc;

extern "C" void onedc_attention_set_plan(int32_t s_buffers, int32_t kv_splits) {
  (void)s_buffers;
  g_force_ks = kv_splits;
}

extern "C" int64_t onedc_attention_ws_floats(int32_t batch, int32_t heads, int32_t head_dim, int32_t sq, int32_t skv) {
  int ks;
  attention_plan(batch, heads, head_dim, sq, skv, &ks);
  if (ks == 1) return 0;
  return (int64_t)ks * batch * sq * heads * (head_dim + 2);
}

extern "C" int onedc_attention(const void* q, int64_t q_ld, const void* k, const void* v, int64_t kv_ld, void* out,
                               int64_t o_ld, int32_t batch, int32_t heads, int32_t head_dim, int32_t sq, int32_t skv,
                               float scale, int32_t impl, float* ws, int64_t ws_floats, void* stream) {
  ONEDC_CHECK(head_dim % 8 == 0 && head_dim >= 8 && head_dim <= 192, "attention: head_dim must be a multiple of 8, <= 192");
  ONEDC_CHECK(q_ld % 8 == 0 && kv_ld % 8 == 0 && o_ld % 8 == 0, "attention: leading dims must be multiples of 8");
  ONEDC_CHECK(sq > 0 && skv > 0 && scale > 0.f, "attention: empty sequence or non-positive scale");
  cudaStream_t st = (cudaStream_t)stream;
  if (impl == 1) {
    const long long total = (long long)batch * heads * sq;
    int blocks = (int)((total + 63) / 64);
    if (blocks > 148 * 32) blocks = 148 * 32;
    ONEDC_CUDA(launch_k(attention_simt_kernel, blocks, 64, 0, st, (const __nv_bfloat16*)q, q_ld, (const __nv_bfloat16*)k,
                                                 (const __nv_bfloat16*)v, kv_ld, (__nv_bfloat16*)out, o_ld, batch, heads,
                                                 head_dim, sq, skv, scale));
    ONEDC_CUDA(cudaGetLastError());
    return 0;
  }
  AttnParams p;
  p.sq = sq;
  p.skv = skv;
  p.d = head_dim;
  p.dk16 = (head_dim + 15) / 16 * 16;
  p.nchunk = (head_dim + 63) / 64;
  p.nblk = (skv + BKV - 1) / BKV;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = (__nv_bfloat16*)out;
  p.o_ld = o_ld;
  p.heads = heads;
  // key-range splits need the caller's scratch
  attention_plan(batch, heads, head_dim, sq, skv, &p.kv_splits);
  if (p.kv_splits > 1 && (ws == nullptr || ws_floats < (int64_t)p.kv_splits * batch * sq * heads * (head_dim + 2))) {
    ONEDC_CHECK(ws == nullptr, "attention: scratch too small (see onedc_attention_ws_floats)");
    p.kv_splits = 1;
  }
  p.bps = (p.nblk + p.kv_splits - 1) / p.kv_splits;
  p.ws_o = ws;
  p.ws_ml = ws != nullptr ? ws + (int64_t)p.kv_splits * batch * sq * heads * head_dim : nullptr;
  p.tmem_cols = (2 * BKV + p.dk16 <= 256) ? 256 : 512;
  CUtensorMap mq, mk, mv;
  {
    uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)sq, (uint64_t)heads, (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)q_ld * 2, (uint64_t)head_dim * 2, (uint64_t)sq * q_ld * 2};
    uint32_t box[4] = {64, 128, 1, 1};
    int rc = make_tensor_map(&mq, q, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)skv, (uint64_t)heads, (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)kv_ld * 2, (uint64_t)head_dim * 2, (uint64_t)skv * kv_ld * 2};
    uint32_t box[4] = {64, BKV, 1, 1};
    int rc = make_tensor_map(&mk, k, 4, dims, str, box);
    if (rc) return rc;
    rc = make_tensor_map(&mv, v, 4, dims, str, box);
    if (rc) return rc;
  }
  const size_t smem = (size_t)p.nchunk * 16384 + 4 * (size_t)p.nchunk * BKV * 128 + 1024;
  static size_t attr = 0;
  if (smem > attr) {
    ONEDC_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((sq + 127) / 128, heads, batch * p.kv_splits);
  ONEDC_CUDA(launch_k(attention_tc_kernel, grid, kAttnThreads, smem, st, mq, mk, mv, p));
  if (p.kv_splits > 1) {
    const long long total = (long long)batch * sq * heads;
    ONEDC_CUDA(launch_k(attention_merge_kernel, (int)((total + 127) / 128), 128, 0, st, (const float*)p.ws_o, (const float*)p.ws_ml,
                        (__nv_bfloat16*)out, o_ld, batch, heads, head_dim, sq, p.kv_splits));
  }
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
