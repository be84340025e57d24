// Flash attention for the UNet's self/cross attention on tcgen05 (sm_100a).
//
//   O[b, s, h*d : (h+1)*d] = softmax(Q K^T * scale) V      per (batch b, head h), non-causal
//
// One CTA = 128 queries of one (b, h).  Q/K/V are read straight out of the [b, s, heads*d] projection
// outputs through 4-D tensor maps {d, s, head, b}: the box is 64 channels wide, so for d = 40/80/160 the
// columns beyond d are out of bounds and TMA zero-fills them (no padded copies, no head split kernel).
//   S = Q K^T     : tcgen05.mma, A = Q (K-major, smem), B = K tile (K-major, smem)  -> TMEM (2 buffers)
//   softmax       : 4 warps, thread = query row, S row held in registers, exp2 with running max/sum,
//                   P written as bf16 into 128B-swizzled smem; O rescaled in TMEM only when a max moved
//   O += P V      : tcgen05.mma, A = P (K-major, smem), B = V tile (MN-major, smem as loaded by TMA)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax + output.  QK^T of block j+1 is
// issued before P V of block j, so the tensor core works on the next scores while softmax runs; two CTAs
// per SM (TMEM 256 columns each) overlap one CTA's softmax with the other's MMAs.
// Launch plans (attention_plan / onedc_attention_set_plan): the default above; one S buffer + three CTAs per SM for
// head_dim <= 64; the key range split over 2..4 CTAs per query tile with an fp32 merge pass (attention_merge_kernel).
// The alternatives are measured slower on B200 for the UNet's shapes (numbers at attention_plan) and stay tested.
//
// Round 2 (tools/experimental/attention_event_pipeline.cu, profiles/attn_*_r2*): an event-driven rewrite -- P in tensor
// memory as the A operand of P V, lazy exponent reference, packed fp32x2 arithmetic, scores issued two blocks ahead, the next
// block's scores prefetched into registers, 32-key blocks with four CTAs per SM -- measured 300 us for S = 9216, d = 40 in
// every variant, like this kernel, and stayed at 300 us with every exponential or every MMA removed: neither the MUFU
// (~55 % busy) nor the tensor pipe (~22 %) is the bound.  It deadlocked under the pipelined decoder (three graphs in flight,
// about one run in three; the MMA warp waiting for the softmax warps' first event) and is therefore not the product
// kernel; this one has run every driver benchmark since round 1.
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

struct AttnParams {
  int sq, skv, d, dk16, nchunk;      // dk16 = round_up(d,16), nchunk = ceil(d/64)
  int nblk;
  float scale_log2;
  __nv_bfloat16* out;
  long long o_ld;
  int heads;
  int tmem_cols;   // 128: one S buffer + O (head_dim <= 64, three CTAs per SM); 256: two S buffers + O; else 512
  int nsbuf;       // S buffers in TMEM (1 or 2)
  // key-range split (wave quantisation): CTA z = batch * kv_splits + split handles key blocks [split*bps, +bps) and,
  // when kv_splits > 1, leaves an unnormalised fp32 O and its (running max, sum) for attention_merge_kernel
  int kv_splits, bps;
  float* ws_o;     // [kv_splits][batch][sq][heads*d] fp32
  float* ws_ml;    // [kv_splits][batch][heads][sq][2] fp32
};

__device__ __forceinline__ void attention_tc_body(const CUtensorMap& map_q, const CUtensorMap& map_k,
                                                  const CUtensorMap& map_v, const AttnParams& p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_free, p_full, pv_done;
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
  uint8_t* sP = sV + 2 * kv_bytes;                 // [128 rows][128 B]  (BKV = 64 bf16 per row)

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(&s_free, 4);       // single S buffer: the four softmax warps hold S(j) in registers
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
  const int nsbuf = p.nsbuf;
  const uint32_t tmem_o = tmem_base + nsbuf * BKV;  // S buffer(s) at columns [0, nsbuf*BKV), O after them

  // The TMA and MMA roles are single-lane jobs run by the WHOLE warp with only the TMA / MMA / commit instructions
  // under `if (leader)` and every address a register bumped by constants: the loop state stays in uniform registers
  // (as in igemm.cu).  Entered under `if (lane == 0)` the MMA issuer spent ~1500 clocks of scalar latency per 64-key
  // block on its 7 MMAs and 4 commits -- more than the softmax warps need for the block -- and bounded the kernel.
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV), sP_a = smem_u32(sP);
  const uint32_t q_full_a = smem_u32(&q_full), k_full_a = smem_u32(&k_full[0]), k_empty_a = smem_u32(&k_empty[0]);
  const uint32_t v_full_a = smem_u32(&v_full[0]), v_empty_a = smem_u32(&v_empty[0]), s_full_a = smem_u32(&s_full[0]);
  const uint32_t s_free_a = smem_u32(&s_free), p_full_a = smem_u32(&p_full), pv_done_a = smem_u32(&pv_done);
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
    const uint32_t idesc_pv = umma_idesc_bf16(128, p.dk16, 0, 1);   // B = V is MN-major
    // descriptor = constant high part | (shared address >> 4)
    const uint64_t dhi_k = umma_smem_desc(0, 16, 1024);              // K-major tiles: Q, K, P
    const uint64_t dhi_v = umma_smem_desc(0, BKV * 128, 1024);       // V: MN(d)-major, 64-wide d chunks BKV*128 B apart
    const uint32_t q_enc = (sQ_a & 0x3FFFF) >> 4, k_enc = (sK_a & 0x3FFFF) >> 4, v_enc = (sV_a & 0x3FFFF) >> 4;
    const uint32_t p_enc = (sP_a & 0x3FFFF) >> 4, kv_enc = (uint32_t)kv_bytes >> 4;
    const int ksteps = p.dk16 / 16;
    mbar_wait_a(q_full_a, 0);
    for (int j = 0; j <= nb; j++) {
      if (j < nb) {
        // ---- S(j) = Q K(j)^T
        const uint32_t st = j & 1, sb = nsbuf == 2 ? st : 0;
        mbar_wait_a(k_full_a + st * 8, (j >> 1) & 1);
        if (nsbuf == 1 && j >= 1) mbar_wait_a(s_free_a, (j - 1) & 1);   // softmax(j-1) has copied S(j-1) out of TMEM
        tc_fence_after();
        if (leader) {
          const uint32_t kst = k_enc + st * kv_enc;
          for (int kk = 0; kk < ksteps; kk++) {
            const uint32_t c = kk >> 2, w = kk & 3;
            const uint64_t da = dhi_k | (uint64_t)(q_enc + c * (16384 >> 4) + w * 2);
            const uint64_t db = dhi_k | (uint64_t)(kst + c * (BKV * 128 >> 4) + w * 2);
            umma_bf16(tmem_base + sb * BKV, da, db, idesc_qk, kk != 0);
          }
          umma_commit_a(s_full_a + sb * 8);
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
#pragma unroll
          for (int kk = 0; kk < BKV / 16; kk++) {
            const uint64_t da = dhi_k | (uint64_t)(p_enc + kk * 2);
            const uint64_t db = dhi_v | (uint64_t)(vst + kk * (2048 >> 4));      // 8 kv rows per 1024-byte atom
            umma_bf16(tmem_o, da, db, idesc_pv, (jp | kk) != 0);
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
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = sP + (row >> 3) * 1024 + (row & 7) * 128;
    for (int j = 0; j < nb; j++) {
      const int sb = nsbuf == 2 ? (j & 1) : 0;
      mbar_wait(&s_full[sb], nsbuf == 2 ? (j >> 1) & 1 : j & 1);
      tc_fence_after();
      uint32_t sr[BKV];
      tmem_ld32(tmem_base + lane_off + sb * BKV, sr);
      tmem_ld32(tmem_base + lane_off + sb * BKV + 32, sr + 32);
      tmem_ld_wait();
      if (nsbuf == 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free);          // the MMA warp may overwrite S with Q K(j+1)^T
      }
      const int nvalid = p.skv - (j0 + j) * BKV;     // columns >= nvalid are zero-filled padding keys
      if (nvalid < BKV) {                            // only the last block can be partial (uniform branch)
#pragma unroll
        for (int c = 0; c < BKV; c++)
          if (c >= nvalid) sr[c] = 0xff800000u;      // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < BKV; c++) mx = fmaxf(mx, __uint_as_float(sr[c]));
      const float m_new = fmaxf(m_run, mx * p.scale_log2);     // scale > 0: max commutes with the scaling
      const float alpha = exp2f(m_run - m_new);      // 0 on the first block (m_run = -inf)
      float rs = 0.f;
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int c = 0; c < BKV; c += 2) {
        // exp(s*scale - m) as one FFMA + one EX2 per element
        const float e0 = exp2f(fmaf(__uint_as_float(sr[c]), p.scale_log2, -m_new));
        const float e1 = exp2f(fmaf(__uint_as_float(sr[c + 1]), p.scale_log2, -m_new));
        pk[c >> 1] = pack_bf16x2(e0, e1);
        rs += e0 + e1;
      }
      l_run = l_run * alpha + rs;
      // P buffer is free and O is final for block j-1 once PV(j-1) has completed
      if (j > 0) {
        mbar_wait(&pv_done, (j - 1) & 1);
        tc_fence_after();
        const bool need = m_new > m_run;
        if (__any_sync(0xffffffffu, need)) {
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
          tmem_st_wait();
        }
      }
      m_run = m_new;
      // write P row: 8 x 16-byte chunks, chunk index XOR (row & 7)  (SWIZZLE_128B, K-major)
#pragma unroll
      for (int ch = 0; ch < BKV / 8; ch++) {
        uint4 v = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();       // generic-proxy smem writes -> visible to the async proxy (UMMA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
    }
    // ------------------------------- epilogue: O / l -> global -------------------------------
    mbar_wait(&pv_done, (nb - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int s = q0 + row;
    __nv_bfloat16* dst = p.out + ((long long)batch * p.sq + s) * p.o_ld + head * p.d;
    if (p.kv_splits > 1) {
      // partial result of this key range: unnormalised O (fp32) + (max, sum); attention_merge_kernel finishes
      const long long sb_ = (long long)split * (gridDim.z / p.kv_splits) + batch;            // (split, batch) plane
      float* wo = p.ws_o + (sb_ * p.sq + s) * (p.heads * p.d) + head * p.d;
      if (s < p.sq) {
        float* ml = p.ws_ml + ((sb_ * p.heads + head) * p.sq + s) * 2;
        ml[0] = m_run;
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

// Two instances of the same body: 2 CTAs per SM (two S buffers, 256 / 512 TMEM columns) and, for head_dim <= 64, 3 CTAs
// per SM (one S buffer, 128 TMEM columns, 65 KB of shared memory, 112 registers).  A CTA needs ~2200 clocks per 64-key
// block either way -- the softmax warps are bound by their own dependent-issue latency -- so the third CTA is worth
// +40 % per SM, but only when the grid fills the extra slots: the host picks the variant per launch by wave count.
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
  attention_tc_body(map_q, map_k, map_v, p);
}
__global__ void __maxnreg__(112)
attention_tc3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
  attention_tc_body(map_q, map_k, map_v, p);
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

// (CTAs per SM, key-range splits).  Measured on B200 (S = 9216, d = 40; 576 CTAs): 2 CTAs/SM 291 us, 3 CTAs/SM 297 us,
// 2 / 3 / 4 key-range splits 313 / 325 / 337 us -- the per-SM throughput is the same with 2 or 3 resident CTAs (several
// units at ~50 %: MUFU, shared memory, TMEM reads, issue), so neither the third CTA nor a better-filled last wave pays
// for its overhead.  Default: 2 CTAs/SM, no split; both variants stay selectable (tests, other shapes).
static int g_force_nsbuf = 0, g_force_ks = 0;
static void attention_plan(int batch, int heads, int head_dim, int sq, int skv, int* nsbuf, int* ks) {
  (void)batch; (void)heads; (void)sq;
  const int dk16 = (head_dim + 15) / 16 * 16, nblk = (skv + BKV - 1) / BKV;
  static const char* e_ns = getenv("ONEDC_ATTN_NSBUF");
  static const char* e_ks = getenv("ONEDC_ATTN_KVSPLIT");
  int want_ns = g_force_nsbuf ? g_force_nsbuf : (e_ns != nullptr ? e_ns[0] - '0' : 2);
  int want_ks = g_force_ks ? g_force_ks : (e_ks != nullptr ? e_ks[0] - '0' : 1);
  if (want_ns == 1 && BKV + dk16 > 128) want_ns = 2;                       // one S buffer + O must fit 128 TMEM columns
  if (want_ks < 1 || want_ks > 4) want_ks = 1;
  while (want_ks > 1 && (long long)(want_ks - 1) * ((nblk + want_ks - 1) / want_ks) >= nblk) want_ks--;   // every split needs work
  *nsbuf = want_ns == 1 ? 1 : 2;
  *ks = want_ks;
}
}  // namespace onedc

using namespace onedc;

extern "C" void onedc_attention_set_plan(int32_t s_buffers, int32_t kv_splits) {
  g_force_nsbuf = s_buffers;
  g_force_ks = kv_splits;
}

extern "C" int64_t onedc_attention_ws_floats(int32_t batch, int32_t heads, int32_t head_dim, int32_t sq, int32_t skv) {
  int nsbuf, ks;
  attention_plan(batch, heads, head_dim, sq, skv, &nsbuf, &ks);
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
  // (CTAs per SM, key-range splits) by wave count; splitting needs the caller's scratch
  attention_plan(batch, heads, head_dim, sq, skv, &p.nsbuf, &p.kv_splits);
  if (p.kv_splits > 1 && (ws == nullptr || ws_floats < (int64_t)p.kv_splits * batch * sq * heads * (head_dim + 2))) {
    ONEDC_CHECK(ws == nullptr, "attention: scratch too small (see onedc_attention_ws_floats)");
    p.kv_splits = 1;
    p.nsbuf = 2;
  }
  p.bps = (p.nblk + p.kv_splits - 1) / p.kv_splits;
  p.ws_o = ws;
  p.ws_ml = ws != nullptr ? ws + (int64_t)p.kv_splits * batch * sq * heads * head_dim : nullptr;
  p.tmem_cols = p.nsbuf == 1 ? 128 : ((2 * BKV + p.dk16 <= 256) ? 256 : 512);
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
  const size_t smem = (size_t)p.nchunk * 16384 + 4 * (size_t)p.nchunk * BKV * 128 + 16384 + 1024;
  static size_t attr = 0;
  if (smem > attr) {
    ONEDC_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ONEDC_CUDA(cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((sq + 127) / 128, heads, batch * p.kv_splits);
  if (p.nsbuf == 1)
    ONEDC_CUDA(launch_k(attention_tc3_kernel, grid, kAttnThreads, smem, st, mq, mk, mv, p));
  else
    ONEDC_CUDA(launch_k(attention_tc_kernel, grid, kAttnThreads, smem, st, mq, mk, mv, p));
  if (p.kv_splits > 1) {
    const long long total = (long long)batch * sq * heads;
    ONEDC_CUDA(launch_k(attention_merge_kernel, (int)((total + 127) / 128), 128, 0, st, (const float*)p.ws_o, (const float*)p.ws_ml,
                        (__nv_bfloat16*)out, o_ld, batch, heads, head_dim, sq, p.kv_splits));
  }
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
