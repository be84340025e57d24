// Flash attention for the UNet's self/cross attention on tcgen05 (sm_100a).
//
//   O[b, s, h*d : (h+1)*d] = softmax(Q K^T * scale) V      per (batch b, head h), non-causal
//
// One CTA = 128 queries of one (b, h).  Q/K/V are read straight out of the [b, s, heads*d] projection
// outputs through 4-D tensor maps {d, s, head, b}: the box is 64 channels wide, so for d = 40/80/160 the
// columns beyond d are out of bounds and TMA zero-fills them (no padded copies, no head split kernel).
//   S = Q K^T     : tcgen05.mma, A = Q (K-major, smem), B = K tile (K-major, smem)  -> TMEM (2 buffers of BKV columns)
//   softmax       : 4 warps, thread = query row, S row in registers; exp2 against a LAZY reference maximum;
//                   P (bf16) goes into a P buffer of its own in TMEM (tcgen05.st), never through shared memory
//   O += P V      : tcgen05.mma with the A operand (P) read from TMEM, B = V tile (MN-major, smem as loaded by TMA)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax + output; two CTAs per SM (64-key blocks, 256
// TMEM columns each) or four (32-key blocks, 128 columns).
//
// Pipeline (round 2).  The softmax warps signal ONE event per block, "P(e) is stored and S(e+1) is already in my
// registers" (they load the next block's scores before the exponentials of the current one).  On event e the MMA warp
// issues P V(e) and Q K(e+3)^T -- the buffer of S(e+1) is free -- so scores are issued two block times before they are
// asked for and no softmax warp ever waits for the MMA warp's wake-up / issue / commit chain (~1000 clocks).  Every
// buffer hand-over is a barrier the consumer waits on; nothing relies on the order in which the tensor pipe executes
// MMAs.  (Round 1 kept P in the first columns of the S buffer it came from and relied on "P V(j) was issued before
// Q K(j+2)^T"; with more than two CTAs per SM that produced NaNs and hangs.)
//
// What the measurements say (profiles/attn_*_r2*, tools/attn_roles.py, tools/micro/{exp_loop,tmem_read,tmem_alloc4}.cu):
//   * S = 9216, d = 40, 8 heads: 300 us = 360 TFLOP/s in EVERY variant tried: 64-key blocks x 2 CTAs per SM, 32-key blocks
//     x 3 or 4 CTAs, K / V rings of 2 or 4 stages, with and without the register prefetch of S(j+1), with the MMA warp's
//     loop at 250 or 130 instructions per block.  It also stays at 300 us with every exponential replaced by a multiply
//     and with every MMA removed: neither the MUFU (16 exp2/clk/SM, ~55 % busy) nor the tensor pipe (~22 %) bounds it.
//   * the inner loop alone (FFMA2 -> 2 x MUFU.EX2 -> F2FP -> FADD2) runs at 16 elements/clk/SM with two warps per
//     scheduler; tcgen05.ld delivers 180 (4 warps) .. 340 (8 warps) B/clk/SM, ten times what the kernel needs.
//   * one CTA alone on an SM needs ~1700 clocks per 64-key block (1300 without the exponentials): a single warp issuing
//     ~400 dependent-ish instructions and ~8 synchronising operations (try_wait, tcgen05.wait, fences, arrive) per block.
//   The remaining suspects are the L2 side (all 72 query tiles of a head stream the same K / V rows in lockstep: 1.7 GB
//   of L2 -> SM traffic per launch, busiest slice at 67 %) and the synchronising operations themselves.  Not resolved.
// The lazy rescale (the exponent reference m_ref moves only when the block maximum exceeds it by more than 8 in log2
// units: P <= 256, O is rescaled in ~1 % of the blocks), packed fp32x2 arithmetic and 3-input max come from round 2's
// first pass and are kept.
// Launch plans (onedc_attention_set_plan): key block 32 / 64; the key range may be split over 2..4 CTAs per query tile
// with an fp32 merge pass (attention_merge_kernel); measured slower for the UNet's shapes, kept selectable and tested.
//
// The SIMT kernel at the bottom is the on-GPU checker (impl = 1), never used by the decode path.
#include "../../include/onedc_b200.h"
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace onedc {

int make_tensor_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);

constexpr int kBkvWide = 64, kBkvNarrow = 32;   // keys per block (template parameter BKV of the kernel)
constexpr int kAttnThreads = 192;
constexpr uint32_t kSleepTma = 100, kSleepMma = 20;   // ns between polls of the producer / issuer warps (see mbar_wait_sleep_a)
constexpr float kLazyTau = 8.f;   // log2 units: P <= 2^8, far inside bf16 / fp32 range

struct AttnParams {
  int sq, skv, d, dk16, nchunk;      // dk16 = round_up(d,16), nchunk = ceil(d/64)
  int nblk;
  float scale_log2;
  __nv_bfloat16* out;
  long long o_ld;
  int heads;
  int tmem_cols;   // power of two >= 2 * BKV (S buffers) + dk16 (O) + np * BKV / 2 (P buffers)
  int np;          // P buffers: 2 where they fit the allocation that one buffer needs anyway, else 1
  // key-range split (wave quantisation): CTA z = batch * kv_splits + split handles key blocks [split*bps, +bps) and,
  // when kv_splits > 1, leaves an unnormalised fp32 O and its (reference max, sum) for attention_merge_kernel
  int kv_splits, bps;
  float* ws_o;     // [kv_splits][batch][sq][heads*d] fp32
  float* ws_ml;    // [kv_splits][batch][heads][sq][2] fp32
  long long* dbg;  // optional phase clocks of the softmax warps (onedc_attention_set_debug): [CTA][4 warps][8] + one CTA's trace
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

template <int BKV, int MINB, bool TIMED, int NST, int NCH, bool PREF>
__global__ void __launch_bounds__(kAttnThreads, MINB)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[NST], k_empty[NST], v_full[NST], v_empty[NST], s_full[2], p_full[2], pv_done[2];
  __shared__ uint32_t tmem_slot;

  // broadcast so that the compiler knows the warp index is warp-uniform (role branches stay uniform)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y;
  const int batch = blockIdx.z / p.kv_splits, split = blockIdx.z - batch * p.kv_splits;
  const int j0 = split * p.bps;                                        // first key block of this CTA
  const int nb = (p.nblk - j0 < p.bps) ? p.nblk - j0 : p.bps;         // its number of key blocks (>= 1, host-checked)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int q_bytes = NCH * 16384;             // [chunk][128 rows][128 B]     (NCH = ceil(head_dim / 64))
  constexpr int kv_bytes = NCH * BKV * 128;        // [chunk][BKV rows][128 B]
  // PREF: the softmax warps load S(j+1) during block j (64 more registers).  Event e then also means "S(e+1) is in
  // registers", and Q K(e+AHEAD)^T may be issued on it.  Without PREF (32-key blocks, four CTAs per SM) a block loads its
  // own scores first, and event e releases the buffer of S(e).
  constexpr int AHEAD = PREF ? 3 : 2;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_bytes;                      // NST stages
  uint8_t* sV = sK + NST * kv_bytes;               // NST stages

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < NST; i++) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&s_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&p_full[i], 4);  // one arrive per softmax warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();
  // tensor memory: S buffers at columns [0, 2 BKV), O behind them, then one or two P buffers of BKV / 2 columns (bf16 pairs)
  const uint32_t tmem_o = tmem_base + 2 * BKV;
  const uint32_t tmem_p0 = tmem_o + (uint32_t)p.dk16;
  const uint32_t p_stride = p.np == 2 ? BKV / 2 : 0;
  // every mbarrier wait of the kernel; the timed build adds a watchdog that reports the stuck (site, index) and traps
  auto wait_bar = [&](const uint32_t bar, const uint32_t parity, const uint32_t sleep_ns, const int site, const int idx) {
    if (!TIMED) {
      if (sleep_ns) mbar_wait_sleep_a(bar, parity, sleep_ns); else mbar_wait_a(bar, parity);
      return;
    }
    for (long long spins = 0;; spins++) {
      uint32_t done;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
      if (done) break;
      if (sleep_ns) __nanosleep(sleep_ns);
      if (spins > (sleep_ns ? 2000000ll : 20000000ll)) {
        if ((threadIdx.x & 31) == 0 && p.dbg != nullptr) {
          // plain stores into (possibly host-mapped) memory: slot by CTA and warp, collisions just overwrite;
          // second word: the barrier's raw state
          volatile long long* o = p.dbg + (size_t)gridDim.x * gridDim.y * gridDim.z * 40;
          const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
          unsigned long long st;
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(st) : "r"(bar));
          const unsigned k = (cta * 6 + (threadIdx.x >> 5)) % 96;
          o[2 * k] = (1ll << 62) | ((long long)blockIdx.x << 40) | ((long long)blockIdx.y << 32) |
                     ((long long)(threadIdx.x >> 5) << 24) | ((long long)site << 16) | (long long)(idx & 0xffff);
          o[2 * k + 1] = (long long)st;
          __threadfence_system();
        }
        for (int w = 0; w < 3000; w++) __nanosleep(1000000);      // let the other stuck warps of the GPU report as well
        __nanosleep(1000000);
        __trap();
      }
    }
  };

  // The TMA and MMA roles are single-lane jobs run by the WHOLE warp with only the TMA / MMA / commit instructions
  // under `if (leader)` and every address a register bumped by constants: the loop state stays in uniform registers
  // (as in igemm.cu).
  const uint32_t sQ_a = smem_u32(sQ), sK_a = smem_u32(sK), sV_a = smem_u32(sV);
  const uint32_t q_full_a = smem_u32(&q_full), k_full_a = smem_u32(&k_full[0]), k_empty_a = smem_u32(&k_empty[0]);
  const uint32_t v_full_a = smem_u32(&v_full[0]), v_empty_a = smem_u32(&v_empty[0]), s_full_a = smem_u32(&s_full[0]);
  const uint32_t p_full_a = smem_u32(&p_full[0]), pv_done_a = smem_u32(&pv_done[0]);
  if (warp == 0) {
    const uint32_t leader = elect_one();
    if (leader) {
      mbar_expect_tx_a(q_full_a, (uint32_t)q_bytes);
      for (int c = 0; c < NCH; c++) tma_load_4d_a(sQ_a + c * 16384, &map_q, q_full_a, c * 64, q0, head, batch);
    }
    __syncwarp();
    // Tiles are requested in the order the MMA warp consumes them: K(0..2), then V(e), K(e+3) per event e (a K tile is
    // needed three blocks before its scores are consumed).  Any other order can deadlock a two-stage ring: K(e+3) queued
    // behind V(e+2), whose slot is only released by P V(e).
    auto load_k = [&](const int j) {
      const uint32_t st = j % NST, ph = (j / NST) & 1;
      wait_bar(k_empty_a + st * 8, ph ^ 1, kSleepTma, 1, j);
      if (leader) {
        mbar_expect_tx_a(k_full_a + st * 8, (uint32_t)kv_bytes);
#pragma unroll
        for (int c = 0; c < NCH; c++)
          tma_load_4d_a(sK_a + st * kv_bytes + c * BKV * 128, &map_k, k_full_a + st * 8, c * 64, (j0 + j) * BKV, head, batch);
      }
      __syncwarp();
    };
    for (int j = 0; j < AHEAD && j < nb; j++) load_k(j);
    for (int e = 0; e < nb; e++) {
      const uint32_t st = e % NST, ph = (e / NST) & 1;
      wait_bar(v_empty_a + st * 8, ph ^ 1, kSleepTma, 2, e);
      if (leader) {
        mbar_expect_tx_a(v_full_a + st * 8, (uint32_t)kv_bytes);
#pragma unroll
        for (int c = 0; c < NCH; c++)
          tma_load_4d_a(sV_a + st * kv_bytes + c * BKV * 128, &map_v, v_full_a + st * 8, c * 64, (j0 + e) * BKV, head, batch);
      }
      __syncwarp();
      if (e + AHEAD < nb) load_k(e + AHEAD);
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    // Event e = "every softmax warp has stored P(e) and holds S(e+1) in registers" (one arrival per warp on p_full; e = -1
    // is the prologue load of S(0)).  It releases P V(e) and, because the buffer of S(e+1) is free again, Q K(e+3)^T: the
    // scores of a block are issued two block times before the softmax warps ask for them, so the chain
    //   P arrive -> this warp wakes -> MMA issue -> MMA -> commit -> softmax warps wake     (~1000 clocks)
    // is never waited for.  (Issued only after P V(j-2), i.e. one block ahead, S(j) arrived 250..450 clocks late in every
    // block -- a third of the softmax warps' time.)
    const uint32_t leader = elect_one();
    const uint32_t idesc_qk = umma_idesc_bf16(128, BKV, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16(128, p.dk16, 0, 1);   // A = P from TMEM (K-major), B = V is MN-major
    // descriptor = constant high part | (shared address >> 4)
    const uint64_t dhi_k = umma_smem_desc(0, 16, 1024);              // K-major tiles: Q, K
    const uint64_t dhi_v = umma_smem_desc(0, BKV * 128, 1024);       // V: MN(d)-major, 64-wide d chunks BKV*128 B apart
    const uint32_t q_enc = (sQ_a & 0x3FFFF) >> 4, k_enc = (sK_a & 0x3FFFF) >> 4, v_enc = (sV_a & 0x3FFFF) >> 4;
    const uint32_t kv_enc = (uint32_t)kv_bytes >> 4;
    const int ksteps = p.dk16 / 16;
    long long mph[4] = {0, 0, 0, 0};
    long long mtp = TIMED ? clock64() : 0;
    auto mmark = [&](int i) {
      if (TIMED) {
        const long long now = clock64();
        mph[i] += now - mtp;
        mtp = now;
      }
    };
    // The loop is unrolled over (e + 1) mod 4: ring stages, S / P buffers and barrier addresses are then compile-time offsets
    // from a handful of registers.  (Indexed by e at run time the loop was ~250 uniform-datapath instructions = ~900 clocks
    // per block -- and THAT, not the MUFU or the tensor pipe, bounded the kernel: removing every exponential or every MMA
    // left its time unchanged.)
    const uint64_t qdesc0 = dhi_k | (uint64_t)q_enc, kdesc0 = dhi_k | (uint64_t)k_enc, vdesc0 = dhi_v | (uint64_t)v_enc;
    constexpr uint32_t kv_enc_c = (uint32_t)kv_bytes >> 4;
    auto issue_qk = [&](auto SC, auto KC) {        // S(j) = Q K(j)^T into S buffer SC, K(j) in ring stage KC
      constexpr int SBUF = decltype(SC)::value, KS = decltype(KC)::value;
      if (leader) {
        const uint32_t d_tmem = tmem_base + SBUF * BKV;
        uint64_t da = qdesc0, db = kdesc0 + KS * kv_enc_c;
#pragma unroll 1
        for (int c = 0; c < NCH; c++) {
          const int ks_c = ksteps - 4 * c < 4 ? ksteps - 4 * c : 4;
          for (int w = 0; w < ks_c; w++)
            umma_bf16(d_tmem, da + 2 * w, db + 2 * w, idesc_qk, (c | w) != 0);
          da += 16384 >> 4;
          db += BKV * 128 >> 4;
        }
        umma_commit_a(s_full_a + SBUF * 8);
        umma_commit_a(k_empty_a + KS * 8);
      }
      __syncwarp();
    };
    wait_bar(q_full_a, 0, kSleepMma, 3, 0);
    wait_bar(k_full_a, 0, kSleepMma, 4, 0);
    tc_fence_after();
    issue_qk(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
    if (nb > 1) {
      wait_bar(k_full_a + 8, 0, kSleepMma, 4, 1);
      tc_fence_after();
      issue_qk(std::integral_constant<int, 1>{}, std::integral_constant<int, 1 % NST>{});
    }
    auto step = [&](auto RC, const int e) {
      constexpr int R = decltype(RC)::value;             // (e + 1) & 3
      constexpr int VS = ((R + 3) & 3) % NST;            // e % NST: V ring stage
      constexpr int KS = ((R + AHEAD + 3) & 3) % NST;    // (e + AHEAD) % NST: K ring stage
      constexpr int EB = (R + 1) & 1;                    // e & 1: P buffer, pv_done barrier
      constexpr int SB = PREF ? (R & 1) : ((R + 1) & 1); // p_full barrier of event e: (e + 1) & 1 with PREF (event -1 exists), else e & 1
      constexpr int QB = (R + AHEAD + 1) & 1;            // (e + AHEAD) & 1: S buffer of Q K(e+AHEAD)^T
      // operands that landed long ago: their barrier round trips stay off the critical path
      if (e >= 0) wait_bar(v_full_a + VS * 8, (e / NST) & 1, kSleepMma, 5, e);
      if (e + AHEAD < nb) wait_bar(k_full_a + KS * 8, ((e + AHEAD) / NST) & 1, kSleepMma, 6, e + AHEAD);
      mmark(0);
      // two barriers, by event parity: a softmax warp may run a whole block ahead of the slowest one (nothing it needs
      // for block e+1 depends on event e), and two arrivals of one warp must not land in the same phase
      wait_bar(p_full_a + SB * 8, ((e + (PREF ? 1 : 0)) >> 1) & 1, kSleepMma, 7, e + 1);
      tc_fence_after();
      mmark(1);
      if (e >= 0) {
        // ---- O += P(e) V(e)
        if (leader) {
          const uint64_t db = vdesc0 + VS * kv_enc_c;
          const uint32_t tmem_p = tmem_p0 + EB * p_stride;                // bf16 pairs: 8 columns per 16 keys
#pragma unroll
          for (int kk = 0; kk < BKV / 16; kk++)                           // 16 kv rows = two 1024-byte atoms
            umma_bf16_ts(tmem_o, tmem_p + kk * 8, db + kk * (2048 >> 4), idesc_pv, (e | kk) != 0);
          umma_commit_a(pv_done_a + EB * 8);
          umma_commit_a(v_empty_a + VS * 8);
        }
        __syncwarp();
      }
      mmark(2);
      if (e + AHEAD < nb) issue_qk(std::integral_constant<int, QB>{}, std::integral_constant<int, KS>{});
      mmark(3);
    };
    for (int e = -1; e < nb; e += 4) {
      if (PREF || e >= 0) step(std::integral_constant<int, 0>{}, e);
      if (e + 1 < nb) step(std::integral_constant<int, 1>{}, e + 1);
      if (e + 2 < nb) step(std::integral_constant<int, 2>{}, e + 2);
      if (e + 3 < nb) step(std::integral_constant<int, 3>{}, e + 3);
    }
    if (TIMED && leader && p.dbg != nullptr)
      for (int i = 0; i < 4; i++)
        p.dbg[(((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 5 + 4) * 8 + i] = mph[i];
  } else {
    // ------------------------------- softmax / output warps -------------------------------
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.f;            // exponent reference (lazy maximum), running sum w.r.t. m_ref
    const uint64_t sc2 = f2_pack(p.scale_log2, p.scale_log2);
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tp = TIMED ? clock64() : 0;
    auto mark = [&](int i) {
      if (TIMED) {
        const long long now = clock64();
        ph[i] += now - tp;
        tp = now;
      }
    };
    // One block of the online softmax.  `cur` holds S(j) (its tcgen05.ld has completed); the load of S(j+1) into `nxt` is
    // issued before the exponentials of block j, so its barrier round trip and tensor-memory latency run under the MUFU
    // work instead of in front of it.
    auto block = [&](const int j, uint32_t (&cur)[BKV], uint32_t (&nxt)[BKV]) {
      if constexpr (!PREF) {
        wait_bar(s_full_a + (j & 1) * 8, (j >> 1) & 1, 0, 9, j);
        tc_fence_after();
        const uint32_t t_c = tmem_base + lane_off + (j & 1) * BKV;
        tmem_ld32(t_c, cur);
        if constexpr (BKV == 64) tmem_ld32(t_c + 32, cur + 32);
        tmem_ld_wait();
      }
      const int nvalid = p.skv - (j0 + j) * BKV;     // columns >= nvalid are zero-filled padding keys
      if (nvalid < BKV) {                            // only the last block can be partial (uniform branch)
#pragma unroll
        for (int c = 0; c < BKV; c++)
          if (c >= nvalid) cur[c] = 0xff800000u;     // -inf
      }
      // block maximum: four independent chains of 3-input max
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < BKV; c += 8) {
#pragma unroll
        for (int u = 0; u < 4; u++)
          mx4[u] = fmax3(mx4[u], __uint_as_float(cur[c + 2 * u]), __uint_as_float(cur[c + 2 * u + 1]));
      }
      const float mxs = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;   // scale > 0
      // lazy reference: keep m_ref while the block stays within 2^tau of it
      const bool need = mxs > m_ref + kLazyTau;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mxs : m_ref;
        const float alpha = ex2_approx(m_ref - m_new);           // 1 where nothing moved, 0 on the first block
        if (j > 0) {
          // O must be final for block j-1 before it is rescaled (P V(j) cannot have been issued: it needs this warp's P)
          wait_bar(pv_done_a + ((j - 1) & 1) * 8, ((j - 1) >> 1) & 1, 0, 8, j);
          tc_fence_after();
          for (int c = 0; c < p.dk16; c += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_o + lane_off + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_o + lane_off + c, o);
          }
        }
        l_run *= alpha;
        m_ref = m_new;
      }
      mark(2);
      // S(j+1) was issued two blocks ago
      if (PREF && j + 1 < nb) {
        const int sn = (j + 1) & 1;
        wait_bar(s_full_a + sn * 8, ((j + 1) >> 1) & 1, 0, 9, j + 1);
        tc_fence_after();
        const uint32_t t_n = tmem_base + lane_off + sn * BKV;
        tmem_ld32(t_n, nxt);
        if constexpr (BKV == 64) tmem_ld32(t_n + 32, nxt + 32);
      }
      mark(0);
      // the P buffer is free once the P V that read it last has completed: P V(j-2) with two buffers, else P V(j-1)
      // (which is issued ~300 clocks after the last warp delivered P(j-1))
      if (p.np == 2) {
        if (j >= 2) wait_bar(pv_done_a + (j & 1) * 8, ((j - 2) >> 1) & 1, 0, 10, j);
      } else if (j >= 1) {
        wait_bar(pv_done_a + ((j - 1) & 1) * 8, ((j - 1) >> 1) & 1, 0, 11, j);
      }
      tc_fence_after();
      mark(4);
      // P = exp2(s * scale - m_ref): one FFMA2 + two EX2 + one pack + one FADD2 per pair; stored 16 keys at a time
      const uint64_t nm2 = f2_pack(-m_ref, -m_ref);
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
      const uint32_t t_p = tmem_p0 + lane_off + (j & 1) * p_stride;
#pragma unroll
      for (int c0 = 0; c0 < BKV; c0 += 16) {
        uint32_t pk[8];
#pragma unroll
        for (int c = c0; c < c0 + 16; c += 2) {
          const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(cur[c]), __uint_as_float(cur[c + 1])), sc2, nm2);
          float x0, x1;
          f2_unpack(x2, x0, x1);
          const float e0 = ex2_approx(x0), e1 = ex2_approx(x1);
          pk[(c - c0) >> 1] = pack_bf16x2(e0, e1);
          acc2[(c >> 1) & 3] = f2_add(acc2[(c >> 1) & 3], f2_pack(e0, e1));
        }
        tmem_st8(t_p + (c0 >> 1), pk);
      }
      mark(3);
      tmem_st_wait();
      tmem_ld_wait();                                // S(j+1) is in `nxt`: its buffer may be overwritten
      mark(5);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[(j + (PREF ? 1 : 0)) & 1]);   // event j
      mark(6);
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(acc2[0], acc2[1]), a0, a1);
        f2_unpack(f2_add(acc2[2], acc2[3]), b0, b1);
        l_run += (a0 + a1) + (b0 + b1);
      }
    };
    if constexpr (PREF) {
      uint32_t sa[BKV], sb2[BKV];
      wait_bar(s_full_a, 0, 0, 12, 0);
      tc_fence_after();
      tmem_ld32(tmem_base + lane_off, sa);
      if constexpr (BKV == 64) tmem_ld32(tmem_base + lane_off + 32, sa + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[0]);          // event -1: S(0) is in registers
      for (int j = 0; j < nb; j += 2) {
        block(j, sa, sb2);
        if (j + 1 < nb) block(j + 1, sb2, sa);
      }
    } else {
      uint32_t sa[BKV];
      for (int j = 0; j < nb; j++) block(j, sa, sa);
    }
    if (TIMED && lane == 0 && p.dbg != nullptr) {
      long long* o = p.dbg + (((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 5 + (warp - 2)) * 8;
      for (int i = 0; i < 8; i++) o[i] = ph[i];
    }
    // ------------------------------- epilogue: O / l -> global -------------------------------
    wait_bar(pv_done_a + ((nb - 1) & 1) * 8, ((nb - 1) >> 1) & 1, 0, 13, nb);
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
// Keys per block.  Two S buffers + O must fit the CTA's tensor-memory allocation (a power of two): with 32-key blocks a
// head_dim <= 64 CTA needs 2 * 32 + dk16 <= 128 columns, so FOUR CTAs share an SM (16 softmax warps, four per scheduler)
// instead of two.  The softmax warps are latency-bound chains (tcgen05.ld -> max -> exp -> tcgen05.st); with two per
// scheduler the MUFU sat idle half of the time (ncu r2: XU 56 %), with four it is the unit that bounds the kernel.
static int g_force_bkv = 0;
static long long* g_attn_dbg = nullptr;
static int attention_bkv(int head_dim) {
  static const char* e = getenv("ONEDC_ATTN_BKV");
  const int want = g_force_bkv ? g_force_bkv : (e != nullptr ? atoi(e) : 0);
  if (head_dim > 64) return kBkvWide;             // the 32-key instance exists for one 64-channel chunk only
  if (want == 32 || want == 64) return want;
  return kBkvWide;
}
static void attention_plan(int batch, int heads, int head_dim, int sq, int skv, int* ks) {
  (void)batch; (void)heads; (void)sq;
  const int BKV = attention_bkv(head_dim);
  const int nblk = (skv + BKV - 1) / BKV;
  static const char* e_ks = getenv("ONEDC_ATTN_KVSPLIT");
  int want_ks = g_force_ks ? g_force_ks : (e_ks != nullptr ? e_ks[0] - '0' : 1);
  if (want_ks < 1 || want_ks > 4) want_ks = 1;
  while (want_ks > 1 && (long long)(want_ks - 1) * ((nblk + want_ks - 1) / want_ks) >= nblk) want_ks--;   // every split needs work
  *ks = want_ks;
}
}  // namespace onedc

using namespace onedc;

extern "C" void onedc_attention_set_plan(int32_t key_block, int32_t kv_splits) {
  g_force_bkv = (key_block == 32 || key_block == 64) ? key_block : 0;
  g_force_ks = kv_splits;
}

extern "C" void onedc_attention_set_debug(void* dev_counters) { g_attn_dbg = reinterpret_cast<long long*>(dev_counters); }

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
  const int BKV = attention_bkv(head_dim);
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
  const int tm_need = 2 * BKV + p.dk16 + BKV / 2;          // two S buffers, O, one P buffer (bf16 pairs)
  p.tmem_cols = tm_need <= 128 ? 128 : tm_need <= 256 ? 256 : 512;
  p.np = tm_need + BKV / 2 <= p.tmem_cols ? 2 : 1;
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
    uint32_t box[4] = {64, (uint32_t)BKV, 1, 1};
    int rc = make_tensor_map(&mk, k, 4, dims, str, box);
    if (rc) return rc;
    rc = make_tensor_map(&mv, v, 4, dims, str, box);
    if (rc) return rc;
  }
  // K / V ring depth: four stages for head_dim <= 64 (4 or 8 KB tiles); two where a tile is 16 / 24 KB (head_dim 80 / 160)
  // and four stages would cost the second CTA per SM
  const int stages = p.nchunk == 1 ? 4 : 2;
  const size_t smem = (size_t)p.nchunk * 16384 + 2 * stages * (size_t)p.nchunk * BKV * 128 + 1024;
  static size_t attr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  p.dbg = g_attn_dbg;
  const int variant = BKV == 32 ? 0 : p.nchunk;            // 0: 32-key blocks; 1 / 2 / 3: 64-key blocks, 64-channel chunks
  const int ai = variant + (p.dbg != nullptr ? 4 : 0);
  auto kern = variant == 0 ? attention_tc_kernel<32, 4, false, 4, 1, false>
            : variant == 1 ? attention_tc_kernel<64, 2, false, 4, 1, true>
            : variant == 2 ? attention_tc_kernel<64, 2, false, 2, 2, true> : attention_tc_kernel<64, 2, false, 2, 3, true>;
  if (p.dbg != nullptr)
    kern = variant == 0 ? attention_tc_kernel<32, 4, true, 4, 1, false>
         : variant == 1 ? attention_tc_kernel<64, 2, true, 4, 1, true>
         : variant == 2 ? attention_tc_kernel<64, 2, true, 2, 2, true> : attention_tc_kernel<64, 2, true, 2, 3, true>;
  if (smem > attr[ai]) {
    ONEDC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[ai] = smem;
  }
  dim3 grid((sq + 127) / 128, heads, batch * p.kv_splits);
  ONEDC_CUDA(launch_k(kern, grid, kAttnThreads, smem, st, mq, mk, mv, p));
  if (p.kv_splits > 1) {
    const long long total = (long long)batch * sq * heads;
    ONEDC_CUDA(launch_k(attention_merge_kernel, (int)((total + 127) / 128), 128, 0, st, (const float*)p.ws_o, (const float*)p.ws_ml,
                        (__nv_bfloat16*)out, o_ld, batch, heads, head_dim, sq, p.kv_splits));
  }
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
