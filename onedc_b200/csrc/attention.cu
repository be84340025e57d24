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
//
// The SIMT kernel at the bottom is the on-GPU checker (impl = 1), never used by the decode path.
#include "../../include/onedc_b200.h"
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
  int tmem_cols;   // 256 when 2*BKV + dk16 fits (two CTAs per SM), else 512
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full, pv_done;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, batch = blockIdx.z;
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
  const uint32_t tmem_o = tmem_base + 2 * BKV;     // S buffers at columns [0, 2*BKV), O after them

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&q_full, (uint32_t)q_bytes);
      for (int c = 0; c < p.nchunk; c++) tma_load_4d(sQ + c * 16384, &map_q, &q_full, c * 64, q0, head, batch);
      for (int j = 0; j < p.nblk; j++) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], (uint32_t)kv_bytes);
        for (int c = 0; c < p.nchunk; c++)
          tma_load_4d(sK + st * kv_bytes + c * BKV * 128, &map_k, &k_full[st], c * 64, j * BKV, head, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], (uint32_t)kv_bytes);
        for (int c = 0; c < p.nchunk; c++)
          tma_load_4d(sV + st * kv_bytes + c * BKV * 128, &map_v, &v_full[st], c * 64, j * BKV, head, batch);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_qk = umma_idesc_bf16(128, BKV, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, p.dk16, 0, 1);   // B = V is MN-major
      mbar_wait(&q_full, 0);
      auto issue_pv = [&](int j) {
        const int st = j & 1;
        mbar_wait(&p_full, j & 1);
        mbar_wait(&v_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + st * kv_bytes);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; kk++) {
          const uint64_t da = umma_smem_desc(pa + kk * 32, 16, 1024);
          // V tile: MN(d)-major, 64-wide d chunks BKV*128 bytes apart, 8 kv rows per 1024-byte atom
          const uint64_t db = umma_smem_desc(va + kk * 2048, BKV * 128, 1024);
          umma_bf16(tmem_o, da, db, idesc_pv, (j | kk) != 0);
        }
        umma_commit(&pv_done);
        umma_commit(&v_empty[st]);
      };
      for (int j = 0; j < p.nblk; j++) {
        const int st = j & 1;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * kv_bytes);
        const int ksteps = p.dk16 / 16;
        for (int kk = 0; kk < ksteps; kk++) {
          const int c = kk >> 2, w = kk & 3;
          const uint64_t da = umma_smem_desc(qa + c * 16384 + w * 32, 16, 1024);
          const uint64_t db = umma_smem_desc(ka + c * BKV * 128 + w * 32, 16, 1024);
          umma_bf16(tmem_base + st * BKV, da, db, idesc_qk, kk != 0);
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(p.nblk - 1);
    }
  } else {
    // ------------------------------- softmax / output warps -------------------------------
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    uint8_t* prow = sP + (row >> 3) * 1024 + (row & 7) * 128;
    for (int j = 0; j < p.nblk; j++) {
      const int st = j & 1;
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[BKV];
      tmem_ld32(tmem_base + lane_off + st * BKV, sr);
      tmem_ld32(tmem_base + lane_off + st * BKV + 32, sr + 32);
      tmem_ld_wait();
      const int nvalid = p.skv - j * BKV;            // columns >= nvalid are zero-filled padding keys
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
    mbar_wait(&pv_done, (p.nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int s = q0 + row;
    __nv_bfloat16* dst = p.out + ((long long)batch * p.sq + s) * p.o_ld + head * p.d;
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

using namespace onedc;

extern "C" int onedc_attention(const void* q, int64_t q_ld, const void* k, const void* v, int64_t kv_ld, void* out,
                               int64_t o_ld, int32_t batch, int32_t heads, int32_t head_dim, int32_t sq, int32_t skv,
                               float scale, int32_t impl, void* stream) {
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
  const size_t smem = (size_t)p.nchunk * 16384 + 4 * (size_t)p.nchunk * BKV * 128 + 16384 + 1024;
  static size_t attr = 0;
  if (smem > attr) {
    ONEDC_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((sq + 127) / 128, heads, batch);
  ONEDC_CUDA(launch_k(attention_tc_kernel, grid, kAttnThreads, smem, st, mq, mk, mv, p));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
