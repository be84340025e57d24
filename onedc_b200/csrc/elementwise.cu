// HBM-bound byte/elementwise kernels of the decode path: entropy-index computation, dequantisation,
// FSQ code expansion, depthwise 3x3, nearest upsample, attention window (un)partition, x0 combine.
#include "../../include/onedc_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace onedc {

// step k, pixel parity p = 2*(y&1) + (x&1): active 32-channel group = p ^ kXor[k]
// (get_mask_four_parts, reference compression_model.py:269-283)
__device__ __constant__ int kXor[4] = {0, 3, 2, 1};

// ------------------------------------------------------------------------------------------------
// scales (NHWC bf16, 4*c4 channels) -> int16 CDF indices in stream order [n][c4][h*w].
// Block = 128 consecutive pixels of one image, in four rounds of 32 pixels.  Phase 1: FOUR lanes share a pixel, each
// takes one 16-byte vector (8 channels) of the pixel's active c4-channel slice, so a warp-wide load touches 8 pixels x
// 64 contiguous bytes (thread-per-pixel touched 32 separate lines per load and was bound by the LSU, 47 % of HBM);
// table lookup; indices to shared memory.  Phase 2 (thread = plane quarter): 16-byte stores of 8 consecutive pixels
// of one plane.  The tile is swizzled (8-pixel group index XOR vector index) so that both phases are conflict-free.
// The lookup uses the compact table: only bf16 patterns in [lo, hi) have an index other than 0 / 255
// (lo = first pattern with index > 0, hi = first pattern with index 255), ~1.2 KB instead of 64 KB.
// 16-byte read-only load that asks L2 for 64 bytes only.  The active slice of a pixel is 64 contiguous bytes out of its 256 or
// 512; with the default 128-byte promotion every slice dragged its neighbour out of DRAM as well (ncu: 1074 MB read for
// 537 MB of scales).
__device__ __forceinline__ uint4 ldg_l2_64(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <int C4>
__global__ void __launch_bounds__(128) scale_to_index_kernel(const __nv_bfloat16* scales, long long ld, const uint8_t* lut,
                                                             int lo, int hi, int16_t* idx_out, int step, int h, int w) {
  static_assert(C4 == 32, "lane mapping assumes four 16-byte vectors per pixel");
  pdl_wait();
  __shared__ __align__(16) int16_t tile[C4][128 + 8];
  __shared__ uint8_t tab[2048];
  const long long hw = (long long)h * w;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * 128;
  const int span = hi - lo;
  const int v = threadIdx.x & 3, pq = threadIdx.x >> 2;
  // the four loads of this thread first (independent of the table)
  uint4 q[4];
  bool okp[4];
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const long long p = p0 + it * 32 + pq;
    okp[it] = p < hw;
    q[it] = make_uint4(0, 0, 0, 0);
    if (okp[it]) {
      const int y = (int)p / w, x = (int)p - y * w;      // h * w < 2^31 (checked by the host wrapper)
      const int g = (2 * (y & 1) + (x & 1)) ^ kXor[step];
      q[it] = ldg_l2_64(reinterpret_cast<const uint4*>(scales + ((long long)n * hw + p) * ld + g * C4) + v);
    }
  }
  for (int i = threadIdx.x; i < span; i += 128) tab[i] = __ldg(lut + lo + i);
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int pl = (it * 32 + pq) ^ (v << 3);                      // swizzled pixel slot
    const uint32_t wds[4] = {q[it].x, q[it].y, q[it].z, q[it].w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
#pragma unroll
      for (int hlf = 0; hlf < 2; hlf++) {
        const int bits = hlf ? (int)(wds[j] >> 16) : (int)(wds[j] & 0xFFFFu);   // sign bit set => >= 0x8000 => 0
        int r = 0;
        if (bits >= hi && bits <= 0x7F80) r = 255;          // NaN patterns (> 0x7F80) keep index 0 like the table
        else if (bits >= lo && bits < hi) r = tab[bits - lo];
        tile[v * 8 + 2 * j + hlf][pl] = (int16_t)r;
      }
    }
  }
  __syncthreads();
  if ((hw & 7) == 0 && p0 + 128 <= hw) {
    // thread = (plane c, 32-pixel quarter): 4 x 16-byte stores
    const int c = threadIdx.x >> 2, qx = (threadIdx.x & 3) * 32, sw = (c >> 3) & 3;
    const uint4* s4 = reinterpret_cast<const uint4*>(&tile[c][qx]);
    uint4* d4 = reinterpret_cast<uint4*>(idx_out + ((long long)n * C4 + c) * hw + p0 + qx);
#pragma unroll
    for (int i = 0; i < 4; i++) d4[i] = s4[i ^ sw];
  } else {
    const long long p = p0 + threadIdx.x;
    if (p < hw) {
      int16_t* dst = idx_out + (long long)n * C4 * hw + p;
#pragma unroll 8
      for (int c = 0; c < C4; c++) dst[(long long)c * hw] = tile[c][threadIdx.x ^ (((c >> 3) & 3) << 3)];
    }
  }
}

// generic elementwise build_indexes (int32 out).  bf16 input: the 64 KB pattern table is staged in shared memory once per block
// (persistent grid-stride blocks), then 8 scales (one 16-byte load) -> 8 table bytes -> two 16-byte stores per thread.
// (One scalar 2-byte load, one global table byte and one 4-byte store per thread reached 42 % of HBM.)
__global__ void __launch_bounds__(512) build_indexes_kernel(const void* scales, int dtype, const uint8_t* lut, const float* thr, int32_t* out,
                                                            long long n) {
  extern __shared__ __align__(16) uint8_t tab[];
  pdl_wait();
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  if (dtype == DT_BF16) {
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(tab)[i] = __ldg(reinterpret_cast<const uint4*>(lut) + i);
    __syncthreads();
    const bool vec = (reinterpret_cast<uintptr_t>(scales) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    const long long nv = vec ? n >> 3 : 0;
    for (long long i = tid; i < nv; i += nth) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(scales) + i);
      const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
      int r[8];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        r[2 * j] = tab[wd[j] & 0xFFFFu];
        r[2 * j + 1] = tab[wd[j] >> 16];
      }
      int4* o = reinterpret_cast<int4*>(out) + 2 * i;
      __stcs(o, make_int4(r[0], r[1], r[2], r[3]));
      __stcs(o + 1, make_int4(r[4], r[5], r[6], r[7]));
    }
    for (long long i = nv * 8 + tid; i < n; i += nth) out[i] = tab[reinterpret_cast<const uint16_t*>(scales)[i]];
    return;
  }
  for (long long i = tid; i < n; i += nth) {
    const float s = reinterpret_cast<const float*>(scales)[i];
    int lo = 0, hi = 255;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s >= __ldg(thr + mid)) lo = mid + 1; else hi = mid;
    }
    out[i] = lo;
  }
}

// ------------------------------------------------------------------------------------------------
// y_hat[n,y,x,g*C4+c] = bf16( sym[n][c][y*w+x] + means[n,y,x,g*C4+c] )   (sym may be null: means only)
// MODE 0: decode (sym given or null).  MODE 1: encode twin, sym is OUTPUT = clamp(rint(y - means)).
// Same thread mapping as scale_to_index_kernel: plane-major 16-byte accesses on the symbol side, four lanes per
// pixel (one 16-byte vector each) on the NHWC side, swizzled shared-memory tile in between.
template <int C4, int MODE>
__global__ void __launch_bounds__(128) dequant_kernel(int16_t* sym, const __nv_bfloat16* yin, long long yin_ld,
                                                      const __nv_bfloat16* means, long long means_ld, __nv_bfloat16* y_hat,
                                                      long long y_ld, int step, int h, int w) {
  static_assert(C4 == 32, "lane mapping assumes four 16-byte vectors per pixel");
  pdl_wait();
  __shared__ __align__(16) int16_t tile[C4][128 + 8];
  const long long hw = (long long)h * w;
  const int n = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * 128;
  const bool vec = ((hw & 7) == 0) && (p0 + 128 <= hw);
  const int v = threadIdx.x & 3, pq = threadIdx.x >> 2;
  // NHWC side first: the loads do not depend on the symbols
  uint4 mq[4], yq[4];
  bool okp[4];
  int gq[4];
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const long long p = p0 + it * 32 + pq;
    okp[it] = p < hw;
    mq[it] = yq[it] = make_uint4(0, 0, 0, 0);
    gq[it] = 0;
    if (okp[it]) {
      const int y = (int)p / w, x = (int)p - y * w;      // h * w < 2^31 (checked by the host wrapper)
      gq[it] = (2 * (y & 1) + (x & 1)) ^ kXor[step];
      const long long pix = (long long)n * hw + p;
      mq[it] = ldg_l2_64(reinterpret_cast<const uint4*>(means + pix * means_ld + gq[it] * C4) + v);
      if (MODE == 1) yq[it] = ldg_l2_64(reinterpret_cast<const uint4*>(yin + pix * yin_ld + gq[it] * C4) + v);
    }
  }
  if (MODE == 0 && sym != nullptr) {
    if (vec) {
      // thread = (plane c, 32-pixel quarter): 4 x 16-byte loads
      const int c = threadIdx.x >> 2, qx = (threadIdx.x & 3) * 32, sw = (c >> 3) & 3;
      const uint4* s4 = reinterpret_cast<const uint4*>(sym + ((long long)n * C4 + c) * hw + p0 + qx);
      uint4* d4 = reinterpret_cast<uint4*>(&tile[c][qx]);
#pragma unroll
      for (int i = 0; i < 4; i++) d4[i ^ sw] = __ldg(s4 + i);
    } else {
      const long long p = p0 + threadIdx.x;
      if (p < hw) {
        const int16_t* src = sym + (long long)n * C4 * hw + p;
#pragma unroll 8
        for (int c = 0; c < C4; c++) tile[c][threadIdx.x ^ (((c >> 3) & 3) << 3)] = src[(long long)c * hw];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int it = 0; it < 4; it++) {
    const int pl = (it * 32 + pq) ^ (v << 3);
    if (okp[it]) {
      const long long pix = (long long)n * hw + p0 + it * 32 + pq;
      const uint32_t mw[4] = {mq[it].x, mq[it].y, mq[it].z, mq[it].w};
      const uint32_t yw[4] = {yq[it].x, yq[it].y, yq[it].z, yq[it].w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float q0, q1;
        if (MODE == 1) {
          // process_with_mask (compression_model.py:224-239) under bf16 autocast: the residual y - means is a bf16
          // tensor BEFORE torch.round (round half to even) sees it
          const float r0 = __bfloat162float(__float2bfloat16(bf16lo(yw[j]) - bf16lo(mw[j])));
          const float r1 = __bfloat162float(__float2bfloat16(bf16hi(yw[j]) - bf16hi(mw[j])));
          q0 = fminf(fmaxf(rintf(r0), -30000.f), 30000.f);
          q1 = fminf(fmaxf(rintf(r1), -30000.f), 30000.f);
          tile[v * 8 + 2 * j][pl] = (int16_t)q0;
          tile[v * 8 + 2 * j + 1][pl] = (int16_t)q1;
        } else if (sym != nullptr) {
          q0 = (float)tile[v * 8 + 2 * j][pl];
          q1 = (float)tile[v * 8 + 2 * j + 1][pl];
        } else {
          q0 = q1 = 0.f;
        }
        // the reference casts the decoded symbols to the activation dtype (bf16) before adding the means
        // (entropy_models.py:374 `.to(dtype)`): |q| > 256 is rounded to bf16 first
        q0 = __bfloat162float(__float2bfloat16(q0));
        q1 = __bfloat162float(__float2bfloat16(q1));
        ow[j] = pack_bf16x2(q0 + bf16lo(mw[j]), q1 + bf16hi(mw[j]));
      }
      reinterpret_cast<uint4*>(y_hat + pix * y_ld + gq[it] * C4)[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      if (step == 0) {   // y_hat_so_far starts as (..)*mask_0: zero everywhere else
#pragma unroll
        for (int og = 0; og < 4; og++)
          if (og != gq[it]) reinterpret_cast<uint4*>(y_hat + pix * y_ld + og * C4)[v] = make_uint4(0, 0, 0, 0);
      }
    }
  }
  if (MODE == 1) {
    __syncthreads();
    if (vec) {
      const int c = threadIdx.x >> 2, qx = (threadIdx.x & 3) * 32, sw = (c >> 3) & 3;
      const uint4* s4 = reinterpret_cast<const uint4*>(&tile[c][qx]);
      uint4* d4 = reinterpret_cast<uint4*>(sym + ((long long)n * C4 + c) * hw + p0 + qx);
#pragma unroll
      for (int i = 0; i < 4; i++) d4[i] = s4[i ^ sw];
    } else {
      const long long p = p0 + threadIdx.x;
      if (p < hw) {
        int16_t* o = sym + (long long)n * C4 * hw + p;
#pragma unroll 8
        for (int c = 0; c < C4; c++) o[(long long)c * hw] = tile[c][threadIdx.x ^ (((c >> 3) & 3) << 3)];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void fsq_codes_kernel(const int32_t* idx, __nv_bfloat16* out, long long n) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = idx[i];
    float c[8];
#pragma unroll
    for (int d = 0; d < 7; d++) c[d] = (float)(((v >> (2 * d)) & 3) - 2) * 0.5f;
    c[7] = 0.f;
    uint4 q;
    q.x = pack_bf16x2(c[0], c[1]); q.y = pack_bf16x2(c[2], c[3]);
    q.z = pack_bf16x2(c[4], c[5]); q.w = pack_bf16x2(c[6], c[7]);
    reinterpret_cast<uint4*>(out)[i] = q;
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, pad 1, NHWC bf16; weights fp32 [9][C], bias fp32 [C].
// Thread = (4-channel vector, column x) of one strip of output rows: it keeps the 3 x 3 window of its column in registers
// and slides it down the strip, so every output costs 3 new 8-byte loads (row y+1 at x-1, x, x+1; two of the three are the
// neighbouring threads' centre columns and hit L1) instead of 9 loads + 18 weight loads, and the nine weight vectors are read
// once per strip.  4 channels per thread: nine fp32 weight vectors + the window fit 80 registers, three blocks per SM (with 8
// channels per thread the kernel needed 138 registers, ran one block per SM and reached 21 % of HBM; round 1's one thread
// per output vector with nine loads each: 18 %).  The strip height is chosen by the host: 16 rows for big tensors, 2..4 for
// the codec's 48 x 48 planes, which need the threads more than the reuse.
__global__ void __launch_bounds__(256, 3) dwconv3x3_kernel(const __nv_bfloat16* x, const float* w9c, const float* bias,
                                                           __nv_bfloat16* out, int n_img, int h, int w, int c, int kDwRows) {
  pdl_wait();
  const int nvec = c >> 2;
  const int strips = (h + kDwRows - 1) / kDwRows;
  const long long total = (long long)n_img * strips * w * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long t = i / nvec;
    const int xx = (int)(t % w);
    t /= w;
    const int strip = (int)(t % strips);
    const int n = (int)(t / strips);
    const int y0 = strip * kDwRows, y1 = y0 + kDwRows < h ? y0 + kDwRows : h;
    float4 wk[9];
#pragma unroll
    for (int k = 0; k < 9; k++) wk[k] = __ldg(reinterpret_cast<const float4*>(w9c + k * c) + v);
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bias) + v);
    const bool okl = xx > 0, okr = xx + 1 < w;
    auto load_row = [&](const int iy, uint2* r) {              // r[0..2] = columns x-1, x, x+1 of row iy (zeros outside)
      r[0] = r[1] = r[2] = make_uint2(0, 0);
      if (iy >= 0 && iy < h) {
        const uint2* rp = reinterpret_cast<const uint2*>(x + (((long long)n * h + iy) * w + xx) * c) + v;
        r[1] = __ldg(rp);
        if (okl) r[0] = __ldg(rp - nvec);
        if (okr) r[2] = __ldg(rp + nvec);
      }
    };
    // two output rows per iteration: rows yy-1 .. yy+2 in win[0..3], six loads issued together, two independent FMA chains
    uint2 win[4][3];
    load_row(y0 - 1, win[0]);
    load_row(y0, win[1]);
    for (int yy = y0; yy < y1; yy += 2) {
      load_row(yy + 1, win[2]);
      load_row(yy + 2 <= y1 ? yy + 2 : h, win[3]);            // row y1 is only read as the halo of row y1 - 1
      float a[2][4];
#pragma unroll
      for (int r = 0; r < 2; r++) {
        a[r][0] = bs.x; a[r][1] = bs.y; a[r][2] = bs.z; a[r][3] = bs.w;
      }
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const float4 wv = wk[ky * 3 + kx];
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const uint2 q = win[ky + r][kx];
            a[r][0] = fmaf(bf16lo(q.x), wv.x, a[r][0]);
            a[r][1] = fmaf(bf16hi(q.x), wv.y, a[r][1]);
            a[r][2] = fmaf(bf16lo(q.y), wv.z, a[r][2]);
            a[r][3] = fmaf(bf16hi(q.y), wv.w, a[r][3]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
        if (yy + r < y1) {
          uint2 o;
          o.x = pack_bf16x2(a[r][0], a[r][1]);
          o.y = pack_bf16x2(a[r][2], a[r][3]);
          reinterpret_cast<uint2*>(out + (((long long)n * h + yy + r) * w + xx) * c)[v] = o;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        win[0][kx] = win[2][kx];
        win[1][kx] = win[3][kx];
      }
    }
  }
}

// nearest 2x upsample, NHWC: thread = one INPUT vector, read once and stored to its four output pixels
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* x, uint4* out, int n_img, int h, int w, int nvec) {
  pdl_wait();
  const long long total = (long long)n_img * h * w * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long pix = i / nvec;
    const int xi = (int)(pix % w);
    const long long t = pix / w;
    const int yi = (int)(t % h);
    const int n = (int)(t / h);
    const uint4 q = __ldg(x + i);
    uint4* o = out + (((long long)n * 2 * h + 2 * yi) * (2 * w) + 2 * xi) * nvec + v;
    o[0] = q;
    o[nvec] = q;
    o[(long long)2 * w * nvec] = q;
    o[(long long)2 * w * nvec + nvec] = q;
  }
}

// windows of win x win pixels -> [n*nwy*nwx][win*win tokens][c]; tokens of an edge window are
// enumerated row-major over its valid vh x vw extent, the rest of the window's rows are zero.
__global__ void __launch_bounds__(256) window_partition_kernel(const uint4* x, uint4* out, int n_img, int h, int w, int nvec,
                                                               int win) {
  pdl_wait();
  const int nwy = (h + win - 1) / win, nwx = (w + win - 1) / win;
  const long long total = (long long)n_img * nwy * nwx * win * win * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long t = i / nvec;
    const int tok = (int)(t % (win * win));
    t /= (win * win);
    const int wx = (int)(t % nwx);
    t /= nwx;
    const int wy = (int)(t % nwy);
    const int n = (int)(t / nwy);
    const int vw = min(win, w - wx * win), vh = min(win, h - wy * win);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (tok < vw * vh) {
      const int ly = tok / vw, lx = tok - ly * vw;
      val = __ldg(x + (((long long)n * h + wy * win + ly) * w + wx * win + lx) * nvec + v);
    }
    out[i] = val;
  }
}

__global__ void __launch_bounds__(256) window_merge_kernel(const uint4* a, const uint4* res, uint4* out, int n_img, int h, int w,
                                                           int nvec, int win) {
  pdl_wait();
  const int nwy = (h + win - 1) / win, nwx = (w + win - 1) / win;
  const long long total = (long long)n_img * h * w * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long pix = i / nvec;
    const int xx = (int)(pix % w);
    const long long t = pix / w;
    const int yy = (int)(t % h);
    const int n = (int)(t / h);
    const int wy = yy / win, wx = xx / win;
    const int vw = min(win, w - wx * win);
    const int tok = (yy - wy * win) * vw + (xx - wx * win);
    uint4 q = __ldg(a + ((((long long)n * nwy + wy) * nwx + wx) * win * win + tok) * nvec + v);
    if (res != nullptr) {
      uint4 r = __ldg(res + i);
      q.x = pack_bf16x2(bf16lo(q.x) + bf16lo(r.x), bf16hi(q.x) + bf16hi(r.x));
      q.y = pack_bf16x2(bf16lo(q.y) + bf16lo(r.y), bf16hi(q.y) + bf16hi(r.y));
      q.z = pack_bf16x2(bf16lo(q.z) + bf16lo(r.z), bf16hi(q.z) + bf16hi(r.z));
      q.w = pack_bf16x2(bf16lo(q.w) + bf16lo(r.w), bf16hi(q.w) + bf16hi(r.w));
    }
    out[i] = q;
  }
}

struct PqW {
  float w[16];
  float b[4];
};
__global__ void x0_prepare_kernel(const float4* reduced, const float4* eps, float sa, float s1m, float inv_scaling, PqW pq,
                                  uint4* out_hilo, float4* x0_out, long long pixels) {
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const float4 r = reduced[i], e = eps[i];
    float x0[4] = {(r.x - s1m * e.x) / sa, (r.y - s1m * e.y) / sa, (r.z - s1m * e.z) / sa, (r.w - s1m * e.w) / sa};
    if (x0_out != nullptr) x0_out[i] = make_float4(x0[0], x0[1], x0[2], x0[3]);
    float z[4], hi[4], lo[4];
#pragma unroll
    for (int o = 0; o < 4; o++) {
      float acc = pq.b[o];
#pragma unroll
      for (int c = 0; c < 4; c++) acc += pq.w[o * 4 + c] * (x0[c] * inv_scaling);
      z[o] = acc;
      hi[o] = __bfloat162float(__float2bfloat16(acc));
      lo[o] = z[o] - hi[o];
    }
    uint4 q;
    q.x = pack_bf16x2(hi[0], hi[1]); q.y = pack_bf16x2(hi[2], hi[3]);
    q.z = pack_bf16x2(lo[0], lo[1]); q.w = pack_bf16x2(lo[2], lo[3]);
    out_hilo[i] = q;
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 convolutions with a handful of output channels (UNet conv_out / vae_reduction 320 -> 4, VAE conv_out 128 -> 3)
// waste the tensor core as an N = 16 implicit GEMM that re-reads the activations nine times (25 TFLOP/s).  They run
// instead as ONE 1x1 GEMM with N = 9 * cout "tap-expanded" columns (activations read once) followed by this gather:
//   out[n, y, x, c] = bias[c] + res[n, y, x, c] + sum_t Y[n, y + dy_t, x + dx_t, t * cout + c]      (zero padding)
// Y fp32 [n, h, w, ldy]; out fp32, NHWC with row pitch out_ld or planar [n][c][h*w].  Thread = output pixel.
template <int COUT>
__global__ void __launch_bounds__(256) tap_gather_kernel(const float* __restrict__ Y, int ldy, const float* __restrict__ bias,
                                                         const float* __restrict__ res, long long res_ld, float* __restrict__ out,
                                                         long long out_ld, int planar, int n_img, int h, int w) {
  pdl_wait();
  const long long hw = (long long)h * w, total = hw * n_img;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(i / hw);
    const int p = (int)(i - (long long)img * hw);
    const int y = p / w, x = p - y * w;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; c++) acc[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float* src = Y + ((long long)img * hw + (long long)yy * w + xx) * ldy + t * COUT;
#pragma unroll
        for (int c = 0; c < COUT; c++) acc[c] += __ldcg(src + c);
      }
    }
    if (res != nullptr) {
#pragma unroll
      for (int c = 0; c < COUT; c++) acc[c] += __ldg(res + i * res_ld + c);
    }
    if (planar) {
#pragma unroll
      for (int c = 0; c < COUT; c++) out[((long long)img * COUT + c) * hw + p] = acc[c];
    } else {
#pragma unroll
      for (int c = 0; c < COUT; c++) out[i * out_ld + c] = acc[c];
    }
  }
}

static inline int ew_blocks(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace onedc

using namespace onedc;

extern "C" int onedc_scale_to_index(const void* scales, int64_t ld, const uint8_t* lut, int32_t lut_lo, int32_t lut_hi,
                                    int16_t* idx_out, int32_t step, int32_t n_img, int32_t h, int32_t w, int32_t c4,
                                    void* stream) {
  ONEDC_CHECK(c4 == 32 && step >= 0 && step < 4 && ld % 8 == 0, "scale_to_index: c4 must be 32, step in 0..3");
  ONEDC_CHECK(lut_lo >= 0 && lut_hi >= lut_lo && lut_hi - lut_lo <= 2048 && lut_hi <= 0x8000,
              "scale_to_index: compact table range [lo, hi) must span at most 2048 positive bf16 patterns");
  ONEDC_CHECK(n_img <= 65535 && (long long)h * w < (1ll << 31), "scale_to_index: batch or plane too large");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 127) / 128), n_img);
  ONEDC_CUDA(launch_k(scale_to_index_kernel<32>, grid, 128, 0, (cudaStream_t)stream, (const __nv_bfloat16*)scales, ld, lut, lut_lo, lut_hi, idx_out,
                                                                     step, h, w));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_build_indexes(const void* scales, int32_t in_dtype, const uint8_t* lut, const float* thresholds,
                                   int32_t* idx_out, int64_t n, void* stream) {
  ONEDC_CHECK((in_dtype == DT_BF16 && lut) || (in_dtype == DT_F32 && thresholds), "build_indexes: missing table");
  static bool attr_done = false;
  if (!attr_done) {
    ONEDC_CUDA(cudaFuncSetAttribute(build_indexes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr_done = true;
  }
  ONEDC_CHECK(in_dtype != DT_BF16 || reinterpret_cast<uintptr_t>(lut) % 16 == 0, "build_indexes: table must be 16-byte aligned");
  int blocks = ew_blocks(n / 8 + 1, 512);
  if (in_dtype == DT_BF16 && blocks > 3 * sm_count()) blocks = 3 * sm_count();          // three 64 KB tables per SM, grid-stride
  ONEDC_CUDA(launch_k(build_indexes_kernel, blocks, 512, in_dtype == DT_BF16 ? 65536 : 0, (cudaStream_t)stream, scales, in_dtype, lut,
                      thresholds, idx_out, n));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_dequant_accum(const int16_t* sym, const void* means, int64_t means_ld, void* y_hat, int64_t y_ld,
                                   int32_t step, int32_t n_img, int32_t h, int32_t w, int32_t c4, void* stream) {
  ONEDC_CHECK(c4 == 32 && step >= 0 && step < 4 && means_ld % 8 == 0 && y_ld % 8 == 0 && (long long)h * w < (1ll << 31),
              "dequant_accum: bad arguments");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 127) / 128), n_img);
  ONEDC_CUDA(launch_k(dequant_kernel<32, 0>, grid, 128, 0, (cudaStream_t)stream, const_cast<int16_t*>(sym), nullptr, 0,
                                                               (const __nv_bfloat16*)means, means_ld,
                                                               (__nv_bfloat16*)y_hat, y_ld, step, h, w));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_quantize_residual(const void* y, int64_t y_in_ld, const void* means, int64_t means_ld, int16_t* sym,
                                       void* y_hat, int64_t y_ld, int32_t step, int32_t n_img, int32_t h, int32_t w,
                                       int32_t c4, void* stream) {
  ONEDC_CHECK(c4 == 32 && step >= 0 && step < 4 && means_ld % 8 == 0 && y_ld % 8 == 0 && y_in_ld % 8 == 0 &&
                  (long long)h * w < (1ll << 31),
              "quantize_residual: bad arguments");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)((hw + 127) / 128), n_img);
  ONEDC_CUDA(launch_k(dequant_kernel<32, 1>, grid, 128, 0, (cudaStream_t)stream, sym, (const __nv_bfloat16*)y, y_in_ld,
                                                               (const __nv_bfloat16*)means, means_ld,
                                                               (__nv_bfloat16*)y_hat, y_ld, step, h, w));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_fsq_codes(const int32_t* idx, void* out, int64_t n, void* stream) {
  ONEDC_CUDA(launch_k(fsq_codes_kernel, ew_blocks(n, 128), 128, 0, (cudaStream_t)stream, idx, (__nv_bfloat16*)out, n));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_dwconv3x3(const void* x, const float* w9c, const float* bias, void* out, int32_t n_img, int32_t h,
                               int32_t w, int32_t c, void* stream) {
  ONEDC_CHECK(c % 8 == 0, "dwconv3x3: C must be a multiple of 8");
  ONEDC_CHECK(reinterpret_cast<uintptr_t>(w9c) % 16 == 0 && reinterpret_cast<uintptr_t>(bias) % 16 == 0, "dwconv3x3: weights / bias must be 16-byte aligned");
  // strip height: halve it until the launch has a few blocks per SM (or the strips are 2 rows)
  int rows = 16;
  while (rows > 2 && (long long)n_img * ((h + rows - 1) / rows) * w * (c / 4) < (long long)sm_count() * 1024) rows >>= 1;
  const long long total = (long long)n_img * ((h + rows - 1) / rows) * w * (c / 4);
  ONEDC_CUDA(launch_k(dwconv3x3_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, w9c, bias,
                                                                           (__nv_bfloat16*)out, n_img, h, w, c, rows));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_upsample2x(const void* x, void* out, int32_t n_img, int32_t h, int32_t w, int32_t c, void* stream) {
  ONEDC_CHECK(c % 8 == 0, "upsample2x: C must be a multiple of 8");
  const long long total = (long long)n_img * h * w * (c / 8);
  ONEDC_CUDA(launch_k(upsample2x_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)out, n_img, h, w, c / 8));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_window_partition(const void* x, void* out, int32_t n_img, int32_t h, int32_t w, int32_t c, int32_t win,
                                      void* stream) {
  ONEDC_CHECK(c % 8 == 0, "window_partition: C must be a multiple of 8");
  const int nwy = (h + win - 1) / win, nwx = (w + win - 1) / win;
  const long long total = (long long)n_img * nwy * nwx * win * win * (c / 8);
  ONEDC_CUDA(launch_k(window_partition_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)out, n_img, h, w,
                                                                                  c / 8, win));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_window_merge(const void* attn_out, const void* residual, void* out, int32_t n_img, int32_t h, int32_t w,
                                  int32_t c, int32_t win, void* stream) {
  ONEDC_CHECK(c % 8 == 0, "window_merge: C must be a multiple of 8");
  const long long total = (long long)n_img * h * w * (c / 8);
  ONEDC_CUDA(launch_k(window_merge_kernel, ew_blocks(total, 256), 256, 0, (cudaStream_t)stream, (const uint4*)attn_out, (const uint4*)residual,
                                                                              (uint4*)out, n_img, h, w, c / 8, win));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_x0_prepare(const float* reduced, const float* eps, float sqrt_alpha, float sqrt_one_minus_alpha,
                                float inv_scaling, const float* pq_w, const float* pq_b, void* out_hilo, float* x0_out,
                                int64_t pixels, void* stream) {
  PqW pq;
  for (int i = 0; i < 16; i++) pq.w[i] = pq_w[i];   // host pointers: 4x4 weight + bias of post_quant_conv
  for (int i = 0; i < 4; i++) pq.b[i] = pq_b[i];
  ONEDC_CUDA(launch_k(x0_prepare_kernel, ew_blocks(pixels, 256), 256, 0, (cudaStream_t)stream, (const float4*)reduced, (const float4*)eps,
                                                                             sqrt_alpha, sqrt_one_minus_alpha, inv_scaling,
                                                                             pq, (uint4*)out_hilo, (float4*)x0_out, pixels));
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int onedc_tap_gather(const float* y, int32_t ldy, int32_t cout, const float* bias, const float* res, int64_t res_ld,
                                float* out, int64_t out_ld, int32_t planar, int32_t n_img, int32_t h, int32_t w, void* stream) {
  ONEDC_CHECK(cout >= 1 && cout <= 4 && ldy >= 9 * cout && (long long)h * w < (1ll << 31), "tap_gather: cout must be 1..4");
  const long long total = (long long)n_img * h * w;
  const int blocks = ew_blocks(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (cout) {
    case 1: ONEDC_CUDA(launch_k(tap_gather_kernel<1>, blocks, 256, 0, st, y, ldy, bias, res, res_ld, out, out_ld, planar, n_img, h, w)); break;
    case 2: ONEDC_CUDA(launch_k(tap_gather_kernel<2>, blocks, 256, 0, st, y, ldy, bias, res, res_ld, out, out_ld, planar, n_img, h, w)); break;
    case 3: ONEDC_CUDA(launch_k(tap_gather_kernel<3>, blocks, 256, 0, st, y, ldy, bias, res, res_ld, out, out_ld, planar, n_img, h, w)); break;
    default: ONEDC_CUDA(launch_k(tap_gather_kernel<4>, blocks, 256, 0, st, y, ldy, bias, res, res_ld, out, out_ld, planar, n_img, h, w)); break;
  }
  ONEDC_CUDA(cudaGetLastError());
  return 0;
}
