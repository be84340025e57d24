/* onedc_b200 -- C ABI of the B200-native OneDC decode hot path.
 *
 * Plain C: pointers, sizes, PODs.  Every device pointer is CALLER-OWNED device memory
 * (the Python host side allocates it with torch); every call enqueues on the caller's
 * cudaStream_t (passed as void*) and returns 0, or a negative code with a message
 * retrievable through onedc_last_error().  No allocation happens on the hot path.
 *
 * What each group replaces in the reference (paths relative to /root/reference/src):
 *
 *  [entropy-index / dequant kernels]
 *    onedc_scale_to_index      compression_model.py:381 (scales*mask -> combine_for_writing :296-301)
 *                              + GaussianEncoder.build_indexes entropy_models.py:355-362
 *                              + .to(int16) entropy_models.py:86
 *    onedc_build_indexes       GaussianEncoder.build_indexes entropy_models.py:355-362 (generic tensor)
 *    onedc_dequant_accum       compression_model.py:383-384 (cat x4 + means, * mask, accumulate)
 *                              and the z-only variant compression_model.py:410-418
 *    onedc_quantize_residual   process_with_mask compression_model.py:224-239 (encode-side twin)
 *    onedc_fsq_codes           FSQ.indices_to_codes, call site codec_module.py:431
 *  [host entropy coder -- replaces the pybind11 module MLCodec_rans / MLCodec_CXX]
 *    onedc_rans_*              cpp/py_rans/py_rans.cpp:261-281 (RansDecoder/RansEncoder bindings),
 *                              cpp/rans/rans.cpp:101-187,303-362
 *    onedc_pmf_to_quantized_cdf cpp/ops/ops.cpp:84-91
 *    (ctypes releases the GIL around these, unlike the reference binding py_rans.cpp:183)
 *  [network kernels -- replace the PyTorch-eager cuDNN/cuBLAS/SDPA calls of the path]
 *    onedc_igemm               every Conv2d 1x1/3x3 (s1/s2) and Linear on the path (dcvc.py:246-250,
 *                              358-359,121,196; vqgan/blocks.py:29-32,61-80; diffusers ResnetBlock2D,
 *                              Transformer2DModel, Attention projections, GEGLU; AutoencoderKL decoder),
 *                              with bias / LeakyReLU / SiLU / GEGLU / ConvFFN3 split-sum / residual /
 *                              PixelShuffle / transposed store fused in the epilogue and channel concat
 *                              folded into the K loop
 *    onedc_attention           F.scaled_dot_product_attention in diffusers AttnProcessor2_0 (UNet self-
 *                              and cross-attention)
 *    onedc_groupnorm_*         nn.GroupNorm(32) (+SiLU) vqgan/blocks.py:28,31,60; decoder_unet.py:20-24
 *    onedc_layernorm           nn.LayerNorm in diffusers BasicTransformerBlock
 *    onedc_softmax_rows        softmax in vqgan/blocks.py:97 and the VAE mid-block attention
 *    onedc_dwconv3x3           depthwise conv dcvc.py:249
 *    onedc_upsample2x          F.interpolate(nearest, 2x) in diffusers Upsample2D
 *    onedc_x0_prepare          get_x0_from_noise modules/dmd/utils.py:279-284 + 1/0.18215
 *                              (model_sd15_with_codec_stage1.py:186) + post_quant_conv
 *    onedc_window_partition / onedc_window_merge   autoencoders_patch_attn.py:20-29
 */
#ifndef ONEDC_B200_H_
#define ONEDC_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes */
#define ONEDC_BF16 0
#define ONEDC_F32 1
/* activation codes */
#define ONEDC_ACT_NONE 0
#define ONEDC_ACT_LRELU 1
#define ONEDC_ACT_SILU 2
#define ONEDC_ACT_GELU 3
/* epilogue modes */
#define ONEDC_EPI_PLAIN 0
#define ONEDC_EPI_PAIR_LRELU 1 /* out[j] = lrelu(x[j],.1) + lrelu(x[j+BN/2],.01)  (ConvFFN3) */
#define ONEDC_EPI_GEGLU 2      /* out[j] = x[j] * gelu(x[j+BN/2]) */
/* store modes */
#define ONEDC_ST_NORMAL 0
#define ONEDC_ST_PIXSHUF 1    /* PixelShuffle(2): GEMM column q*ps_c + c -> pixel (2y+(q>>1), 2x+(q&1)), channel c */
#define ONEDC_ST_TRANSPOSED 2 /* out[(img*ncols + col)*out_ld + pixel]  (NCHW / V^T) */
#define ONEDC_ST_QUAD 3       /* output pixel (2y + (quad>>1), 2x + (quad&1)) of a 2H x 2W image: one phase of a
                                 nearest-2x-upsample + 3x3 conv folded into four 2x2 convs on the low-res input */

const char* onedc_last_error(void);
int onedc_version(void);
/* number of kernels launched by this library in this process since the last reset */
int64_t onedc_launch_count(int reset);
/* programmatic dependent launch of every kernel of this library (default off: measured 3 % slower
   inside the graph-replayed step on B200; env ONEDC_PDL=1 turns it on).
   Returns the previous setting.  Takes effect for launches (and graph captures) made after the call. */
int onedc_set_pdl(int on);

/* ---- implicit-GEMM convolution / GEMM on tcgen05 ------------------------------------------ */
typedef struct {
  const void* a_ptr[2]; /* bf16 NHWC activations; second source is concatenated along channels */
  int32_t a_c[2];       /* channels per source (second may be 0) */
  int64_t a_pix_stride[2]; /* elements between pixels (>= channels; allows channel-sliced views) */
  int32_t n_img, h_in, w_in;
  int32_t ksize;        /* 1 or 3 (pad 1) */
  int32_t stride;       /* 1 or 2 */
  const void* w_ptr;    /* bf16 [taps][cout][ktot], ktot contiguous */
  int32_t cout, ktot;
  int64_t w_row_stride, w_z_stride; /* elements */
  int32_t w_batched;    /* 1: one weight matrix per image (batched GEMM, taps must be 1) */
  const float* bias;    /* fp32 [cout] or NULL */
  int32_t epi_mode, act;
  float slope;
  const void* res;      /* residual added after the activation, [pixels, res_ld], or NULL */
  int32_t res_dtype;
  int64_t res_ld;
  void* out;
  int32_t out_dtype;
  int64_t out_ld;
  int32_t out_col_off;
  int32_t store_mode;
  int32_t ps_c;
  int32_t bn;           /* N tile, multiple of 16 (32 for pair modes), <= 256; 0 = auto */
  int32_t impl;         /* 0 = tcgen05 kernel, 1 = SIMT checking kernel (debug only), 2 = dry run: validate and
                           decide (tiling, split-K, fused statistics) without launching (bench bookkeeping) */
  /* split-K switch (layers with too few tiles to fill the GPU): non-NULL splitk_ws / splitk_counters allow it, NULL
   * disables it.  Since round 2 the splits of a tile run as one thread-block cluster and reduce over distributed
   * shared memory: the buffers are not touched any more (kept in the ABI; splitk_ws_floats still bounds tiles x splits) */
  void* splitk_ws;
  int64_t splitk_ws_floats;
  void* splitk_counters;
  int32_t splitk_max_tiles; /* number of counters */
  /* optional custom filter taps (stride 1): ntaps > 0 overrides ksize; tap t reads input pixel
   * (y + tap_dy[t], x + tap_dx[t]) with weights w_ptr[t]; `quad` selects the ONEDC_ST_QUAD phase */
  int32_t ntaps;
  int32_t tap_dy[9], tap_dx[9];
  int32_t quad;
  /* optional fused GroupNorm statistics of the stored output: gn_acc = device fp64 [n_img][gn_groups][2] (sum, sum
   * of squares per group of cout/gn_groups consecutive channels), zero-initialised by the caller, accumulated with
   * atomics.  Only done when the whole launch runs on the vector-store path without split-K and the group size is
   * 4, 8, 16 or 32 channels; gn_fused_out (written by the call) says whether it was. */
  void* gn_acc;
  int32_t gn_groups;
  int32_t gn_fused_out;
  /* optional L2 prefetch hint: a device range (16-byte aligned, size a multiple of 16) that a later launch will read
   * -- in practice the NEXT layer's weights, which would otherwise arrive from DRAM at the head of that launch; the
   * CTAs issue cp.async.bulk.prefetch.L2 for disjoint slices of it before starting their own work */
  const void* prefetch_ptr;
  int64_t prefetch_bytes;
  /* 1 = the launch plan must not depend on the batch size or the SM count: no split-K, no column-copy / transposed
   * tile, so every output element is accumulated in the same (tap, 64-channel chunk) order whatever n_img is.  Set
   * for every layer that feeds the entropy parameters (hyper-synthesis, prior nets): encoder and decoder must see
   * bit-identical scales (compression_model.py:369-407), also when one of them batches. */
  int32_t deterministic;
  /* 1 = w_ptr holds one more [cout][ktot] matrix behind its taps, the identity (cout == channels of `res`).  The kernel
   * may then add a bf16 residual on the tensor core: the residual tensor becomes one more K chunk group of the centre
   * pixel multiplied by that identity (exact: bf16 x 1.0 accumulated in fp32), loaded by TMA in the main loop instead of
   * by 2-byte loads in the epilogue.  Used for the transposed tile (cout <= 128) with no activation. */
  int32_t w_identity_tap;
} onedc_igemm_desc;

/* diagnostics: when set (device pointer to 148 x 16 int64, zeroed by the caller) every igemm CTA records clock
   counters per warp role: [0] producer total, [1] producer waiting for a free A slot, [2] ... for a free B slot,
   [4] MMA issuer total, [5] waiting for A data, [6] waiting for B data, [7] waiting for a free accumulator,
   [8] epilogue total, [9] waiting for a finished accumulator, [10] (split-K) waiting for the other splits.
   NULL turns it off. */
void onedc_igemm_set_debug(void* dev_counters);
int onedc_igemm(onedc_igemm_desc* d, void* stream);

/* ---- flash attention on tcgen05 (multi-head, head_dim 40/80/160) -------------------------- */
/* q: [batch, sq, q_ld] bf16 (head h at columns h*d), k/v: [batch, skv, kv_ld], out: [batch, sq, o_ld] */
/* ws: fp32 scratch of onedc_attention_ws_floats() elements (0 = none needed) for launches whose key range is split
   over several CTAs per query tile to fill the last wave; NULL disables the split */
/* overrides the launch plan: s_buffers 1 = one S buffer / three CTAs per SM (head_dim <= 64), 2 = two S buffers / two
   CTAs per SM; kv_splits 1..4; 0 = default (2, 1: measured fastest on B200, see attention.cu) */
void onedc_attention_set_plan(int32_t s_buffers, int32_t kv_splits);
int64_t onedc_attention_ws_floats(int32_t batch, int32_t heads, int32_t head_dim, int32_t sq, int32_t skv);
int onedc_attention(const void* q, int64_t q_ld, const void* k, const void* v, int64_t kv_ld, void* out,
                    int64_t o_ld, int32_t batch, int32_t heads, int32_t head_dim, int32_t sq, int32_t skv,
                    float scale, int32_t impl, float* ws, int64_t ws_floats, void* stream);

/* ---- normalisation / elementwise --------------------------------------------------------- */
/* GroupNorm over the channel concatenation of up to two NHWC sources. partial: fp32 scratch of
 * onedc_groupnorm_ws_floats() floats (per-block group sums); counters: n_img uint32, zero on entry and zero again
 * on exit (the last block of an image merges the block sums in a fixed order); stats: fp32 [n_img][groups][2]
 * (mean, rstd).  valid_px (device int32[n_img] or NULL): number of real pixels per image when the rest is zero
 * padding (edge windows of the VAE attention): the statistics then cover the real pixels only. */
int64_t onedc_groupnorm_ws_floats(int32_t n_img, int64_t hw, int32_t c_total);
int onedc_groupnorm_stats(const void* x0, int32_t c0, int64_t ld0, const void* x1, int32_t c1, int64_t ld1,
                          int32_t in_dtype, int32_t n_img, int64_t hw, int32_t groups, float eps, float* partial,
                          float* stats, uint32_t* counters, const int32_t* valid_px, void* stream);
/* statistics come either from `stats` (onedc_groupnorm_stats) or, when acc0 != NULL, from the fp64 (sum, sum of
 * squares) accumulators that onedc_igemm fused into the producers' epilogues: per group [n_img][groups][2] (single
 * source, acc_per_channel = 0) or per channel [n_img][c0][2] / [n_img][c1][2] (acc_per_channel = 1; acc1 belongs to
 * the second source of a concatenation) */
int onedc_groupnorm_apply(const void* x0, int32_t c0, int64_t ld0, const void* x1, int32_t c1, int64_t ld1,
                          int32_t in_dtype, int32_t n_img, int64_t hw, int32_t groups, const float* stats,
                          const double* acc0, const double* acc1, int32_t acc_per_channel, float eps,
                          const float* gamma, const float* beta, int32_t silu, void* out, int64_t out_ld, void* stream);
int onedc_layernorm(const void* x, int64_t ld, int64_t rows, int32_t c, const float* gamma, const float* beta,
                    float eps, void* out, int64_t out_ld, void* stream);
/* scores fp32 [rows, ld] -> bf16 probabilities [rows, out_ld]; columns >= valid are written as 0 */
int onedc_softmax_rows(const float* scores, int64_t ld, int64_t rows, int32_t cols, int32_t valid, float scale,
                       void* out, int64_t out_ld, void* stream);
/* same, with a per-batch valid length (device int32 array), rows_per_batch rows share one entry */
int onedc_softmax_rows_batched(const float* scores, int64_t ld, int64_t rows, int32_t cols,
                               const int32_t* valid_per_batch, int32_t rows_per_batch, float scale, void* out,
                               int64_t out_ld, void* stream);
int onedc_dwconv3x3(const void* x, const float* w9c, const float* bias, void* out, int32_t n_img, int32_t h,
                    int32_t w, int32_t c, void* stream);
int onedc_upsample2x(const void* x, void* out, int32_t n_img, int32_t h, int32_t w, int32_t c, void* stream);
int onedc_window_partition(const void* x, void* out, int32_t n_img, int32_t h, int32_t w, int32_t c, int32_t win,
                           void* stream);
int onedc_window_merge(const void* attn_out, const void* residual, void* out, int32_t n_img, int32_t h, int32_t w,
                       int32_t c, int32_t win, void* stream);
/* x0 = (reduced - sqrt(1-a) eps)/sqrt(a) in fp32, then /0.18215 and post_quant_conv (4x4 + bias);
 * writes bf16 NHWC with 8 channels: [hi(4), lo(4)] split of the fp32 value; x0_out (fp32, optional) */
/* second half of a 3x3 convolution with 1..4 output channels run as a 1x1 GEMM with 9*cout tap-expanded columns:
   out[n,y,x,c] = bias[c] + res[n,y,x,c] + sum_t y[n, y+dy_t, x+dx_t, t*cout + c]; y fp32 NHWC with ldy columns, out fp32
   NHWC (row pitch out_ld) or planar [n][cout][h*w] */
int onedc_tap_gather(const float* y, int32_t ldy, int32_t cout, const float* bias, const float* res, int64_t res_ld,
                     float* out, int64_t out_ld, int32_t planar, int32_t n_img, int32_t h, int32_t w, void* stream);
int onedc_x0_prepare(const float* reduced, const float* eps, float sqrt_alpha, float sqrt_one_minus_alpha,
                     float inv_scaling, const float* pq_w, const float* pq_b, void* out_hilo, float* x0_out,
                     int64_t pixels, void* stream);

/* ---- entropy-index kernels --------------------------------------------------------------- */
/* lut: device uint8[65536] (bf16 bit pattern -> index); [lut_lo, lut_hi) = the positive bit patterns whose index is
 * neither 0 nor 255 (first pattern with index > 0, first pattern with index 255): the kernel keeps only that slice
 * in shared memory.  scales: NHWC [n,h,w,c4*4] with pixel stride ld.
 * idx_out: int16 [n][c4][h][w]  (the symbol order of the y stream). */
int onedc_scale_to_index(const void* scales, int64_t ld, const uint8_t* lut, int32_t lut_lo, int32_t lut_hi,
                         int16_t* idx_out, int32_t step, int32_t n_img, int32_t h, int32_t w, int32_t c4, void* stream);
/* generic build_indexes: in_dtype bf16 -> LUT, fp32 -> 255 ascending thresholds (device fp32[255]) */
int onedc_build_indexes(const void* scales, int32_t in_dtype, const uint8_t* lut, const float* thresholds,
                        int32_t* idx_out, int64_t n, void* stream);
/* y_hat[n,h,w,32g+c] = bf16(sym[n][c][h][w] + means[n,h,w,32g+c]) at this step's active positions
 * (sym == NULL: means only, the z-only model).  step 0 also zero-fills the inactive positions. */
int onedc_dequant_accum(const int16_t* sym, const void* means, int64_t means_ld, void* y_hat, int64_t y_ld,
                        int32_t step, int32_t n_img, int32_t h, int32_t w, int32_t c4, void* stream);
/* encode-side twin: sym[n][c][h][w] = clamp(round_half_even(bf16(y - means)), +-30000) at active positions and
 * y_hat as in dequant_accum */
int onedc_quantize_residual(const void* y, int64_t y_in_ld, const void* means, int64_t means_ld, int16_t* sym,
                            void* y_hat, int64_t y_ld, int32_t step, int32_t n_img, int32_t h, int32_t w,
                            int32_t c4, void* stream);
/* z indices (int32 [n,hz,wz]) -> bf16 NHWC codes, 8 channels (7 codes + one zero pad) */
int onedc_fsq_codes(const int32_t* idx, void* out, int64_t n, void* stream);

/* ---- host entropy coder (no GPU involved; thread-safe per handle) -------------------------- */
int onedc_pmf_to_quantized_cdf(const float* pmf, int32_t n, int32_t precision, int32_t* cdf_out);
typedef struct onedc_rans_tables onedc_rans_tables;
onedc_rans_tables* onedc_rans_tables_create(const int32_t* cdf, int32_t rows, int32_t row_stride,
                                            const int32_t* cdf_sizes, const int32_t* offsets);
void onedc_rans_tables_destroy(onedc_rans_tables* t);
typedef struct onedc_rans_decoder onedc_rans_decoder;
onedc_rans_decoder* onedc_rans_decoder_create(void);
void onedc_rans_decoder_destroy(onedc_rans_decoder* d);
/* stream = reference y stream (flag byte + [sizes] + payload); the bytes are copied */
int onedc_rans_decoder_set_stream(onedc_rans_decoder* d, const uint8_t* stream, size_t n);
int onedc_rans_decoder_decode(onedc_rans_decoder* d, const onedc_rans_tables* t, const int16_t* indexes,
                              int32_t n, int16_t* out);
typedef struct onedc_rans_encoder onedc_rans_encoder;
onedc_rans_encoder* onedc_rans_encoder_create(void);
void onedc_rans_encoder_destroy(onedc_rans_encoder* e);
void onedc_rans_encoder_reset(onedc_rans_encoder* e);
int onedc_rans_encoder_encode(onedc_rans_encoder* e, const onedc_rans_tables* t, const int16_t* symbols,
                              const int16_t* indexes, int32_t n);
/* flush and return the stream size (flag byte included); copy it out with _get_stream */
int64_t onedc_rans_encoder_flush(onedc_rans_encoder* e);
int onedc_rans_encoder_get_stream(const onedc_rans_encoder* e, uint8_t* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* ONEDC_B200_H_ */
