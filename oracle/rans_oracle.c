/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product
 * (onedc_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.
 *
 * Plain-C restatement of the reference's host entropy coder for the y stream:
 *   - pmf_to_quantized_cdf        /root/reference/src/cpp/ops/ops.cpp:24-82
 *   - byte-wise rANS primitives   /root/reference/src/cpp/rans/rans_byte.h:48-155
 *   - bypass bit put/get          /root/reference/src/cpp/rans/rans.cpp:35-71
 *   - encode_with_indexes + flush /root/reference/src/cpp/rans/rans.cpp:101-187
 *   - decode_stream               /root/reference/src/cpp/rans/rans.cpp:303-362
 *   - multi-stream header         /root/reference/src/cpp/py_rans/py_rans.cpp:91-136,150-181
 * Pinned against the reference itself (oracle/_ref, compiled from the reference
 * sources by oracle/Makefile) in tests/test_oracle_pinned.py and against
 * tests/golden/rans_*.bin.
 *
 * State: 32-bit, L = 2^23, 16-bit probability precision, 2-bit bypass groups.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16
#define BYPASS_BITS 2
#define BYPASS_MAX 3
#define RANS_L (1u << 23)

/* ops.cpp:24-82.  cdf must hold n+1 entries.  Returns 0 on success. */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  cdf[0] = 0;
  for (int i = 0; i < n; i++)
    cdf[i + 1] = (uint32_t)(roundf(pmf[i] * (float)(1 << precision)) + 0.5f);
  uint32_t total = 0;
  for (int i = 0; i <= n; i++) total += cdf[i];
  for (int i = 0; i <= n; i++)
    cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
  for (int i = 1; i <= n; i++) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; i++) {
    if (cdf[i] == cdf[i + 1]) {
      uint32_t best_freq = ~0u;
      int best_steal = -1;
      for (int j = 0; j < n; j++) {
        uint32_t freq = cdf[j + 1] - cdf[j];
        if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
      }
      if (best_steal < 0) return -1;
      if (best_steal < i) { for (int j = best_steal + 1; j <= i; j++) cdf[j]--; }
      else                { for (int j = i + 1; j <= best_steal; j++) cdf[j]++; }
    }
  }
  return 0;
}

/* ---- encoder: queue of (start, range) pairs, range==0 => bypass bits ---- */
typedef struct { uint16_t start, range; } sym_t;

typedef struct {
  sym_t *q; size_t n, cap;
} symq_t;

static void q_push(symq_t *s, uint16_t start, uint16_t range) {
  if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 4096; s->q = (sym_t *)realloc(s->q, s->cap * sizeof(sym_t)); }
  s->q[s->n].start = start; s->q[s->n].range = range; s->n++;
}

/* rans.cpp:101-160: append symbols (value-offset, escape coding) to the queue. */
static void queue_symbols(symq_t *s, const int16_t *symbols, const int16_t *indexes, int n,
                          const int32_t *cdf, int cdf_stride, const int32_t *cdf_sizes,
                          const int32_t *offsets) {
  for (int i = 0; i < n; i++) {
    int32_t ci = indexes[i];
    if (ci < 0) continue;
    const int32_t *row = cdf + (size_t)ci * cdf_stride;
    int32_t max_value = cdf_sizes[ci] - 2;
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw = 0;
    if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
    q_push(s, (uint16_t)row[value], (uint16_t)(row[value + 1] - row[value]));
    if (value == max_value) {
      int32_t nb = 0;
      while ((raw >> (nb * BYPASS_BITS)) != 0) nb++;
      int32_t v = nb;
      while (v >= BYPASS_MAX) { q_push(s, BYPASS_MAX, 0); v -= BYPASS_MAX; }
      q_push(s, (uint16_t)v, 0);
      for (int32_t j = 0; j < nb; j++) q_push(s, (uint16_t)((raw >> (j * BYPASS_BITS)) & BYPASS_MAX), 0);
    }
  }
}

/* rans.cpp:162-187 + rans_byte.h:63-107.  Unlike the reference (which sizes the
 * output at one byte per queued entry and can under-run it) this
 * restatement sizes the buffer safely; the bytes produced are identical whenever
 * the reference does not crash. */
static size_t flush_queue(const symq_t *s, uint8_t **out) {
  size_t cap = s->n * 4 + 16;
  uint8_t *buf = (uint8_t *)malloc(cap);
  uint8_t *end = buf + cap, *p = end;
  uint32_t x = RANS_L;
  for (size_t k = s->n; k-- > 0;) {
    sym_t sy = s->q[k];
    if (sy.range != 0) {
      uint32_t freq = sy.range, x_max = freq << 15;
      while (x >= x_max) { *--p = (uint8_t)(x & 0xff); x >>= 8; }
      x = ((x / freq) << PRECISION) + (x % freq) + sy.start;
    } else {
      uint32_t freq = 1u << (PRECISION - BYPASS_BITS), x_max = freq << 15;
      while (x >= x_max) { *--p = (uint8_t)(x & 0xff); x >>= 8; }
      x = (x << BYPASS_BITS) | sy.start;
    }
  }
  p -= 4;
  p[0] = (uint8_t)(x >> 0); p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
  size_t nbytes = (size_t)(end - p);
  *out = (uint8_t *)malloc(nbytes);
  memcpy(*out, p, nbytes);
  free(buf);
  return nbytes;
}

/* One-shot encoder for `ngroups` consecutive encode_with_indexes calls followed by
 * flush + get_encoded_stream with stream_part == 1 (flag byte 0x01,
 * py_rans.cpp:116-117).  Returns the stream length; *out is malloc'ed (free with
 * oracle_free). */
size_t oracle_rans_encode(const int16_t *const *symbols, const int16_t *const *indexes,
                          const int *counts, int ngroups, const int32_t *cdf, int cdf_stride,
                          const int32_t *cdf_sizes, const int32_t *offsets, uint8_t **out) {
  symq_t s = {0, 0, 0};
  for (int g = 0; g < ngroups; g++)
    queue_symbols(&s, symbols[g], indexes[g], counts[g], cdf, cdf_stride, cdf_sizes, offsets);
  uint8_t *payload;
  size_t n = flush_queue(&s, &payload);
  free(s.q);
  *out = (uint8_t *)malloc(n + 1);
  (*out)[0] = 0x01; /* ((1-1)<<4) | (perStreamHeader==2 ? 1 : 0) */
  memcpy(*out + 1, payload, n);
  free(payload);
  return n + 1;
}

void oracle_free(void *p) { free(p); }

/* ---- decoder ---- */
typedef struct {
  const uint8_t *ptr;
  uint32_t x;
} dec_t;

static uint32_t get_bits(dec_t *d, uint32_t nbits) { /* rans.cpp:55-71 */
  uint32_t x = d->x, val = x & ((1u << nbits) - 1);
  x >>= nbits;
  if (x < RANS_L) { x = (x << 8) | *d->ptr++; }
  d->x = x;
  return val;
}

/* Stateful decoder handle (one cursor, calls must arrive in stream order). */
void *oracle_rans_dec_open(const uint8_t *stream, size_t n) {
  /* py_rans.cpp:150-181 with numberOfStreams == 1: skip flag byte; rans_byte.h:115-127 */
  (void)n;
  dec_t *d = (dec_t *)malloc(sizeof(dec_t));
  const uint8_t *p = stream + 1;
  d->x = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  d->ptr = p + 4;
  return d;
}

void oracle_rans_dec_close(void *h) { free(h); }

/* rans.cpp:303-362 */
void oracle_rans_decode(void *h, const int16_t *indexes, int n, const int32_t *cdf, int cdf_stride,
                        const int32_t *cdf_sizes, const int32_t *offsets, int16_t *out) {
  dec_t *d = (dec_t *)h;
  for (int i = 0; i < n; i++) {
    int32_t ci = indexes[i];
    if (ci < 0) { out[i] = 0; continue; }
    const int32_t *row = cdf + (size_t)ci * cdf_stride;
    int32_t len = cdf_sizes[ci], max_value = len - 2;
    uint32_t cum = d->x & ((1u << PRECISION) - 1);
    int k = 0;
    while (k < len && !((uint32_t)row[k] > cum)) k++;   /* linear find_if */
    uint32_t s = (uint32_t)(k - 1);
    uint32_t start = (uint32_t)row[s], freq = (uint32_t)(row[s + 1] - row[s]);
    uint32_t x = d->x;
    x = freq * (x >> PRECISION) + (x & ((1u << PRECISION) - 1)) - start;
    while (x < RANS_L) x = (x << 8) | *d->ptr++;
    d->x = x;
    int32_t value = (int32_t)s;
    if (value == max_value) {
      int32_t val = (int32_t)get_bits(d, BYPASS_BITS), nb = val;
      while (val == BYPASS_MAX) { val = (int32_t)get_bits(d, BYPASS_BITS); nb += val; }
      int32_t raw = 0;
      for (int j = 0; j < nb; j++) { val = (int32_t)get_bits(d, BYPASS_BITS); raw |= val << (j * BYPASS_BITS); }
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
    }
    out[i] = (int16_t)(value + offsets[ci]);
  }
}
