// Host entropy coder of the y stream: byte-wise rANS (32-bit state, L = 2^23, 16-bit probabilities,
// 2-bit bypass groups for escaped values), bit-compatible with the stream format of the reference's
// MLCodec_rans module (cpp/rans/rans.cpp, cpp/py_rans/py_rans.cpp) but written for this library:
//   * C ABI, no Python objects, no per-call vector copies -> callable from any thread without the GIL
//   * symbol lookup through a per-row 256-bucket start table instead of a linear scan from 0 (a wider per-bucket
//     record holding start / frequency / value was tried and measured no faster on the B200 host: 7.9 vs 7.6 ns/symbol)
//   * the encoder sizes its output buffer from the actual code length (the reference allocates one
//     byte per queued entry and can under-run it)
// Also: pmf -> 16-bit quantised CDF (the MLCodec_CXX.pmf_to_quantized_cdf replacement).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/onedc_b200.h"

namespace {
constexpr int kPrecision = 16;
constexpr uint32_t kMask = (1u << kPrecision) - 1;
constexpr uint32_t kL = 1u << 23;
constexpr int kBypassBits = 2;
constexpr int kBypassMax = 3;
}  // namespace

struct onedc_rans_tables {
  int rows, stride;
  std::vector<int32_t> cdf;      // [rows][stride]
  std::vector<int32_t> sizes;    // entries per row (symbols + 2)
  std::vector<int32_t> offsets;
  std::vector<uint8_t> bucket;   // [rows][256]: largest s with cdf[s] <= (b << 8)
};

struct Cursor {
  const uint8_t* ptr;
  const uint8_t* end;
  uint32_t x;
};

struct onedc_rans_decoder {
  std::vector<uint8_t> data;
  std::vector<Cursor> cur;
};

struct onedc_rans_encoder {
  struct Sym { uint16_t start, range; };
  std::vector<Sym> q;
  std::vector<uint8_t> stream;
};

extern "C" int onedc_pmf_to_quantized_cdf(const float* pmf, int32_t n, int32_t precision, int32_t* out) {
  std::vector<uint32_t> cdf((size_t)n + 1);
  cdf[0] = 0;
  for (int i = 0; i < n; i++) cdf[i + 1] = (uint32_t)(roundf(pmf[i] * (float)(1 << precision)) + 0.5f);
  uint32_t total = 0;
  for (uint32_t v : cdf) total += v;
  if (total == 0) return -1;
  for (auto& v : cdf) v = (uint32_t)((((uint64_t)1 << precision) * v) / total);
  for (int i = 1; i <= n; i++) cdf[i] += cdf[i - 1];
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; i++) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best_freq = ~0u;
    int best = -1;
    for (int j = 0; j < n; j++) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best_freq) { best_freq = f; best = j; }
    }
    if (best < 0) return -1;
    if (best < i) { for (int j = best + 1; j <= i; j++) cdf[j]--; }
    else          { for (int j = i + 1; j <= best; j++) cdf[j]++; }
  }
  for (int i = 0; i <= n; i++) out[i] = (int32_t)cdf[i];
  return 0;
}

extern "C" onedc_rans_tables* onedc_rans_tables_create(const int32_t* cdf, int32_t rows, int32_t row_stride,
                                                       const int32_t* cdf_sizes, const int32_t* offsets) {
  auto* t = new (std::nothrow) onedc_rans_tables();
  if (!t) return nullptr;
  t->rows = rows;
  t->stride = row_stride;
  t->cdf.assign(cdf, cdf + (size_t)rows * row_stride);
  t->sizes.assign(cdf_sizes, cdf_sizes + rows);
  t->offsets.assign(offsets, offsets + rows);
  t->bucket.resize((size_t)rows * 256);
  for (int r = 0; r < rows; r++) {
    const int32_t* row = &t->cdf[(size_t)r * row_stride];
    const int nsym = t->sizes[r] - 1;        // cdf entries 0..nsym (nsym symbols incl. escape)
    int s = 0;
    for (int b = 0; b < 256; b++) {
      const uint32_t c = (uint32_t)b << 8;
      while (s + 1 < nsym && (uint32_t)row[s + 1] <= c) s++;
      t->bucket[(size_t)r * 256 + b] = (uint8_t)s;
    }
  }
  return t;
}
extern "C" void onedc_rans_tables_destroy(onedc_rans_tables* t) { delete t; }

extern "C" onedc_rans_decoder* onedc_rans_decoder_create(void) { return new (std::nothrow) onedc_rans_decoder(); }
extern "C" void onedc_rans_decoder_destroy(onedc_rans_decoder* d) { delete d; }

extern "C" int onedc_rans_decoder_set_stream(onedc_rans_decoder* d, const uint8_t* stream, size_t n) {
  if (!d || !stream || n < 5) return -1;
  d->data.assign(stream, stream + n);
  d->data.resize(n + 8, 0);                 // reads past the end of a truncated stream stay in bounds
  const uint8_t* p = d->data.data();
  const int nstreams = (p[0] >> 4) + 1;
  const int szlen = ((p[0] & 0x0f) == 1) ? 2 : 4;
  size_t off = 1;
  std::vector<size_t> sizes;
  size_t total = 0;
  for (int i = 0; i + 1 < nstreams; i++) {
    if (off + szlen > n) return -1;
    size_t s = 0;
    for (int b = 0; b < szlen; b++) s |= (size_t)p[off + b] << (8 * b);
    off += szlen;
    sizes.push_back(s);
    total += s;
  }
  if (off + total > n) return -1;
  sizes.push_back(n - off - total);
  d->cur.clear();
  for (int i = 0; i < nstreams; i++) {
    if (sizes[i] < 4) return -1;
    Cursor c;
    const uint8_t* q = p + off;
    c.x = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
    c.ptr = q + 4;
    c.end = d->data.data() + d->data.size();
    d->cur.push_back(c);
    off += sizes[i];
  }
  return 0;
}

static inline uint32_t get_bits(Cursor& c, uint32_t nbits) {
  uint32_t x = c.x;
  const uint32_t val = x & ((1u << nbits) - 1);
  x >>= nbits;
  if (x < kL && c.ptr < c.end) x = (x << 8) | *c.ptr++;
  c.x = x;
  return val;
}

static void decode_span(Cursor& c, const onedc_rans_tables* t, const int16_t* idx, int n, int16_t* out) {
  const int32_t* cdf = t->cdf.data();
  const int stride = t->stride;
  uint32_t x = c.x;
  const uint8_t* ptr = c.ptr;
  const uint8_t* end = c.end;
  for (int i = 0; i < n; i++) {
    const int r = idx[i];
    if (r < 0) { out[i] = 0; continue; }
    const int32_t* row = cdf + (size_t)r * stride;
    const int max_value = t->sizes[r] - 2;
    const uint32_t cum = x & kMask;
    int s = t->bucket[(size_t)r * 256 + (cum >> 8)];
    while ((uint32_t)row[s + 1] <= cum) s++;
    const uint32_t start = (uint32_t)row[s], freq = (uint32_t)row[s + 1] - start;
    x = freq * (x >> kPrecision) + cum - start;
    while (x < kL && ptr < end) x = (x << 8) | *ptr++;
    int value = s;
    if (s == max_value) {
      c.x = x; c.ptr = ptr;
      int val = (int)get_bits(c, kBypassBits), nb = val;
      while (val == kBypassMax) { val = (int)get_bits(c, kBypassBits); nb += val; }
      int raw = 0;
      for (int j = 0; j < nb; j++) raw |= (int)get_bits(c, kBypassBits) << (j * kBypassBits);
      value = raw >> 1;
      if (raw & 1) value = -value - 1; else value += max_value;
      x = c.x; ptr = c.ptr;
    }
    out[i] = (int16_t)(value + t->offsets[r]);
  }
  c.x = x;
  c.ptr = ptr;
}

extern "C" int onedc_rans_decoder_decode(onedc_rans_decoder* d, const onedc_rans_tables* t, const int16_t* indexes, int32_t n,
                                         int16_t* out) {
  if (!d || !t || d->cur.empty()) return -1;
  const int nd = (int)d->cur.size();
  const int each = n / nd;
  for (int i = 0; i < nd; i++) {          // stream parts split the symbols evenly (py_rans.cpp:183-221)
    const int cnt = (i < nd - 1) ? each : n - each * (nd - 1);
    decode_span(d->cur[i], t, indexes + (size_t)i * each, cnt, out + (size_t)i * each);
  }
  return 0;
}

extern "C" onedc_rans_encoder* onedc_rans_encoder_create(void) { return new (std::nothrow) onedc_rans_encoder(); }
extern "C" void onedc_rans_encoder_destroy(onedc_rans_encoder* e) { delete e; }
extern "C" void onedc_rans_encoder_reset(onedc_rans_encoder* e) { e->q.clear(); e->stream.clear(); }

extern "C" int onedc_rans_encoder_encode(onedc_rans_encoder* e, const onedc_rans_tables* t, const int16_t* symbols,
                                         const int16_t* indexes, int32_t n) {
  if (!e || !t) return -1;
  e->q.reserve(e->q.size() + (size_t)n + 16);
  for (int i = 0; i < n; i++) {
    const int r = indexes[i];
    if (r < 0) continue;
    const int32_t* row = &t->cdf[(size_t)r * t->stride];
    const int max_value = t->sizes[r] - 2;
    int value = (int)symbols[i] - t->offsets[r];
    uint32_t raw = 0;
    if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; }
    else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; }
    e->q.push_back({(uint16_t)row[value], (uint16_t)(row[value + 1] - row[value])});
    if (value == max_value) {
      int nb = 0;
      while ((raw >> (nb * kBypassBits)) != 0) nb++;
      int v = nb;
      while (v >= kBypassMax) { e->q.push_back({(uint16_t)kBypassMax, 0}); v -= kBypassMax; }
      e->q.push_back({(uint16_t)v, 0});
      for (int j = 0; j < nb; j++) e->q.push_back({(uint16_t)((raw >> (j * kBypassBits)) & kBypassMax), 0});
    }
  }
  return 0;
}

extern "C" int64_t onedc_rans_encoder_flush(onedc_rans_encoder* e) {
  if (!e) return -1;
  // worst case per entry: 16 bits for a coded symbol, 2 bits for a bypass group -> 2 bytes is a safe bound
  std::vector<uint8_t> buf(e->q.size() * 2 + 16);
  uint8_t* endp = buf.data() + buf.size();
  uint8_t* p = endp;
  uint32_t x = kL;
  for (size_t k = e->q.size(); k-- > 0;) {
    const auto s = e->q[k];
    if (s.range != 0) {
      const uint32_t freq = s.range, x_max = freq << 15;
      while (x >= x_max) { *--p = (uint8_t)(x & 0xff); x >>= 8; }
      x = ((x / freq) << kPrecision) + (x % freq) + s.start;
    } else {
      const uint32_t x_max = (1u << (kPrecision - kBypassBits)) << 15;
      while (x >= x_max) { *--p = (uint8_t)(x & 0xff); x >>= 8; }
      x = (x << kBypassBits) | s.start;
    }
  }
  p -= 4;
  p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
  const size_t n = (size_t)(endp - p);
  e->stream.resize(n + 1);
  e->stream[0] = 0x01;                     // one stream part, 2-byte size fields (py_rans.cpp:116-117)
  memcpy(e->stream.data() + 1, p, n);
  return (int64_t)e->stream.size();
}

extern "C" int onedc_rans_encoder_get_stream(const onedc_rans_encoder* e, uint8_t* out, size_t cap) {
  if (!e || cap < e->stream.size()) return -1;
  memcpy(out, e->stream.data(), e->stream.size());
  return 0;
}
