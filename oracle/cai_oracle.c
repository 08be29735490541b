/*
 * cai_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the integer arithmetic of the reference's codec hot path, used ONLY as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under compressai_environment_b200/ may import, link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (a) the golden vectors in tests/golden/ (generated from the compiled reference, oracle/_ref,
 *       by tests/golden/make_golden.py), including the reference's own KAT
 *       tests/test_ops.py:104-106, and
 *   (b) when oracle/_ref is importable, against the reference itself on random streams.
 *
 * Reference locations restated here (paths relative to /root/reference):
 *   third_party/ryg_rans/rans64.h:59-142                      state machine (L = 2^31, 32-bit renorm)
 *   compressai/cpp_exts/rans/rans_interface.cpp:49-52         precision 16, bypass 4 bits
 *   compressai/cpp_exts/rans/rans_interface.cpp:69-105        raw-bit put/get
 *   compressai/cpp_exts/rans/rans_interface.cpp:108-200       symbol mapping + reverse flush
 *   compressai/cpp_exts/rans/rans_interface.cpp:215-359       decode, stateful decode
 *   compressai/cpp_exts/ops/ops.cpp:40-109                    pmf -> quantized cdf
 *
 * Deliberate differences from the reference (documented defects):
 *   - streams of 0 or 1 symbols are well defined here (the reference under-runs its buffer);
 *   - the encoder never buffers an entry list: it walks the symbols backwards and re-derives each
 *     symbol's escape expansion on the fly, which produces the same word sequence.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PRECISION 16
#define ORC_BYPASS_BITS 4
#define ORC_BYPASS_MAX 15u
#define ORC_L (1ull << 31)

/* ---- word sink that fills a caller buffer from its END towards its start ---------------------- */
typedef struct {
  uint32_t *base;
  int64_t cap;  /* words */
  int64_t used; /* words written so far (at base[cap-used .. cap)) */
  int overflow;
} orc_sink;

static inline void sink_push(orc_sink *s, uint32_t w) {
  if (s->used >= s->cap) {
    s->overflow = 1;
    return;
  }
  s->used += 1;
  s->base[s->cap - s->used] = w;
}

/* rans64.h:77-93 with scale_bits = 16 */
static inline uint64_t enc_put(uint64_t x, orc_sink *s, uint32_t start, uint32_t freq) {
  const uint64_t x_max = ((ORC_L >> ORC_PRECISION) << 32) * (uint64_t)freq;
  if (x >= x_max) {
    sink_push(s, (uint32_t)x);
    x >>= 32;
  }
  return ((x / freq) << ORC_PRECISION) + (x % freq) + start;
}

/* rans_interface.cpp:69-87 with nbits = 4 */
static inline uint64_t enc_put_bits(uint64_t x, orc_sink *s, uint32_t val) {
  const uint64_t x_max = ((ORC_L >> 16) << 32) * (uint64_t)(1u << (16 - ORC_BYPASS_BITS));
  if (x >= x_max) {
    sink_push(s, (uint32_t)x);
    x >>= 32;
  }
  return (x << ORC_BYPASS_BITS) | val;
}

/*
 * Encode one string.  Returns the number of 32-bit words produced (>= 2) or -1 on overflow / bad
 * input.  The words are left at out[cap_words - ret .. cap_words) in DECODE order, i.e. the byte
 * string the reference returns is exactly those words, little endian.
 * (rans_interface.cpp:108-200)
 */
int64_t orc_rans_encode(const int32_t *symbols, const int32_t *indexes, int64_t n,
                        const int32_t *cdfs, const int32_t *cdf_len, const int32_t *offsets,
                        int32_t K, int32_t Lmax, uint32_t *out, int64_t cap_words) {
  orc_sink s = {out, cap_words, 0, 0};
  uint64_t x = ORC_L;
  for (int64_t i = n - 1; i >= 0; --i) {
    const int32_t k = indexes[i];
    if (k < 0 || k >= K) return -1;
    const int32_t *row = cdfs + (int64_t)k * Lmax;
    const int32_t max_value = cdf_len[k] - 2;
    if (max_value < 0 || max_value + 1 >= Lmax) return -1;
    int32_t v = symbols[i] - offsets[k];
    uint32_t raw = 0;
    if (v < 0) {
      raw = (uint32_t)(-2 * v - 1);
      v = max_value;
    } else if (v >= max_value) {
      raw = (uint32_t)(2 * (v - max_value));
      v = max_value;
    }
    if (v == max_value) {
      /* forward order would be: SYM, count nibble(s), payload nibbles LSB first.
       * We are walking backwards, so: payload MSB..LSB, count nibbles reversed, then SYM. */
      int32_t nb = 0;
      while (nb < 8 && (raw >> (nb * ORC_BYPASS_BITS)) != 0) ++nb;
      for (int32_t j = nb - 1; j >= 0; --j)
        x = enc_put_bits(x, &s, (raw >> (j * ORC_BYPASS_BITS)) & ORC_BYPASS_MAX);
      /* count: unary chunks of 15 then the remainder (the chunk loop is dead for nb <= 8) */
      int32_t rem = nb, n15 = 0;
      while (rem >= (int32_t)ORC_BYPASS_MAX) {
        rem -= ORC_BYPASS_MAX;
        ++n15;
      }
      x = enc_put_bits(x, &s, (uint32_t)rem);
      for (int32_t t = 0; t < n15; ++t) x = enc_put_bits(x, &s, ORC_BYPASS_MAX);
    }
    const uint32_t start = (uint32_t)row[v] & 0xFFFFu;
    const uint32_t freq = ((uint32_t)row[v + 1] - (uint32_t)row[v]) & 0xFFFFu;
    if (freq == 0) return -1;
    x = enc_put(x, &s, start, freq);
  }
  /* rans64.h:96-103: word[0] = low half, word[1] = high half */
  sink_push(&s, (uint32_t)(x >> 32));
  sink_push(&s, (uint32_t)x);
  return s.overflow ? -1 : s.used;
}

/* ---- decoder ---------------------------------------------------------------------------------- */
typedef struct {
  uint64_t x;
  int64_t pos; /* next word to read */
} orc_dec_state;

static inline uint32_t dec_get_bits(uint64_t *px, const uint32_t *w, int64_t nw, int64_t *pos) {
  uint64_t x = *px;
  const uint32_t val = (uint32_t)(x & ORC_BYPASS_MAX);
  x >>= ORC_BYPASS_BITS;
  if (x < ORC_L) {
    const uint32_t nx = (*pos < nw) ? w[*pos] : 0u; /* the reference reads past the end; we feed zeros */
    *pos += 1;
    x = (x << 32) | nx;
  }
  *px = x;
  return val;
}

void orc_rans_dec_init(orc_dec_state *st, const uint32_t *w, int64_t nw) {
  const uint64_t lo = nw > 0 ? w[0] : 0u, hi = nw > 1 ? w[1] : 0u;
  st->x = lo | (hi << 32);
  st->pos = 2;
}

/* rans_interface.cpp:215-284 / :294-359 (identical loops; state carried in *st) */
int orc_rans_decode_stream(orc_dec_state *st, const uint32_t *w, int64_t nw, const int32_t *indexes,
                           int64_t n, const int32_t *cdfs, const int32_t *cdf_len,
                           const int32_t *offsets, int32_t K, int32_t Lmax, int32_t *out) {
  uint64_t x = st->x;
  int64_t pos = st->pos;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t k = indexes[i];
    if (k < 0 || k >= K) return -1;
    const int32_t *row = cdfs + (int64_t)k * Lmax;
    const int32_t len = cdf_len[k];
    const int32_t max_value = len - 2;
    const uint32_t cf = (uint32_t)(x & 0xFFFFu);
    /* first j with row[j] > cf (linear in the reference; any exact search gives the same j) */
    int32_t lo = 0, hi = len; /* answer in [0, len] */
    while (lo < hi) {
      const int32_t mid = (lo + hi) >> 1;
      if ((uint32_t)row[mid] > cf)
        hi = mid;
      else
        lo = mid + 1;
    }
    const int32_t sidx = lo - 1;
    if (sidx < 0 || sidx + 1 >= len) return -2;
    const uint32_t start = (uint32_t)row[sidx];
    const uint32_t freq = (uint32_t)row[sidx + 1] - start;
    x = (uint64_t)freq * (x >> ORC_PRECISION) + cf - start;
    if (x < ORC_L) {
      const uint32_t nx = (pos < nw) ? w[pos] : 0u;
      pos += 1;
      x = (x << 32) | nx;
    }
    int32_t value = sidx;
    if (value == max_value) {
      int32_t val = (int32_t)dec_get_bits(&x, w, nw, &pos);
      int32_t nb = val;
      while (val == (int32_t)ORC_BYPASS_MAX) {
        val = (int32_t)dec_get_bits(&x, w, nw, &pos);
        nb += val;
      }
      uint32_t raw = 0;
      for (int32_t j = 0; j < nb; ++j) {
        val = (int32_t)dec_get_bits(&x, w, nw, &pos);
        if (j < 8) raw |= (uint32_t)val << (j * ORC_BYPASS_BITS);
      }
      /* the reference keeps raw in an int32 and uses an arithmetic shift (:266-277) */
      const int32_t sraw = (int32_t)raw;
      value = sraw >> 1;
      if (sraw & 1)
        value = -value - 1;
      else
        value += max_value;
    }
    out[i] = value + offsets[k];
  }
  st->x = x;
  st->pos = pos;
  return 0;
}

int orc_rans_decode(const uint32_t *w, int64_t nw, const int32_t *indexes, int64_t n,
                    const int32_t *cdfs, const int32_t *cdf_len, const int32_t *offsets, int32_t K,
                    int32_t Lmax, int32_t *out) {
  orc_dec_state st;
  orc_rans_dec_init(&st, w, nw);
  return orc_rans_decode_stream(&st, w, nw, indexes, n, cdfs, cdf_len, offsets, K, Lmax, out);
}

/* ---- pmf -> quantized cdf  (ops.cpp:40-109) ---------------------------- */
/* returns 0 ok, -1 negative / non-finite entry, -2 all-zero pmf.  cdf has m + 1 entries. */
int orc_pmf_to_quantized_cdf(const float *pmf, int32_t m, int32_t precision, uint32_t *cdf) {
  for (int32_t i = 0; i < m; ++i)
    if (pmf[i] < 0 || !isfinite(pmf[i])) return -1;
  cdf[0] = 0;
  const float scale = (float)(1 << precision);
  for (int32_t i = 0; i < m; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * scale); /* half away from 0 */
  uint32_t total = 0;
  for (int32_t i = 0; i <= m; ++i) total += cdf[i];
  if (total == 0) return -2;
  for (int32_t i = 0; i <= m; ++i) cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
  for (int32_t i = 1; i <= m; ++i) cdf[i] += cdf[i - 1];
  cdf[m] = 1u << precision;
  for (int32_t i = 0; i < m; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best_freq = ~0u;
    int32_t best = -1;
    for (int32_t j = 0; j < m; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best_freq) {
        best_freq = f;
        best = j;
      }
    }
    if (best < 0) return -3; /* nothing to steal from (the reference asserts) */
    if (best < i)
      for (int32_t j = best + 1; j <= i; ++j) cdf[j] -= 1;
    else
      for (int32_t j = i + 1; j <= best; ++j) cdf[j] += 1;
  }
  return 0;
}

/* Batched front end with the reference caller's row convention (entropy_models.py:204-212):
 * row k = cat(pmf[k, :pmf_len[k]], tail[k]) -> cdf row of pmf_len[k] + 2 entries, zero padded. */
int orc_pmf_rows_to_cdf(const float *pmf, const int32_t *pmf_len, const float *tail, int32_t K,
                        int32_t Lp, int32_t precision, int32_t *cdf_out /* K x (Lp+2) */) {
  float *tmp = (float *)malloc(sizeof(float) * (size_t)(Lp + 1));
  uint32_t *c = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(Lp + 2));
  int rc = 0;
  for (int32_t k = 0; k < K && rc == 0; ++k) {
    const int32_t m = pmf_len[k];
    memcpy(tmp, pmf + (int64_t)k * Lp, sizeof(float) * (size_t)m);
    tmp[m] = tail[k];
    rc = orc_pmf_to_quantized_cdf(tmp, m + 1, precision, c);
    int32_t *row = cdf_out + (int64_t)k * (Lp + 2);
    memset(row, 0, sizeof(int32_t) * (size_t)(Lp + 2));
    if (rc == 0)
      for (int32_t i = 0; i < m + 2; ++i) row[i] = (int32_t)c[i];
  }
  free(tmp);
  free(c);
  return rc;
}

/* ---- multi-string convenience used by the CPU baseline: one call codes strings [0, B); callers
 * get multi-core runs by giving disjoint string ranges to several threads (ctypes drops the GIL). */
int64_t orc_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int64_t n_per, int64_t B,
                              const int32_t *cdfs, const int32_t *cdf_len, const int32_t *offsets,
                              int32_t K, int32_t Lmax, uint32_t *out, int64_t cap_words_per,
                              int64_t *n_words) {
  int64_t bad = 0;
  for (int64_t b = 0; b < B; ++b) {
    n_words[b] = orc_rans_encode(symbols + b * n_per, indexes + b * n_per, n_per, cdfs, cdf_len,
                                 offsets, K, Lmax, out + b * cap_words_per, cap_words_per);
    if (n_words[b] < 0) bad += 1;
  }
  return bad;
}

int64_t orc_rans_decode_batch(const uint32_t *slots, int64_t cap_words_per, const int64_t *n_words,
                              const int32_t *indexes, int64_t n_per, int64_t B, const int32_t *cdfs,
                              const int32_t *cdf_len, const int32_t *offsets, int32_t K, int32_t Lmax,
                              int32_t *out) {
  int64_t bad = 0;
  for (int64_t b = 0; b < B; ++b) {
    const uint32_t *w = slots + b * cap_words_per + (cap_words_per - n_words[b]);
    if (orc_rans_decode(w, n_words[b], indexes + b * n_per, n_per, cdfs, cdf_len, offsets, K, Lmax,
                        out + b * n_per) != 0)
      bad += 1;
  }
  return bad;
}
