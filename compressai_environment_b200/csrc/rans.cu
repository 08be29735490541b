// rans.cu -- batched, bit-exact Rans64 coder for sm_100a.
//
// What it replaces (reference tree):
//   compressai/cpp_exts/rans/rans_interface.cpp:108-213  BufferedRansEncoder / RansEncoder
//   compressai/cpp_exts/rans/rans_interface.cpp:215-359  RansDecoder (stateless and streaming)
//   third_party/ryg_rans/rans64.h:59-142                 state machine
//
// Design.  One rANS stream is a strict serial chain (state_i depends on state_{i+1}), so the only
// parallelism is across strings: ONE WARP PER STRING, thousands of strings per launch.  Inside a warp
//   * the 32 lanes cooperate on everything that is NOT the chain: coalesced loads of symbols / indexes,
//     CDF lookups in the shared-memory table (staged with one TMA bulk copy per CTA), escape
//     expansion, exact reciprocals for the encoder's division -- one symbol per lane, software
//     pipelined two chunks ahead of the chain;
//   * the chain itself runs warp-uniformly over a 32-entry parameter buffer in shared memory (one
//     LDS.128 broadcast per symbol), so its per-symbol latency is: renorm test, one 64x64 high
//     multiply, one shift, one multiply-add (encoder) / one LUT hit, one multiply-add (decoder);
//   * output words / symbols are collected one per lane and written as full 128-byte lines.
//
// Exact division.  Rans64EncPut needs q = x / freq, r = x % freq with x < 2^63, freq < 2^16
// (rans64.h:92).  For freq >= 2 let l = ceil(log2 freq), m = ceil(2^(63+l) / freq) (fits 64 bits);
// then q = umul64hi(x, m) >> (l-1) for every x < 2^63 (round-up reciprocal: the error term
// e = m*freq - 2^(63+l) < freq <= 2^l, so x*e < 2^(63+l)).  The new state is
//   ((x/f) << 16) + x%f + start = x + start + q * (2^16 - f).
// freq == 1 uses m = 2^64-1, shift 0 (q = x - 1) and folds the missing 2^16 - 1 into the bias.
// The 65536-entry table of m lives in HBM/L2 and is gathered one chunk ahead of the chain.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace cai {

constexpr int kMaxWarpsPerCta = 32;
constexpr uint32_t kRansL_hi = 0u;  // documentation only: L = 2^31

__device__ uint64_t g_rcp[65536];

// host: exact reciprocal table
static void build_rcp_host(uint64_t *tab) {
  tab[0] = 0;
  tab[1] = ~0ull;
  for (uint32_t f = 2; f < 65536; ++f) {
    uint32_t l = 0;
    while ((1u << l) < f) ++l;
    const unsigned __int128 num = (static_cast<unsigned __int128>(1) << (63 + l)) + (f - 1);
    tab[f] = static_cast<uint64_t>(num / f);
  }
}

static int ensure_rcp_table(int device) {
  static std::mutex mu;
  static bool done[64] = {false};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= 64) return CAI_E_NO_DEVICE;
  if (done[device]) return CAI_OK;
  static uint64_t *host = nullptr;
  if (!host) {
    host = new uint64_t[65536];
    build_rcp_host(host);
  }
  CAI_CUDA(cudaMemcpyToSymbol(g_rcp, host, sizeof(uint64_t) * 65536));
  done[device] = true;
  return CAI_OK;
}

// ---- shared-memory carve-up ---------------------------------------------------------------------------
// dynamic smem: [ blob (table) | per-warp staging ]
struct __align__(16) EncParam {
  uint32_t m_lo;
  uint32_t m_hi;        // bit 31 (bit 63 of m, always 1 in the true value) carries the escape flag inverted: see pack
  uint32_t bias_shift;  // bias (17 bits) | shift << 20
  uint32_t thr;         // freq << 15 : renormalise iff (x >> 32) >= thr
};
struct __align__(16) DecParam {
  uint32_t cdf_off;
  int32_t maxv;
  int32_t offset;
  uint32_t lut_off;
};

constexpr int kEncWarpBytes = 2 * 32 * (sizeof(EncParam) + sizeof(uint32_t));  // 1280
constexpr int kDecWarpBytes = 2 * 32 * (sizeof(DecParam) + sizeof(int32_t));     // 1280: parameter ring + index staging

template <bool kSmem>
struct TableView {
  const RowMeta *meta;
  const uint16_t *cdf;
  const uint2 *lut;
  int32_t K;
  int32_t lut_shift;
};

// ---- encoder -------------------------------------------------------------------------------------------
struct EncEmit {
  uint32_t *lane_ptr;  // slot_end - 1 - lane : word k (k = 32 * g + lane) goes to lane_ptr[-32 * g]
  uint32_t cap;        // slot capacity in words (< 2^31)
  uint32_t cnt;        // words emitted so far
  uint32_t buf;
  uint32_t lane;
  __device__ __forceinline__ void flush_full() {  // cnt is a non-zero multiple of 32
    const uint32_t k0 = cnt - 32u;
    if (k0 + lane < cap) *(lane_ptr - static_cast<int64_t>(k0)) = buf;
  }
  __device__ __forceinline__ void push(uint32_t w) {
    if ((cnt & 31u) == lane) buf = w;
    cnt += 1;
    if ((cnt & 31u) == 0) flush_full();
  }
  // predicated variant for the hot path: no branch unless a 128-byte line completes
  __device__ __forceinline__ void push_if(bool on, uint32_t w) {
    buf = (on && (cnt & 31u) == lane) ? w : buf;
    cnt += on ? 1u : 0u;
    if (on && (cnt & 31u) == 0) flush_full();
  }
  __device__ __forceinline__ void finish() {
    const uint32_t rem = cnt & 31u;
    const uint32_t k0 = cnt - rem;
    if (lane < rem && k0 + lane < cap) *(lane_ptr - static_cast<int64_t>(k0)) = buf;
  }
};

template <bool kSmem>
__global__ void __launch_bounds__(1024, 1)
rans_encode_kernel(const unsigned char *__restrict__ blob, uint32_t enc_bytes, const int32_t *__restrict__ symbols,
                   const int32_t *__restrict__ indexes, const int64_t *__restrict__ str_begin,
                   int64_t n_per_string, int32_t B, uint32_t *__restrict__ slots, int64_t slot_words,
                   int32_t *__restrict__ n_words, int32_t *__restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;

  const unsigned char *tbl = blob;
  uint32_t stage_off = 0;
  if (kSmem) {
    stage_blob(smem, blob, enc_bytes, &s_bar);
    tbl = smem;
    stage_off = (enc_bytes + 127u) & ~127u;
  }
  const BlobHeader *hdr = reinterpret_cast<const BlobHeader *>(tbl);
  const int32_t K = hdr->K;
  const uint4 *meta = reinterpret_cast<const uint4 *>(tbl + hdr->off_meta);
  const uint16_t *cdf16 = reinterpret_cast<const uint16_t *>(tbl + hdr->off_cdf);

  EncParam *pbuf = reinterpret_cast<EncParam *>(smem + stage_off + warp * kEncWarpBytes);
  uint32_t *rbuf = reinterpret_cast<uint32_t *>(pbuf + 64);

  for (int64_t b = static_cast<int64_t>(blockIdx.x) * warps + warp; b < B;
       b += static_cast<int64_t>(gridDim.x) * warps) {
    const int64_t beg = str_begin ? str_begin[b] : b * n_per_string;
    const int64_t n = str_begin ? (str_begin[b + 1] - beg) : n_per_string;
    const int32_t *sym = symbols + beg;
    const int32_t *idx = indexes + beg;
    int32_t st = CAI_S_OK;

    EncEmit em;
    em.lane_ptr = slots + (b + 1) * slot_words - 1 - lane;
    em.cap = static_cast<uint32_t>(slot_words);
    em.cnt = 0;
    em.buf = 0;
    em.lane = static_cast<uint32_t>(lane);

    uint64_t x = 1ull << 31;
    const int64_t nchunks = (n + 31) >> 5;

    // pipeline registers
    int32_t r_sym = 0, r_idx = 0;                          // stage 1: raw loads for chunk j-1
    uint32_t p_bias_shift = 0, p_thr = 0, p_raw = 0;       // stage 2: lookups for chunk j
    uint32_t p_freq = 1, p_esc = 0;

    auto stage1 = [&](int64_t j) {
      const int64_t i = (j << 5) + lane;
      if (i < n) {
        r_sym = __ldg(sym + i);
        r_idx = __ldg(idx + i);
      } else {
        r_sym = 0;
        r_idx = 0;
      }
    };
    auto stage2 = [&](int64_t j) {
      const int64_t i = (j << 5) + lane;
      int32_t k = r_idx;
      if (i < n && (k < 0 || k >= K)) {
        st = CAI_S_BAD_INDEX;
        k = 0;
      }
      const uint4 m = meta[k];  // cdf_off, len, offset, lut_off
      const int32_t maxv = static_cast<int32_t>(m.y) - 2;
      int32_t v = static_cast<int32_t>(static_cast<uint32_t>(r_sym) - m.z);
      uint32_t raw = 0, esc = 0;
      if (v < 0) {
        raw = static_cast<uint32_t>(-2) * static_cast<uint32_t>(v) - 1u;
        v = maxv;
        esc = 1;
      } else if (v >= maxv) {
        raw = 2u * static_cast<uint32_t>(v - maxv);
        v = maxv;
        esc = 1;
      }
      if (maxv < 0) v = 0;  // malformed row: stay in bounds
      const uint32_t c0 = cdf16[m.x + v];
      const uint32_t c1 = cdf16[m.x + v + 1];
      uint32_t freq = (c1 - c0) & 0xFFFFu;
      if (freq == 0) freq = 1;  // malformed table: keep the chain well defined
      uint32_t shift = 0, bias = c0;
      if (freq == 1)
        bias += 65535u;
      else
        shift = 31 - __clz(freq - 1);
      p_bias_shift = bias | (shift << 20);
      p_thr = freq << 15;
      p_raw = raw;
      p_esc = esc;
      p_freq = freq;
    };

    if (nchunks > 0) {
      stage1(nchunks - 1);
      stage2(nchunks - 1);
      if (nchunks > 1) stage1(nchunks - 2);
    }
    uint64_t p_m = __ldg(&g_rcp[p_freq]);

    for (int64_t j = nchunks - 1; j >= 0; --j) {
      const int bsel = static_cast<int>(j & 1);
      EncParam *pp = pbuf + bsel * 32;
      uint32_t *rr = rbuf + bsel * 32;
      {
        EncParam e;
        e.m_lo = static_cast<uint32_t>(p_m);
        // true m always has bit 63 set; store the escape flag there (1 = no escape keeps the bit)
        e.m_hi = static_cast<uint32_t>(p_m >> 32) & (p_esc ? 0x7FFFFFFFu : 0xFFFFFFFFu);
        e.bias_shift = p_bias_shift;
        e.thr = p_thr;
        pp[lane] = e;
        rr[lane] = p_raw;
      }
      __syncwarp();
      if (j >= 1) {
        stage2(j - 1);
        p_m = __ldg(&g_rcp[p_freq]);
        if (j >= 2) stage1(j - 2);
      }
      const int64_t rem = n - (j << 5);
      const int nvalid = rem < 32 ? static_cast<int>(rem) : 32;
      uint4 e_next = *reinterpret_cast<const uint4 *>(pp + (nvalid - 1));
#pragma unroll 4
      for (int l = nvalid - 1; l >= 0; --l) {
        const uint4 e = e_next;
        if (l > 0) e_next = *reinterpret_cast<const uint4 *>(pp + (l - 1));  // hide the LDS latency of the next step
        // Rans64EncPut (rans64.h:77-93): renormalise iff x >= freq << 47, then x = x + bias + (x / freq) * (2^16 - freq).
        // Common case: ordinary symbol, no renormalisation -> straight-line code behind ONE branch.
        if (!(e.y & 0x80000000u) || static_cast<uint32_t>(x >> 32) >= e.w) {
          if (!(e.y & 0x80000000u)) {
            // escape: the entry list read backwards is payload nibbles MSB..LSB, then the count nibble, i.e. the
            // (nb+1)-nibble integer V = raw << 4 | nb pushed MSB first.  The reference renormalises before a nibble
            // iff x >= 2^59 (rans_interface.cpp:69-87); with L = bitlen(x) exactly J = (59 - L) / 4 + 1 nibbles fit
            // before that happens, so nibbles are pushed in groups (at most 3 rounds) instead of one by one.
            const uint32_t raw = rr[l];
            const int nb = raw ? ((35 - __clz(raw)) >> 2) : 0;
            const uint64_t V = (static_cast<uint64_t>(raw) << 4) | static_cast<uint32_t>(nb);
            int rem = nb + 1;
            while (rem > 0) {
              const int L = 64 - __clzll(x);
              if (L > 59) {
                em.push(static_cast<uint32_t>(x));
                x >>= 32;
                continue;
              }
              const int J = ((59 - L) >> 2) + 1;
              const int c = J < rem ? J : rem;
              const uint64_t part = (V >> (4 * (rem - c))) & ((1ull << (4 * c)) - 1ull);
              x = (x << (4 * c)) | part;
              rem -= c;
            }
          }
          const uint32_t xh = static_cast<uint32_t>(x >> 32);
          if (xh >= e.w) {
            em.push(static_cast<uint32_t>(x));
            x = static_cast<uint64_t>(xh);
          }
        }
        const uint64_t m = (static_cast<uint64_t>(e.y | 0x80000000u) << 32) | e.x;
        const uint32_t shift = e.z >> 20;
        const uint32_t bias = e.z & 0xFFFFFu;
        const uint32_t cmpl = 65536u - (e.w >> 15);
        const uint64_t q = __umul64hi(x, m) >> shift;
        x = x + bias + q * cmpl;
      }
    }
    em.push(static_cast<uint32_t>(x >> 32));
    em.push(static_cast<uint32_t>(x));
    em.finish();
    if (em.cnt > em.cap) st = CAI_S_OVERFLOW;
    st = __reduce_max_sync(0xffffffffu, st);
    if (lane == 0) {
      n_words[b] = static_cast<int32_t>(em.cnt > em.cap ? em.cap : em.cnt);
      if (status) status[b] = st;
    }
    __syncwarp();
  }
}

// ---- decoder -------------------------------------------------------------------------------------------
struct WordFeed {
  const uint32_t *w;
  uint32_t nw;     // words in this string (< 2^31)
  uint32_t pos;    // next word to consume
  uint32_t cbase;  // multiple of 32: wcur holds [cbase, cbase+32), wnext the following 32
  uint32_t wcur, wnext, next;
  uint32_t lane;
  __device__ __forceinline__ uint32_t load(uint32_t base) const {
    const uint32_t i = base + lane;
    return (i < nw) ? __ldg(w + i) : 0u;
  }
  __device__ __forceinline__ void init(uint32_t p) {
    pos = p;
    cbase = p & ~31u;
    wcur = load(cbase);
    wnext = load(cbase + 32u);
    next = __shfl_sync(0xffffffffu, wcur, static_cast<int>(pos & 31u));
  }
  // consume one word; the following word is fetched from the lane buffers right away so that it is already
  // in a register when the next renormalisation needs it (off the critical path)
  __device__ __forceinline__ uint32_t take() {
    const uint32_t r = next;
    pos += 1;
    if ((pos & 31u) == 0) {
      wcur = wnext;
      cbase += 32u;
      wnext = load(cbase + 32u);
    }
    next = __shfl_sync(0xffffffffu, wcur, static_cast<int>(pos & 31u));
    return r;
  }
};

template <bool kSmem>
__global__ void __launch_bounds__(1024, 1)
rans_decode_kernel(const unsigned char *__restrict__ blob, uint32_t blob_bytes, const uint32_t *__restrict__ words,
                   const int64_t *__restrict__ word_begin, const int32_t *__restrict__ word_count,
                   const int32_t *__restrict__ indexes,
                   const int64_t *__restrict__ str_begin, int64_t n_per_string, int32_t B,
                   int32_t *__restrict__ out, uint64_t *__restrict__ state, int32_t resume,
                   int32_t *__restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;

  const unsigned char *tbl = blob;
  uint32_t stage_off = 0;
  if (kSmem) {
    stage_blob(smem, blob, blob_bytes, &s_bar);
    tbl = smem;
    stage_off = (blob_bytes + 127u) & ~127u;
  }
  const BlobHeader *hdr = reinterpret_cast<const BlobHeader *>(tbl);
  const int32_t K = hdr->K;
  const int lut_shift = hdr->lut_shift;
  const uint4 *meta = reinterpret_cast<const uint4 *>(tbl + hdr->off_meta);
  const uint16_t *cdf16 = reinterpret_cast<const uint16_t *>(tbl + hdr->off_cdf);
  const uint2 *lut = reinterpret_cast<const uint2 *>(tbl + hdr->off_lut);

  DecParam *pbuf = reinterpret_cast<DecParam *>(smem + stage_off + warp * kDecWarpBytes);

  for (int64_t b = static_cast<int64_t>(blockIdx.x) * warps + warp; b < B;
       b += static_cast<int64_t>(gridDim.x) * warps) {
    const int64_t beg = str_begin ? str_begin[b] : b * n_per_string;
    const int64_t n = str_begin ? (str_begin[b + 1] - beg) : n_per_string;
    const int32_t *idx = indexes + beg;
    int32_t *dst = out + beg;
    int32_t st = CAI_S_OK;

    WordFeed wf;
    wf.w = words + word_begin[b];
    wf.nw = static_cast<uint32_t>(word_count ? static_cast<int64_t>(word_count[b]) : (word_begin[b + 1] - word_begin[b]));
    wf.lane = static_cast<uint32_t>(lane);
    uint64_t x;
    if (resume && state) {
      x = state[2 * b];
      wf.init(static_cast<uint32_t>(state[2 * b + 1]));
    } else {
      wf.init(0);
      const uint32_t lo = wf.take();
      const uint32_t hi = wf.take();
      x = static_cast<uint64_t>(lo) | (static_cast<uint64_t>(hi) << 32);
    }

    const int64_t nchunks = (n + 31) >> 5;
    // Index prefetch, one chunk ahead, with cp.async into shared memory.  (As a register prefetch the load's
    // scoreboard wait landed on the first branch of the symbol loop: ~13% of the kernel waiting for DRAM.)
    int32_t *ibuf = reinterpret_cast<int32_t *>(pbuf + 64);
    auto stage1 = [&](int64_t j) {
      const int64_t i = (j << 5) + lane;
      int32_t *dst_s = ibuf + static_cast<int>(j & 1) * 32 + lane;
      if (i < n)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_s)), "l"(idx + i) : "memory");
      else
        *dst_s = 0;
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nchunks > 0) stage1(0);

    for (int64_t j = 0; j < nchunks; ++j) {
      DecParam *pp = pbuf + static_cast<int>(j & 1) * 32;
      {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        int32_t k = ibuf[static_cast<int>(j & 1) * 32 + lane];
        if (k < 0 || k >= K) {
          if ((j << 5) + lane < n) st = CAI_S_BAD_INDEX;
          k = 0;
        }
        const uint4 m = meta[k];
        DecParam d;
        d.cdf_off = m.x;
        d.maxv = static_cast<int32_t>(m.y) - 2;
        d.offset = static_cast<int32_t>(m.z);
        d.lut_off = m.w;
        pp[lane] = d;
      }
      __syncwarp();
      if (j + 1 < nchunks) stage1(j + 1);
      const int64_t rem = n - (j << 5);
      const int nvalid = rem < 32 ? static_cast<int>(rem) : 32;
      int32_t myout = 0;
      uint4 d_next = *reinterpret_cast<const uint4 *>(pp);
#pragma unroll 2
      for (int l = 0; l < nvalid; ++l) {
        const uint4 d = d_next;
        if (l + 1 < nvalid) d_next = *reinterpret_cast<const uint4 *>(pp + (l + 1));  // prefetch next parameters
        const int32_t maxv = static_cast<int32_t>(d.y);
        const uint32_t cf = static_cast<uint32_t>(x) & 0xFFFFu;
        const uint2 e = lut[d.w + (cf >> lut_shift)];
        // Fast path (the common case on real streams): the bucket names one ordinary symbol (the table builder
        // marks buckets of the escape symbol as "search", freq field 0) and the new state needs no refill.  It is
        // straight-line code with ONE branch; every rare event takes the general path below.
        const uint64_t xn = static_cast<uint64_t>(e.x >> 16) * (x >> 16) + (cf - (e.x & 0xFFFFu));
        int32_t v = static_cast<int32_t>(e.y);
        if ((e.x >> 16) != 0u && (xn >> 31) != 0ull) {
          x = xn;
        } else {
          uint32_t start = e.x & 0xFFFFu;
          uint32_t freq = e.x >> 16;
          int32_t s = static_cast<int32_t>(e.y);
          if (freq == 0 && (e.y & 0x80000000u)) {
            // the whole bucket lies inside the escape symbol (table.cu marks it): start is in the entry
            s = maxv;
            start = e.y & 0xFFFFu;
            freq = 0x10000u - start;
          } else if (freq == 0) {
            // bucket spans several symbols: warp-cooperative forward search from s.  Lane i tests symbol s + i
            // (cdf[s+i] <= cf < cdf[s+i+1]); exactly one lane can hit, and a max-reduction broadcasts its
            // (freq, start) pair without a second round of dependent shared-memory loads.
            for (;;) {
              const int32_t cand = s + lane;
              uint32_t packed = 0;
              if (cand <= maxv) {
                const uint32_t c_lo = cdf16[d.x + cand];
                const uint32_t c_hi = (cand == maxv) ? 0x10000u : static_cast<uint32_t>(cdf16[d.x + cand + 1]);
                if (c_lo <= cf && cf < c_hi) packed = ((c_hi - c_lo) << 16) | c_lo;
              }
              const uint32_t win = __reduce_max_sync(0xffffffffu, packed);
              if (win) {
                s += __ffs(__ballot_sync(0xffffffffu, packed != 0)) - 1;
                start = win & 0xFFFFu;
                freq = win >> 16;
                break;
              }
              s += 32;
              if (s > maxv) {  // malformed table: stay in bounds, keep the chain defined
                s = maxv < 0 ? 0 : maxv;
                start = cdf16[d.x + s];
                freq = 1;
                break;
              }
            }
          }
          x = static_cast<uint64_t>(freq) * (x >> 16) + (cf - start);
          if (x < (1ull << 31)) x = (x << 32) | wf.take();
          v = s;
          if (s == maxv) {
            // bypass / escape decoding (rans_interface.cpp:256-278).  Count nibble(s) one at a time, then the
            // payload nibbles in groups: with L = bitlen(x), the reference refills after pop number
            // ceil((L - 31) / 4) (that pop leaves x < 2^31), so that many nibbles can be taken at once.
            // Common case first: the count nibble t0 (< 15) and its t0 payload nibbles can all be popped before the
            // reference would refill (pop number js0 = ceil((L - 31) / 4) is the one that leaves x < 2^31), so they
            // come off in one shift, followed by at most one refill.  Anything else takes the general loop.
            uint32_t raw;
            {
              const uint32_t t0 = static_cast<uint32_t>(x) & 15u;
              const int js0 = (64 - __clzll(x) - 28) >> 2;
              if (t0 < 15u && static_cast<int>(t0) + 1 <= js0) {
                raw = static_cast<uint32_t>((x >> 4) & ((1ull << (4u * t0)) - 1ull));
                x >>= 4u * (t0 + 1u);
                if (static_cast<int>(t0) + 1 == js0) x = (x << 32) | wf.take();
              } else {
                uint32_t t = static_cast<uint32_t>(x) & 15u;
                x >>= 4;
                if (x < (1ull << 31)) x = (x << 32) | wf.take();
                int32_t nb = static_cast<int32_t>(t);
                while (t == 15u) {
                  t = static_cast<uint32_t>(x) & 15u;
                  x >>= 4;
                  if (x < (1ull << 31)) x = (x << 32) | wf.take();
                  nb += static_cast<int32_t>(t);
                }
                uint64_t acc = 0;
                int done = 0;
                int rem = nb;
                while (rem > 0) {
                  const int L = 64 - __clzll(x);
                  int js = (L - 31 + 3) >> 2;
                  if (js < 1) js = 1;  // only reachable on corrupt / truncated streams (x < 2^31)
                  const int c = rem < js ? rem : js;
                  const uint64_t bits = x & ((1ull << (4 * c)) - 1ull);
                  x >>= 4 * c;
                  if (done < 8) acc |= bits << (4 * done);
                  done += c;
                  rem -= c;
                  if (c == js) x = (x << 32) | wf.take();
                }
                raw = static_cast<uint32_t>(acc);
              }
            }
            const int32_t sraw = static_cast<int32_t>(raw);
            v = sraw >> 1;
            v = (sraw & 1) ? (-v - 1) : (v + maxv);
          }
        }
        if (lane == l) myout = v + static_cast<int32_t>(d.z);
      }
      if (lane < nvalid) dst[(j << 5) + lane] = myout;
    }
    if (wf.pos > wf.nw) st = st ? st : CAI_S_TRUNCATED;
    st = __reduce_max_sync(0xffffffffu, st);
    if (lane == 0) {
      if (state) {
        state[2 * b] = x;
        state[2 * b + 1] = static_cast<uint64_t>(wf.pos);
      }
      if (status) status[b] = st;
    }
    __syncwarp();
  }
}

// ---- lane-per-string kernels -----------------------------------------------------------------------------
// From 32 strings up, ONE LANE codes one string: the 32 chains of a warp share one instruction stream, so every
// issued instruction advances 32 strings (the warp-per-string kernels above spend the whole warp's issue slots
// on one chain, and measurably slow each other down from 4 warps per scheduler on).  The chains still cost their
// dependent latency per symbol, but a launch of 32 strings now occupies one warp instead of 32, and 4096 strings
// 128 warps instead of 4096.
//   * symbols / indexes reach the lanes through warp-transposed staging: per 32-symbol chunk the warp copies, for each
//     of its 32 strings, 32 consecutive values (one coalesced 128-byte line, cp.async) into a shared-memory tile
//     with an odd pitch; lane s then walks row s.  Decoded symbols leave through the same transposition.
//   * everything off the chain is software-pipelined one / two symbols ahead inside a chunk: row metadata (LDS),
//     then the per-(row, symbol) encoder parameters (one 16-byte read-only load from a table built once per
//     cai_table: reciprocal, bias | shift, freq << 15 -- the same values the warp kernel derives per symbol).
//   * each lane owns its output cursor (encoder: words stored backwards from its slot end) or its input cursor
//     (decoder: next word prefetched into a register when the previous one is consumed).
constexpr int kTilePitch = 33;
constexpr int kLaneTileInts = 32 * kTilePitch;
constexpr int kEncLaneWarpBytes = 2 * 2 * kLaneTileInts * 4 + 32 * 16;  // sym + idx tiles, double buffered; beg / n
constexpr int kDecLaneWarpBytes = 3 * kLaneTileInts * 4 + 32 * 16;      // idx tiles (x2), output tile; beg / n

__global__ void enc_params_kernel(const unsigned char *__restrict__ blob, uint4 *__restrict__ out) {
  const BlobHeader *hdr = reinterpret_cast<const BlobHeader *>(blob);
  const RowMeta m = reinterpret_cast<const RowMeta *>(blob + hdr->off_meta)[blockIdx.x];
  const uint16_t *row = reinterpret_cast<const uint16_t *>(blob + hdr->off_cdf) + m.cdf_off;
  const int32_t maxv = m.len - 2;
  for (int32_t v = threadIdx.x; v <= maxv; v += blockDim.x) {
    const uint32_t c0 = row[v], c1 = row[v + 1];
    uint32_t freq = (c1 - c0) & 0xFFFFu;
    if (freq == 0) freq = 1;
    uint32_t shift = 0, bias = c0;
    if (freq == 1)
      bias += 65535u;
    else
      shift = 31 - __clz(freq - 1);
    const uint64_t mm = g_rcp[freq];
    // the last symbol of a row is the escape symbol: flag it in bit 63 of m (always set in the true value)
    out[m.cdf_off + v] = make_uint4(static_cast<uint32_t>(mm),
                                    static_cast<uint32_t>(mm >> 32) & (v == maxv ? 0x7FFFFFFFu : 0xFFFFFFFFu),
                                    bias | (shift << 20), freq << 15);
  }
}

static int ensure_enc_params(cai_table *t, cudaStream_t stream) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (t->enc_params) return CAI_OK;
  void *p = nullptr;
  CAI_CUDA(cudaMalloc(&p, sizeof(uint4) * (static_cast<size_t>(t->n_cdf_entries) + 8)));
  enc_params_kernel<<<t->K, 256, 0, stream>>>(t->blob, static_cast<uint4 *>(p));
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // once per table: visible to every stream afterwards
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("enc_params_kernel failed: %s", cudaGetErrorString(e));
    return CAI_E_CUDA;
  }
  t->enc_params = p;
  return CAI_OK;
}

// Copy chunk j (32 values of each of the warp's 32 strings) of `src` into a tile: row s = string wbase + s.
__device__ __forceinline__ void lane_tile_load(int32_t *tile, const int32_t *__restrict__ src, const int64_t *s_beg,
                                               const int64_t *s_n, int64_t j, int lane) {
#pragma unroll 4
  for (int s2 = 0; s2 < 32; ++s2) {
    const int64_t i = (j << 5) + lane;
    int32_t *dst = tile + s2 * kTilePitch + lane;
    if (i < s_n[s2])
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src + s_beg[s2] + i) : "memory");
  }
}

__global__ void __launch_bounds__(256, 1)
rans_encode_lanes_kernel(const unsigned char *__restrict__ blob, const uint4 *__restrict__ eparams,
                         const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes,
                         const int64_t *__restrict__ str_begin, int64_t n_per_string, int32_t B,
                         uint32_t *__restrict__ slots, int64_t slot_words, int32_t *__restrict__ n_words,
                         int32_t *__restrict__ status, int32_t meta_in_smem) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const BlobHeader *hdr = reinterpret_cast<const BlobHeader *>(blob);
  const int32_t K = hdr->K;
  const uint4 *meta = reinterpret_cast<const uint4 *>(blob + hdr->off_meta);
  uint32_t woff = 0;
  if (meta_in_smem) {
    uint4 *sm = reinterpret_cast<uint4 *>(smem);
    for (int k = threadIdx.x; k < K; k += blockDim.x) sm[k] = __ldg(meta + k);
    __syncthreads();
    meta = sm;
    woff = (static_cast<uint32_t>(K) * 16u + 127u) & ~127u;
  }
  unsigned char *wsm = smem + woff + warp * kEncLaneWarpBytes;
  int32_t *tile_sym = reinterpret_cast<int32_t *>(wsm);                    // [2][kLaneTileInts]
  int32_t *tile_idx = tile_sym + 2 * kLaneTileInts;                       // [2][kLaneTileInts]
  int64_t *s_beg = reinterpret_cast<int64_t *>(tile_idx + 2 * kLaneTileInts);  // [32]
  int64_t *s_n = s_beg + 32;                                              // [32]

  for (int64_t wbase = (static_cast<int64_t>(blockIdx.x) * warps + warp) * 32; wbase < B;
       wbase += static_cast<int64_t>(gridDim.x) * warps * 32) {
    const int64_t b = wbase + lane;
    const bool live = b < B;
    const int64_t beg = live ? (str_begin ? str_begin[b] : b * n_per_string) : 0;
    const int64_t n = live ? (str_begin ? (str_begin[b + 1] - beg) : n_per_string) : 0;
    __syncwarp();
    s_beg[lane] = beg;
    s_n[lane] = n;
    __syncwarp();
    int64_t nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int64_t other = __shfl_xor_sync(0xffffffffu, nmax, o);
      nmax = other > nmax ? other : nmax;
    }
    const int64_t nchunks = (nmax + 31) >> 5;

    uint64_t x = 1ull << 31;
    uint32_t cnt = 0;
    const uint32_t cap = static_cast<uint32_t>(slot_words);
    uint32_t *outp = slots + (b + 1) * slot_words - 1;  // word k goes to outp[-k]
    int32_t st = CAI_S_OK;
    auto push = [&](uint32_t w) {
      if (cnt < cap) *(outp - cnt) = w;
      cnt += 1;
    };

    if (nchunks > 0) {
      lane_tile_load(tile_sym + ((nchunks - 1) & 1) * kLaneTileInts, symbols, s_beg, s_n, nchunks - 1, lane);
      lane_tile_load(tile_idx + ((nchunks - 1) & 1) * kLaneTileInts, indexes, s_beg, s_n, nchunks - 1, lane);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int64_t j = nchunks - 1; j >= 0; --j) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (j >= 1) {  // the other buffer was chunk j + 1's: every lane is done with it
        lane_tile_load(tile_sym + ((j - 1) & 1) * kLaneTileInts, symbols, s_beg, s_n, j - 1, lane);
        lane_tile_load(tile_idx + ((j - 1) & 1) * kLaneTileInts, indexes, s_beg, s_n, j - 1, lane);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int32_t *tsym = tile_sym + (j & 1) * kLaneTileInts + lane * kTilePitch;
      const int32_t *tidx = tile_idx + (j & 1) * kLaneTileInts + lane * kTilePitch;
      const int64_t i0 = j << 5;

      // stage M: row metadata of symbol t;  stage A: clamp / escape split + parameter load;  then the chain
      bool m_ok = false, a_ok = false;
      int32_t m_sym = 0;
      uint4 m_meta = make_uint4(0u, 2u, 0u, 0u);
      uint4 a_e = make_uint4(0u, 0u, 0u, 0u);
      uint32_t a_raw = 0;
      auto stage_m = [&](int t) {
        m_ok = t >= 0 && (i0 + t) < n;
        int32_t k = 0;
        m_sym = 0;
        if (m_ok) {
          k = tidx[t];
          m_sym = tsym[t];
          if (k < 0 || k >= K) {
            st = CAI_S_BAD_INDEX;
            k = 0;
          }
        }
        m_meta = meta[k];
      };
      auto stage_a = [&]() {
        a_ok = m_ok;
        const int32_t maxv = static_cast<int32_t>(m_meta.y) - 2;
        int32_t v = static_cast<int32_t>(static_cast<uint32_t>(m_sym) - m_meta.z);
        uint32_t raw = 0;
        if (v < 0) {
          raw = static_cast<uint32_t>(-2) * static_cast<uint32_t>(v) - 1u;
          v = maxv;
        } else if (v >= maxv) {
          raw = 2u * static_cast<uint32_t>(v - maxv);
          v = maxv;
        }
        if (maxv < 0) v = 0;  // malformed row: stay in bounds
        a_raw = raw;
        a_e = __ldg(eparams + (a_ok ? (m_meta.x + static_cast<uint32_t>(v)) : 0u));
      };
      stage_m(31);
      stage_a();
      stage_m(30);
#pragma unroll 2
      for (int t = 31; t >= 0; --t) {
        const uint4 e = a_e;
        const uint32_t raw = a_raw;
        const bool ok = a_ok;
        stage_a();
        stage_m(t - 2);
        if (ok) {
          // Rans64EncPut (rans64.h:77-93) with the exact reciprocal; escapes as in the warp kernel above
          if (!(e.y & 0x80000000u)) {
            const int nb = raw ? ((35 - __clz(raw)) >> 2) : 0;
            const uint64_t V = (static_cast<uint64_t>(raw) << 4) | static_cast<uint32_t>(nb);
            int rem = nb + 1;
            while (rem > 0) {
              const int L = 64 - __clzll(x);
              if (L > 59) {
                push(static_cast<uint32_t>(x));
                x >>= 32;
                continue;
              }
              const int J = ((59 - L) >> 2) + 1;
              const int c = J < rem ? J : rem;
              const uint64_t part = (V >> (4 * (rem - c))) & ((1ull << (4 * c)) - 1ull);
              x = (x << (4 * c)) | part;
              rem -= c;
            }
          }
          const uint32_t xh = static_cast<uint32_t>(x >> 32);
          if (xh >= e.w) {
            push(static_cast<uint32_t>(x));
            x = static_cast<uint64_t>(xh);
          }
          const uint64_t m = (static_cast<uint64_t>(e.y | 0x80000000u) << 32) | e.x;
          const uint32_t shift = e.z >> 20;
          const uint32_t bias = e.z & 0xFFFFFu;
          const uint32_t cmpl = 65536u - (e.w >> 15);
          const uint64_t q = __umul64hi(x, m) >> shift;
          x = x + bias + q * cmpl;
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
      push(static_cast<uint32_t>(x >> 32));
      push(static_cast<uint32_t>(x));
      if (cnt > cap) st = CAI_S_OVERFLOW;
      n_words[b] = static_cast<int32_t>(cnt > cap ? cap : cnt);
      if (status) status[b] = st;
    }
  }
}

template <bool kSmem>
__global__ void __launch_bounds__(256, 1)
rans_decode_lanes_kernel(const unsigned char *__restrict__ blob, uint32_t blob_bytes,
                         const uint32_t *__restrict__ words, const int64_t *__restrict__ word_begin,
                         const int32_t *__restrict__ word_count, const int32_t *__restrict__ indexes,
                         const int64_t *__restrict__ str_begin, int64_t n_per_string, int32_t B,
                         int32_t *__restrict__ out, int32_t *__restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t s_bar;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  const unsigned char *tbl = blob;
  uint32_t stage_off = 0;
  if (kSmem) {
    stage_blob(smem, blob, blob_bytes, &s_bar);
    tbl = smem;
    stage_off = (blob_bytes + 127u) & ~127u;
  }
  const BlobHeader *hdr = reinterpret_cast<const BlobHeader *>(tbl);
  const int32_t K = hdr->K;
  const int lut_shift = hdr->lut_shift;
  const int32_t nb_row = hdr->lut_buckets;
  const uint4 *meta = reinterpret_cast<const uint4 *>(tbl + hdr->off_meta);
  const uint16_t *cdf16 = reinterpret_cast<const uint16_t *>(tbl + hdr->off_cdf);
  const uint2 *lut = reinterpret_cast<const uint2 *>(tbl + hdr->off_lut);

  unsigned char *wsm = smem + stage_off + warp * kDecLaneWarpBytes;
  int32_t *tile_idx = reinterpret_cast<int32_t *>(wsm);                      // [2][kLaneTileInts]
  int32_t *tile_out = tile_idx + 2 * kLaneTileInts;                         // [kLaneTileInts]
  int64_t *s_beg = reinterpret_cast<int64_t *>(tile_out + kLaneTileInts);   // [32]
  int64_t *s_n = s_beg + 32;                                                // [32]

  for (int64_t wbase = (static_cast<int64_t>(blockIdx.x) * warps + warp) * 32; wbase < B;
       wbase += static_cast<int64_t>(gridDim.x) * warps * 32) {
    const int64_t b = wbase + lane;
    const bool live = b < B;
    const int64_t beg = live ? (str_begin ? str_begin[b] : b * n_per_string) : 0;
    const int64_t n = live ? (str_begin ? (str_begin[b + 1] - beg) : n_per_string) : 0;
    __syncwarp();
    s_beg[lane] = beg;
    s_n[lane] = n;
    __syncwarp();
    int64_t nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int64_t other = __shfl_xor_sync(0xffffffffu, nmax, o);
      nmax = other > nmax ? other : nmax;
    }
    const int64_t nchunks = (nmax + 31) >> 5;

    // word feed of this lane's string: the next word sits in a register before the chain needs it
    const uint32_t *wp = words + (live ? word_begin[b] : 0);
    const uint32_t nw = live ? static_cast<uint32_t>(word_count ? static_cast<int64_t>(word_count[b])
                                                                  : (word_begin[b + 1] - word_begin[b]))
                             : 0u;
    uint32_t pos = 0;
    uint32_t nxt = nw > 0 ? __ldg(wp) : 0u;
    auto take = [&]() -> uint32_t {
      const uint32_t r = nxt;
      pos += 1;
      nxt = pos < nw ? __ldg(wp + pos) : 0u;
      return r;
    };
    uint64_t x;
    {
      const uint32_t lo = take();
      const uint32_t hi = take();
      x = static_cast<uint64_t>(lo) | (static_cast<uint64_t>(hi) << 32);
    }
    int32_t st = CAI_S_OK;

    if (nchunks > 0) lane_tile_load(tile_idx, indexes, s_beg, s_n, 0, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int64_t j = 0; j < nchunks; ++j) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();  // tile j visible; every lane has finished chunk j - 1 (other index buffer, output tile)
      if (j + 1 < nchunks) lane_tile_load(tile_idx + ((j + 1) & 1) * kLaneTileInts, indexes, s_beg, s_n, j + 1, lane);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const int32_t *tidx = tile_idx + (j & 1) * kLaneTileInts + lane * kTilePitch;
      int32_t *tout = tile_out + lane * kTilePitch;
      const int64_t i0 = j << 5;

      bool d_ok = false;
      uint4 d_meta = make_uint4(0u, 2u, 0u, 0u);
      auto stage_d = [&](int t) {
        d_ok = t < 32 && (i0 + t) < n;
        int32_t k = 0;
        if (d_ok) {
          k = tidx[t];
          if (k < 0 || k >= K) {
            st = CAI_S_BAD_INDEX;
            k = 0;
          }
        }
        d_meta = meta[k];
      };
      stage_d(0);
#pragma unroll 2
      for (int t = 0; t < 32; ++t) {
        const uint4 d = d_meta;  // cdf_off, len, offset, lut_off
        const bool ok = d_ok;
        stage_d(t + 1);
        if (ok) {
          const int32_t maxv = static_cast<int32_t>(d.y) - 2;
          const uint32_t cf = static_cast<uint32_t>(x) & 0xFFFFu;
          const uint32_t bk = cf >> lut_shift;
          const uint2 e = lut[d.w + bk];
          const uint64_t xn = static_cast<uint64_t>(e.x >> 16) * (x >> 16) + (cf - (e.x & 0xFFFFu));
          int32_t v = static_cast<int32_t>(e.y);
          if ((e.x >> 16) != 0u && (xn >> 31) != 0ull) {
            x = xn;  // one ordinary symbol, no refill
          } else {
            uint32_t start = e.x & 0xFFFFu;
            uint32_t freq = e.x >> 16;
            int32_t s = static_cast<int32_t>(e.y);
            if (freq == 0 && (e.y & 0x80000000u)) {
              s = maxv;  // the whole bucket lies inside the escape symbol
              start = e.y & 0xFFFFu;
              freq = 0x10000u - start;
            } else if (freq == 0) {
              // bucket spans several symbols: lane-local binary search for the last s in [s, hi] with cdf[s] <= cf;
              // hi = the symbol holding the next bucket's first value (or the row's last symbol)
              int32_t hi = maxv;
              if (static_cast<int32_t>(bk) + 1 < nb_row) {
                const uint32_t ny = lut[d.w + bk + 1].y;
                hi = (ny & 0x80000000u) ? maxv : static_cast<int32_t>(ny);
              }
              if (hi > maxv) hi = maxv;
              if (s < 0) s = 0;
              while (s < hi) {
                const int32_t mid = (s + hi + 1) >> 1;
                if (static_cast<uint32_t>(cdf16[d.x + mid]) <= cf)
                  s = mid;
                else
                  hi = mid - 1;
              }
              start = cdf16[d.x + s];
              const uint32_t c_hi = (s >= maxv) ? 0x10000u : static_cast<uint32_t>(cdf16[d.x + s + 1]);
              freq = c_hi - start;
              if (freq == 0 || freq > 0x10000u) freq = 1;  // malformed table: keep the chain defined
            }
            x = static_cast<uint64_t>(freq) * (x >> 16) + (cf - start);
            if (x < (1ull << 31)) x = (x << 32) | take();
            v = s;
            if (s == maxv) {
              // bypass / escape decoding (rans_interface.cpp:256-278), grouped nibble pops as in the warp kernel
              uint32_t raw;
              const uint32_t t0 = static_cast<uint32_t>(x) & 15u;
              const int js0 = (64 - __clzll(x) - 28) >> 2;
              if (t0 < 15u && static_cast<int>(t0) + 1 <= js0) {
                raw = static_cast<uint32_t>((x >> 4) & ((1ull << (4u * t0)) - 1ull));
                x >>= 4u * (t0 + 1u);
                if (static_cast<int>(t0) + 1 == js0) x = (x << 32) | take();
              } else {
                uint32_t tt = static_cast<uint32_t>(x) & 15u;
                x >>= 4;
                if (x < (1ull << 31)) x = (x << 32) | take();
                int32_t nb = static_cast<int32_t>(tt);
                while (tt == 15u) {
                  tt = static_cast<uint32_t>(x) & 15u;
                  x >>= 4;
                  if (x < (1ull << 31)) x = (x << 32) | take();
                  nb += static_cast<int32_t>(tt);
                }
                uint64_t acc = 0;
                int done = 0;
                int rem = nb;
                while (rem > 0) {
                  const int L = 64 - __clzll(x);
                  int js = (L - 31 + 3) >> 2;
                  if (js < 1) js = 1;  // only reachable on corrupt / truncated streams (x < 2^31)
                  const int c = rem < js ? rem : js;
                  const uint64_t bits = x & ((1ull << (4 * c)) - 1ull);
                  x >>= 4 * c;
                  if (done < 8) acc |= bits << (4 * done);
                  done += c;
                  rem -= c;
                  if (c == js) x = (x << 32) | take();
                }
                raw = static_cast<uint32_t>(acc);
              }
              const int32_t sraw = static_cast<int32_t>(raw);
              v = sraw >> 1;
              v = (sraw & 1) ? (-v - 1) : (v + maxv);
            }
          }
          tout[t] = v + static_cast<int32_t>(d.z);
        }
      }
      __syncwarp();
      // transposed store: row s2 of the output tile = 32 consecutive symbols of string wbase + s2
#pragma unroll 4
      for (int s2 = 0; s2 < 32; ++s2) {
        const int64_t i = i0 + lane;
        if (i < s_n[s2]) out[s_beg[s2] + i] = tile_out[s2 * kTilePitch + lane];
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live) {
      if (pos > nw) st = st ? st : CAI_S_TRUNCATED;
      if (status) status[b] = st;
    }
  }
}

// ---- compaction ----------------------------------------------------------------------------------------
__global__ void scan_words_kernel(const int32_t *__restrict__ n_words, int32_t B, int64_t *__restrict__ out_begin) {
  __shared__ int64_t s_scan[1024];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int32_t base = 0; base < B; base += blockDim.x) {
    const int32_t k = base + threadIdx.x;
    const int64_t n = (k < B) ? static_cast<int64_t>(n_words[k]) : 0;
    s_scan[threadIdx.x] = n;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
      const int64_t v = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    if (k < B) out_begin[k] = s_carry + s_scan[threadIdx.x] - n;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry += s_scan[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) out_begin[B] = s_carry;
}

__global__ void compact_kernel(const uint32_t *__restrict__ slots, int64_t slot_words,
                               const int32_t *__restrict__ n_words, int32_t B,
                               const int64_t *__restrict__ out_begin, uint32_t *__restrict__ out,
                               int64_t out_cap) {
  for (int32_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int64_t n = n_words[b];
    const uint32_t *src = slots + (static_cast<int64_t>(b) + 1) * slot_words - n;
    const int64_t o = out_begin[b];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
      if (o + i < out_cap) out[o + i] = src[i];
  }
}

}  // namespace cai

using namespace cai;

extern "C" {

int64_t cai_rans_slot_words(int64_t n_symbols) {
  if (n_symbols < 0) n_symbols = 0;
  // <= 16 + 4*9 = 52 bits per symbol, 31 bits of initial state, 2 flush words; rounded to 128-byte lines
  const int64_t w = (52 * n_symbols + 31 + 31) / 32 + 2;
  return (w + 31) & ~static_cast<int64_t>(31);
}

static int plan_grid(const DeviceProps &dp, int32_t B, int *warps, int *grid) {
  // A string is a serial chain: its warp issues ~0.2 instructions per cycle, so several strings share an SM
  // sub-partition without slowing each other much.  Few, fat CTAs (>= 8 warps) keep the coder on a handful of
  // SMs -- each CTA pins a copy of the table in shared memory -- and leave the rest of the chip to the
  // transform kernels running concurrently on other streams.  Measured on the headline step (32 strings per launch):
  // 16 warps per CTA (2 SMs per launch) lengthen the chains by 15-20 % (decode 73 -> 85 ms per launch) but halve the
  // coder's SM-time, and the step gains 1.3 % (104.5 -> 103.1 ms); 32 warps per CTA lose (138 ms).
  int w = (B + dp.sm_count / 4 - 1) / (dp.sm_count / 4 > 0 ? dp.sm_count / 4 : 1);
  const int kw = knobs().coder_warps;
  if (kw >= 1 && kw <= kMaxWarpsPerCta) w = w > kw ? w : kw;
  else if (w < 16) w = 16;
  if (w > kMaxWarpsPerCta) {
    // thousands of strings (C3): the kernels are issue-bound (ncu r02: SM throughput 73-77 %), so every SM must carry
    // its share -- 4096 strings as 128 CTAs x 32 warps left 20 of the 148 SMs idle; 147 CTAs x 28 warps use them all
    w = (B + dp.sm_count - 1) / dp.sm_count;
    if (w > kMaxWarpsPerCta) w = kMaxWarpsPerCta;
    if (w < 16) w = 16;
  }
  if (w > B) w = B < 1 ? 1 : B;
  int g = (B + w - 1) / w;
  if (g > dp.sm_count) g = dp.sm_count;  // persistent: each warp strides over strings
  if (g < 1) g = 1;
  *warps = w;
  *grid = g;
  return 0;
}

// Lane-per-string kernels from CAI_CODER_LANES strings per launch up; off by default.  Measured on B200 (r2):
// bit-exact, and a launch of 32 strings occupies one warp instead of 32 -- but every instruction of the per-symbol
// step (~115 warp instructions with the divergent refill / search / escape paths serialised) now sits on ONE
// in-order issue stream at ~5 cycles apiece, so a step costs 600-800 cycles against 160-250 cycles per symbol of
// the warp-per-string kernels whose off-chain work is spread over 32 lanes: C3 9.2 / 5.5 vs 18.3 / 15.4 Gsym/s.
static bool use_lanes(int32_t B) {
  const int k = knobs().coder_lanes;
  return k > 0 && B >= k;
}

// 32 strings per warp; as many CTAs as there are SMs before a CTA gets a second warp (chains are latency bound:
// spreading the warps buys more than packing them), at most `max_warps` warps per CTA, persistent beyond that.
static void plan_lanes(const DeviceProps &dp, int32_t B, int max_warps, int *warps, int *grid) {
  const int nw = (B + 31) / 32;
  int w = (nw + dp.sm_count - 1) / dp.sm_count;
  if (w > max_warps) w = max_warps;
  if (w < 1) w = 1;
  int g = (nw + w - 1) / w;
  const int cap = dp.sm_count * 4;
  if (g > cap) g = cap;
  *warps = w;
  *grid = g < 1 ? 1 : g;
}

int cai_rans_encode_batch(cai_table_t t, const int32_t *symbols, const int32_t *indexes,
                          const int64_t *str_begin, int64_t n_per_string, int32_t B, uint32_t *slots,
                          int64_t slot_words, int32_t *n_words, int32_t *status, cai_stream_t stream_) {
  CAI_CHECK_ARG(t != nullptr, "cai_rans_encode_batch: NULL table");
  CAI_CHECK_ARG(B >= 0, "cai_rans_encode_batch: B < 0");
  if (B == 0) return CAI_OK;
  CAI_CHECK_ARG(slots && n_words, "cai_rans_encode_batch: NULL output");
  CAI_CHECK_ARG(str_begin || n_per_string >= 0, "cai_rans_encode_batch: negative string length");
  CAI_CHECK_ARG((symbols && indexes) || (!str_begin && n_per_string == 0),
                "cai_rans_encode_batch: NULL symbols / indexes");
  CAI_CHECK_ARG(slot_words >= 2 && (slot_words % 32) == 0,
                "cai_rans_encode_batch: slot_words must be a multiple of 32 (use cai_rans_slot_words)");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  rc = ensure_rcp_table(dp.device);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (use_lanes(B)) {
    // lane-per-string: 32 strings per warp (see the kernel comment); tables are read through L1 / L2 off the chain
    rc = ensure_enc_params(t, stream);
    if (rc != CAI_OK) return rc;
    int lw, lgrid;
    plan_lanes(dp, B, 8, &lw, &lgrid);
    const int meta_in_smem = static_cast<size_t>(t->K) * 16 <= 32 * 1024;
    const size_t smem = (meta_in_smem ? ((static_cast<size_t>(t->K) * 16 + 127) & ~static_cast<size_t>(127)) : 0) +
                        static_cast<size_t>(lw) * kEncLaneWarpBytes;
    int max_dyn = 0;
    rc = optin_max_smem(reinterpret_cast<const void *>(rans_encode_lanes_kernel), dp, &max_dyn);
    if (rc != CAI_OK) return rc;
    CAI_CHECK_ARG(smem <= static_cast<size_t>(max_dyn), "cai_rans_encode_batch: staging does not fit shared memory");
    rans_encode_lanes_kernel<<<lgrid, lw * 32, smem, stream>>>(t->blob, static_cast<const uint4 *>(t->enc_params),
                                                              symbols, indexes, str_begin, n_per_string, B, slots,
                                                              slot_words, n_words, status, meta_in_smem);
    CAI_LAUNCH_CHECK();
    return CAI_OK;
  }
  int warps, grid;
  plan_grid(dp, B, &warps, &grid);
  if (t->enc_in_smem) {
    const size_t smem = ((t->enc_bytes + 127u) & ~127u) + static_cast<size_t>(warps) * kEncWarpBytes;
    int max_dyn = 0;
    rc = optin_max_smem(reinterpret_cast<const void *>(rans_encode_kernel<true>), dp, &max_dyn);  // once per device
    if (rc != CAI_OK) return rc;
    CAI_CHECK_ARG(smem <= static_cast<size_t>(max_dyn), "cai_rans_encode_batch: table does not fit shared memory");
    rans_encode_kernel<true><<<grid, warps * 32, smem, stream>>>(t->blob, t->enc_bytes, symbols, indexes,
                                                                 str_begin, n_per_string, B, slots,
                                                                 slot_words, n_words, status);
  } else {
    const size_t smem = static_cast<size_t>(warps) * kEncWarpBytes;
    rans_encode_kernel<false><<<grid, warps * 32, smem, stream>>>(t->blob, t->enc_bytes, symbols, indexes,
                                                                  str_begin, n_per_string, B, slots,
                                                                  slot_words, n_words, status);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_rans_compact(const uint32_t *slots, int64_t slot_words, const int32_t *n_words, int32_t B,
                     int64_t *out_begin, uint32_t *out_words, int64_t out_capacity_words,
                     cai_stream_t stream_) {
  CAI_CHECK_ARG(B >= 0 && out_begin && n_words, "cai_rans_compact: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  scan_words_kernel<<<1, 1024, 0, stream>>>(n_words, B, out_begin);
  CAI_LAUNCH_CHECK();
  if (out_words && B > 0) {
    CAI_CHECK_ARG(slots != nullptr, "cai_rans_compact: NULL slots");
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc != CAI_OK) return rc;
    int grid = B < dp.sm_count * 8 ? B : dp.sm_count * 8;
    compact_kernel<<<grid, 256, 0, stream>>>(slots, slot_words, n_words, B, out_begin, out_words,
                                             out_capacity_words);
    CAI_LAUNCH_CHECK();
  }
  return CAI_OK;
}

int cai_rans_decode_batch(cai_table_t t, const uint32_t *words, const int64_t *word_begin,
                          const int32_t *word_count, const int32_t *indexes, const int64_t *str_begin, int64_t n_per_string,
                          int32_t B, int32_t *out, uint64_t *state, int32_t resume, int32_t *status,
                          cai_stream_t stream_) {
  CAI_CHECK_ARG(t != nullptr, "cai_rans_decode_batch: NULL table");
  CAI_CHECK_ARG(B >= 0, "cai_rans_decode_batch: B < 0");
  if (B == 0) return CAI_OK;
  CAI_CHECK_ARG(words && word_begin, "cai_rans_decode_batch: NULL stream");
  CAI_CHECK_ARG(!resume || state, "cai_rans_decode_batch: resume needs a state array");
  CAI_CHECK_ARG(str_begin || n_per_string >= 0, "cai_rans_decode_batch: negative string length");
  CAI_CHECK_ARG((indexes && out) || (!str_begin && n_per_string == 0),
                "cai_rans_decode_batch: NULL indexes / out");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const uint32_t bytes = static_cast<uint32_t>(t->blob_bytes);
  if (use_lanes(B) && !state) {
    int max_dyn = 0;
    const void *fn = t->in_smem ? reinterpret_cast<const void *>(rans_decode_lanes_kernel<true>)
                                : reinterpret_cast<const void *>(rans_decode_lanes_kernel<false>);
    rc = optin_max_smem(fn, dp, &max_dyn);
    if (rc != CAI_OK) return rc;
    const size_t base = t->in_smem ? ((bytes + 127u) & ~127u) : 0;
    int fit = static_cast<int>((static_cast<size_t>(max_dyn) - base) / kDecLaneWarpBytes);
    CAI_CHECK_ARG(static_cast<size_t>(max_dyn) > base && fit >= 1, "cai_rans_decode_batch: table does not fit shared memory");
    int lw, lgrid;
    plan_lanes(dp, B, fit < 8 ? fit : 8, &lw, &lgrid);
    if (t->in_smem && lgrid > dp.sm_count) lgrid = dp.sm_count;  // one CTA per SM: each stages the table
    const size_t smem = base + static_cast<size_t>(lw) * kDecLaneWarpBytes;
    if (t->in_smem)
      rans_decode_lanes_kernel<true><<<lgrid, lw * 32, smem, stream>>>(t->blob, bytes, words, word_begin, word_count,
                                                                      indexes, str_begin, n_per_string, B, out, status);
    else
      rans_decode_lanes_kernel<false><<<lgrid, lw * 32, smem, stream>>>(t->blob, bytes, words, word_begin, word_count,
                                                                       indexes, str_begin, n_per_string, B, out, status);
    CAI_LAUNCH_CHECK();
    return CAI_OK;
  }
  int warps, grid;
  plan_grid(dp, B, &warps, &grid);
  if (t->in_smem) {
    const size_t smem = ((bytes + 127u) & ~127u) + static_cast<size_t>(warps) * kDecWarpBytes;
    int max_dyn = 0;
    rc = optin_max_smem(reinterpret_cast<const void *>(rans_decode_kernel<true>), dp, &max_dyn);  // once per device
    if (rc != CAI_OK) return rc;
    CAI_CHECK_ARG(smem <= static_cast<size_t>(max_dyn), "cai_rans_decode_batch: table does not fit shared memory");
    rans_decode_kernel<true><<<grid, warps * 32, smem, stream>>>(t->blob, bytes, words, word_begin, word_count, indexes,
                                                                 str_begin, n_per_string, B, out, state,
                                                                 resume, status);
  } else {
    const size_t smem = static_cast<size_t>(warps) * kDecWarpBytes;
    rans_decode_kernel<false><<<grid, warps * 32, smem, stream>>>(t->blob, bytes, words, word_begin, word_count, indexes,
                                                                  str_begin, n_per_string, B, out, state,
                                                                  resume, status);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

}  // extern "C"
