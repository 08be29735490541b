// pmf.cu -- pmf_to_quantized_cdf for K rows in one launch (one CTA per row).
//
// What it replaces: compressai/cpp_exts/ops/ops.cpp:40-109 called once per row through pybind
// (entropy_models.py:89-92, :204-212; 64 + C calls per update(), 37 ms for the default Gaussian table).
//
// The reference's "steal" loop (ops.cpp:74-100) is O(m^2) and sequential: for every zero-width symbol i
// (ascending) it finds the FIRST symbol with the smallest frequency > 1 and moves one unit of
// frequency from it to i by shifting the CDF entries in between.  Restated in the frequency domain
// the shift is exactly "freq[best] -= 1, freq[i] += 1", and because the donor that was just
// decremented is then strictly the smallest donor, the loop drains donors one after the other in
// (freq, index) order until as many units as there are zero-width symbols have been collected.
// That is a parallel algorithm: sort donors by (freq, index), exclusive-scan their spare units
// (freq - 1), clamp against the number of zero-width symbols, and prefix-sum the repaired
// frequencies back into a CDF.  Bit-exact with the sequential loop (tests: test_pmf_* ).
#include "common.cuh"

namespace cai {

constexpr int kPmfThreads = 256;
constexpr int kPmfMaxM = 8192;

__device__ __forceinline__ uint32_t block_reduce_add(uint32_t v, uint32_t *s_red) {
  __syncthreads();
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (int w = 0; w < kPmfThreads / 32; ++w) t += s_red[w];
  return t;
}

// exclusive scan (mod 2^32) of a[0..n) in shared memory; returns the total.  All threads call.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t *a, int n, uint32_t *s_part) {
  const int per = (n + kPmfThreads - 1) / kPmfThreads;
  const int lo = threadIdx.x * per, hi = min(lo + per, n);
  uint32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  __syncthreads();
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int t = 0; t < kPmfThreads; ++t) {
      const uint32_t v = s_part[t];
      s_part[t] = run;
      run += v;
    }
    s_part[kPmfThreads] = run;
  }
  __syncthreads();
  uint32_t run = s_part[threadIdx.x];
  for (int i = lo; i < hi; ++i) {
    const uint32_t v = a[i];
    a[i] = run;
    run += v;
  }
  __syncthreads();
  return s_part[kPmfThreads];
}

__global__ void __launch_bounds__(kPmfThreads)
pmf_to_cdf_kernel(const float *__restrict__ pmf, const int32_t *__restrict__ pmf_len, const float *__restrict__ tail,
                  int32_t Lp, int32_t precision, int32_t P /* pow2 >= max m */, int32_t *__restrict__ cdf,
                  int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);    // P
  uint32_t *freq = reinterpret_cast<uint32_t *>(smem_raw + sizeof(unsigned long long) * P);  // P
  uint32_t *spare = freq + P;                                                      // P (scan workspace)
  __shared__ uint32_t s_red[kPmfThreads / 32];
  __shared__ uint32_t s_part[kPmfThreads + 1];
  __shared__ int s_bad;

  const int k = blockIdx.x;
  const int W = Lp + (tail ? 2 : 1);  // output row width
  int32_t *row = cdf + static_cast<int64_t>(k) * W;
  int len = pmf_len[k];
  len = len < 0 ? 0 : (len > Lp ? Lp : len);
  const int m = len + (tail ? 1 : 0);
  const float *src = pmf + static_cast<int64_t>(k) * Lp;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();

  // 1. validate + round half away from zero (std::round on float, ops.cpp:57-58)
  const float scale = static_cast<float>(1 << precision);
  uint32_t part = 0;
  for (int i = threadIdx.x; i < P; i += kPmfThreads) {
    uint32_t c = 0;
    if (i < m) {
      const float p = (i < len) ? src[i] : tail[k];
      if (p < 0.f || !isfinite(p)) s_bad = 1;
      c = __float2uint_rz(roundf(__fmul_rn(p, scale)));
    }
    freq[i] = c;
    part += c;
  }
  const uint32_t total = block_reduce_add(part, s_red);
  int st = CAI_S_OK;
  if (s_bad)
    st = CAI_S_BAD_PMF;
  else if (total == 0 || m == 0)
    st = CAI_S_ZERO_PMF;
  if (st != CAI_S_OK) {
    for (int i = threadIdx.x; i < W; i += kPmfThreads) row[i] = 0;
    if (threadIdx.x == 0) status[k] = st;
    return;
  }

  // 2. rescale so the total is 2^precision (floor), force the last symbol to absorb the slack
  part = 0;
  const unsigned long long one = 1ull << precision;
  for (int i = threadIdx.x; i < m; i += kPmfThreads) {
    const uint32_t f = static_cast<uint32_t>((one * freq[i]) / total);
    freq[i] = f;
    if (i < m - 1) part += f;
  }
  const uint32_t head = block_reduce_add(part, s_red);
  if (threadIdx.x == 0) freq[m - 1] = static_cast<uint32_t>(one) - head;
  __syncthreads();

  // 3. zero-width symbols and donors
  part = 0;
  for (int i = threadIdx.x; i < P; i += kPmfThreads) {
    unsigned long long key = ~0ull;
    if (i < m) {
      const uint32_t f = freq[i];
      if (f == 0) part += 1;
      if (f > 1) key = (static_cast<unsigned long long>(f) << 32) | static_cast<uint32_t>(i);
    }
    keys[i] = key;
  }
  const uint32_t n_zero = block_reduce_add(part, s_red);

  if (n_zero > 0) {
    // 4. bitonic sort of the donors by (freq, index)
    for (int size = 2; size <= P; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        __syncthreads();
        for (int t = threadIdx.x; t < (P >> 1); t += kPmfThreads) {
          const int i = 2 * t - (t & (stride - 1));
          const int j = i + stride;
          const bool up = (i & size) == 0;
          const unsigned long long a = keys[i], b = keys[j];
          if ((a > b) == up) {
            keys[i] = b;
            keys[j] = a;
          }
        }
      }
    }
    __syncthreads();
    // 5. exclusive scan of the spare units in sorted order, clamp against n_zero
    for (int i = threadIdx.x; i < P; i += kPmfThreads) {
      const unsigned long long key = keys[i];
      spare[i] = (key == ~0ull) ? 0u : (static_cast<uint32_t>(key >> 32) - 1u);
    }
    const uint32_t avail = block_excl_scan(spare, P, s_part);
    if (avail < n_zero) st = CAI_S_NO_DONOR;
    for (int i = threadIdx.x; i < P; i += kPmfThreads) {
      const unsigned long long key = keys[i];
      if (key == ~0ull) continue;
      const uint32_t f = static_cast<uint32_t>(key >> 32);
      const uint32_t before = spare[i];
      if (before >= n_zero) continue;
      const uint32_t want = n_zero - before;
      const uint32_t give = want < (f - 1u) ? want : (f - 1u);
      freq[static_cast<uint32_t>(key)] = f - give;
    }
    __syncthreads();
    // 6. every zero-width symbol receives one unit (only as many as could be stolen, in index order)
    if (st == CAI_S_OK) {
      for (int i = threadIdx.x; i < m; i += kPmfThreads)
        if (freq[i] == 0) freq[i] = 1;
    } else {
      // not enough donors: the reference would leave the remaining zero-width symbols untouched;
      // hand out the available units to the first zero-width symbols in order
      for (int i = threadIdx.x; i < P; i += kPmfThreads) spare[i] = (i < m && freq[i] == 0) ? 1u : 0u;
      block_excl_scan(spare, P, s_part);
      for (int i = threadIdx.x; i < m; i += kPmfThreads)
        if (freq[i] == 0 && spare[i] < avail) freq[i] = 1;
    }
    __syncthreads();
  }

  // 7. back to a CDF
  for (int i = m + threadIdx.x; i < P; i += kPmfThreads) freq[i] = 0;
  __syncthreads();
  block_excl_scan(freq, P, s_part);
  for (int i = threadIdx.x; i < W; i += kPmfThreads) {
    int32_t v = 0;
    if (i < m)
      v = static_cast<int32_t>(freq[i]);
    else if (i == m)
      v = static_cast<int32_t>(one);
    row[i] = v;
  }
  if (threadIdx.x == 0) status[k] = st;
}

}  // namespace cai

using namespace cai;

extern "C" int cai_pmf_to_quantized_cdf(const float *pmf, const int32_t *pmf_len, const float *tail, int32_t K,
                                        int32_t Lp, int32_t precision, int32_t *cdf, int32_t *status,
                                        cai_stream_t stream_) {
  CAI_CHECK_ARG(K >= 0, "cai_pmf_to_quantized_cdf: K < 0");
  if (K == 0) return CAI_OK;
  CAI_CHECK_ARG(pmf && pmf_len && cdf && status, "cai_pmf_to_quantized_cdf: NULL pointer");
  CAI_CHECK_ARG(precision >= 1 && precision <= 16, "cai_pmf_to_quantized_cdf: precision %d not in [1, 16]", precision);
  CAI_CHECK_ARG(Lp >= 1 && Lp + 1 <= kPmfMaxM, "cai_pmf_to_quantized_cdf: row length %d not in [1, %d]", Lp,
                kPmfMaxM - 1);
  int P = 2;
  while (P < Lp + 1) P <<= 1;
  const size_t smem = static_cast<size_t>(P) * (sizeof(unsigned long long) + 2 * sizeof(uint32_t));
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  int max_dyn = 0;
  rc = optin_max_smem(reinterpret_cast<const void *>(pmf_to_cdf_kernel), dp, &max_dyn);  // once per device
  if (rc != CAI_OK) return rc;
  CAI_CHECK_ARG(smem <= static_cast<size_t>(max_dyn), "cai_pmf_to_quantized_cdf: row length %d needs too much shared memory", Lp);
  pmf_to_cdf_kernel<<<K, kPmfThreads, smem, static_cast<cudaStream_t>(stream_)>>>(pmf, pmf_len, tail, Lp, precision,
                                                                                  P, cdf, status);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}
