// table.cu -- packed CDF tables resident in HBM + host-side plumbing of the C ABI.
//
// Replaces the per-call Python-list -> std::vector<std::vector<int>> conversion of the reference
// (compressai/cpp_exts/rans/rans_interface.cpp:108-113, :215-221).  The int32 [K, Lmax] table of
// EntropyModel._quantized_cdf is packed once into a blob laid out for the coder kernels:
//
//   [ BlobHeader 64 B | RowMeta K x 16 B | cdf uint16 ragged rows | decode LUT K x nb x 8 B ]
//
// * cdf values are stored as uint16; the terminal 65536 wraps to 0, which is harmless because the
//   kernels never compare against the last entry of a row (its position encodes "infinity") and
//   frequencies are taken modulo 2^16 exactly like the reference's static_cast<uint16_t>
//   (rans_interface.cpp:142-144).
// * LUT entry for (row, bucket b = cf >> shift): {start | freq << 16, s0}.  s0 is the symbol that
//   contains the first cumulative frequency of the bucket; freq != 0 means the whole bucket lies
//   inside symbol s0 (decode needs no search), freq == 0 means "search forward from s0".
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "common.cuh"

namespace cai {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int get_device_props(DeviceProps *out) {
  static thread_local DeviceProps cache[64];
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return CAI_E_NO_DEVICE;
  }
  DeviceProps &p = cache[dev];
  if (p.device != dev) {
    CAI_CUDA(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
    CAI_CUDA(cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    p.device = dev;
  }
  *out = p;
  return CAI_OK;
}

int optin_max_smem(const void *kernel, const DeviceProps &dp, int *max_dynamic, int carveout) {
  static std::mutex mu;
  static std::set<std::pair<int, const void *>> done;
  static std::set<std::pair<int, const void *>> carved;
  std::lock_guard<std::mutex> lk(mu);
  cudaFuncAttributes fa;
  const std::pair<int, const void *> key(dp.device, kernel);
  if (!done.count(key) || max_dynamic) {
    CAI_CUDA(cudaFuncGetAttributes(&fa, kernel));
    const int lim = dp.max_smem_optin - static_cast<int>(fa.sharedSizeBytes);
    if (!done.count(key)) {
      CAI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
      done.insert(key);
    }
    if (max_dynamic) *max_dynamic = lim;
  }
  if (carveout >= 0 && !carved.count(key)) {
    CAI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
    carved.insert(key);
  }
  return CAI_OK;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

const Knobs &knobs() {
  static const Knobs k = [] {
    Knobs r;
    r.conv_stages = env_int("CAI_CONV_STAGES", 0);
    r.conv_generic = getenv("CAI_CONV_GENERIC") ? 1 : 0;
    r.conv_carveout = env_int("CAI_CONV_CARVEOUT", -1);
    r.conv_debug = env_int("CAI_CONV_DEBUG", 0);
    r.patch_generic = getenv("CAI_PATCH_GENERIC") ? 1 : 0;
    r.coder_warps = env_int("CAI_CODER_WARPS", 0);
    r.lut_buckets = env_int("CAI_LUT_BUCKETS", 0);
    r.table_smem_kb = env_int("CAI_TABLE_SMEM_KB", -1);
    r.coder_lanes = env_int("CAI_CODER_LANES", 0);
    r.conv_persist = env_int("CAI_CONV_PERSIST", -1);
    r.coder_lut_adapt = env_int("CAI_LUT_ADAPT", -1);
    r.tma_epi_warps = env_int("CAI_TMA_EPI_WARPS", 0);
    return r;
  }();
  return k;
}

// ---- kernels ---------------------------------------------------------------------------------------

// One block: validate lengths and lay the rows out back to back (each row padded to 8 entries = 16 B).
__global__ void table_layout_kernel(const int32_t *__restrict__ cdf_len, int32_t K, int32_t Lmax,
                                    uint32_t *__restrict__ row_off, uint32_t *__restrict__ total) {
  __shared__ uint32_t s_carry;
  __shared__ uint32_t s_scan[1024];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int32_t base = 0; base < K; base += blockDim.x) {
    const int32_t k = base + threadIdx.x;
    uint32_t n = 0;
    if (k < K) {
      int32_t len = cdf_len[k];
      len = len < 0 ? 0 : (len > Lmax ? Lmax : len);
      n = (static_cast<uint32_t>(len) + 7u) & ~7u;
    }
    s_scan[threadIdx.x] = n;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {  // Hillis-Steele inclusive scan
      uint32_t v = threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    if (k < K) row_off[k] = s_carry + s_scan[threadIdx.x] - n;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry += s_scan[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void table_pack_kernel(const int32_t *__restrict__ cdfs, const int32_t *__restrict__ cdf_len,
                                  const int32_t *__restrict__ offsets, const uint32_t *__restrict__ row_off,
                                  BlobHeader hdr, unsigned char *__restrict__ blob) {
  const int32_t k = blockIdx.x;
  RowMeta *meta = reinterpret_cast<RowMeta *>(blob + hdr.off_meta);
  uint16_t *cdf16 = reinterpret_cast<uint16_t *>(blob + hdr.off_cdf);
  int32_t len = cdf_len[k];
  len = len < 0 ? 0 : (len > hdr.Lmax ? hdr.Lmax : len);
  const uint32_t off = row_off[k];
  if (k == 0 && threadIdx.x == 0) *reinterpret_cast<BlobHeader *>(blob) = hdr;
  if (threadIdx.x == 0) {
    RowMeta m;
    m.cdf_off = off;
    m.len = len;
    m.offset = offsets[k];
    m.lut_off = static_cast<uint32_t>(k) * static_cast<uint32_t>(hdr.lut_buckets);
    meta[k] = m;
  }
  const uint32_t padded = (static_cast<uint32_t>(len) + 7u) & ~7u;
  const int32_t *row = cdfs + static_cast<int64_t>(k) * hdr.Lmax;
  for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x)
    cdf16[off + i] = (i < static_cast<uint32_t>(len)) ? static_cast<uint16_t>(row[i]) : uint16_t(0);
}

__global__ void table_lut_kernel(BlobHeader hdr, unsigned char *__restrict__ blob) {
  const int32_t k = blockIdx.x;
  const RowMeta m = reinterpret_cast<const RowMeta *>(blob + hdr.off_meta)[k];
  const uint16_t *row = reinterpret_cast<const uint16_t *>(blob + hdr.off_cdf) + m.cdf_off;
  uint2 *lut = reinterpret_cast<uint2 *>(blob + hdr.off_lut) + m.lut_off;
  const int32_t nsym = m.len - 1;  // symbols 0 .. len-2 ; cdf[len-1] is the terminal 2^16
  for (int32_t b = threadIdx.x; b < hdr.lut_buckets; b += blockDim.x) {
    uint2 e = make_uint2(0u, 0u);
    if (nsym >= 1) {
      const uint32_t cf0 = static_cast<uint32_t>(b) << hdr.lut_shift;
      const uint32_t cf1 = cf0 + (1u << hdr.lut_shift) - 1u;
      // largest s in [0, nsym-1] with row[s] <= cf0
      int32_t lo = 0, hi = nsym - 1;
      while (lo < hi) {
        const int32_t mid = (lo + hi + 1) >> 1;
        if (static_cast<uint32_t>(row[mid]) <= cf0)
          lo = mid;
        else
          hi = mid - 1;
      }
      const uint32_t start = row[lo];
      const uint32_t next = (lo + 1 >= nsym) ? 65536u : static_cast<uint32_t>(row[lo + 1]);
      e.y = static_cast<uint32_t>(lo);
      // a bucket is "direct" (freq field != 0) only if it names ONE ORDINARY symbol; buckets that straddle a symbol
      // boundary, and buckets of the escape symbol (the last one), are left as "search from e.y" so that the
      // decoder's fast path needs a single test
      if (next > cf1 && next > start) {
        if (lo != nsym - 1)
          e.x = start | ((next - start) << 16);
        else
          e.y = 0x80000000u | start;  // whole bucket inside the escape symbol: no search needed, freq = 2^16 - start
      }
    }
    lut[b] = e;
  }
}

}  // namespace cai

using namespace cai;

extern "C" {

int cai_abi_version(void) { return CAI_ABI_VERSION; }

const char *cai_last_error(void) { return g_err; }

int cai_device_info(int *sm_count, int *max_smem_per_block) {
  DeviceProps p;
  int rc = get_device_props(&p);
  if (rc != CAI_OK) return rc;
  if (sm_count) *sm_count = p.sm_count;
  if (max_smem_per_block) *max_smem_per_block = p.max_smem_optin;
  return CAI_OK;
}

int cai_table_create(const int32_t *cdfs, const int32_t *cdf_len, const int32_t *offsets, int32_t K,
                     int32_t Lmax, cai_stream_t stream_, cai_table_t *out) {
  CAI_CHECK_ARG(out != nullptr, "cai_table_create: out is NULL");
  *out = nullptr;
  CAI_CHECK_ARG(cdfs && cdf_len && offsets, "cai_table_create: NULL table pointer");
  CAI_CHECK_ARG(K >= 1 && K <= (1 << 20), "cai_table_create: K=%d out of range", K);
  CAI_CHECK_ARG(Lmax >= 2 && Lmax <= 65537, "cai_table_create: Lmax=%d out of range [2, 65537]", Lmax);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;

  uint32_t *row_off = nullptr;
  CAI_CUDA(cudaMalloc(&row_off, sizeof(uint32_t) * (static_cast<size_t>(K) + 1)));
  table_layout_kernel<<<1, 1024, 0, stream>>>(cdf_len, K, Lmax, row_off, row_off + K);
  uint32_t total_entries = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(&total_entries, row_off + K, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    cudaFree(row_off);
    set_error("cai_table_create: layout pass failed: %s", cudaGetErrorString(e));
    return CAI_E_CUDA;
  }

  BlobHeader h{};
  h.magic = kBlobMagic;
  h.K = K;
  h.Lmax = Lmax;
  h.n_cdf_entries = total_entries;
  h.off_meta = sizeof(BlobHeader);
  h.off_cdf = h.off_meta + static_cast<uint32_t>(K) * sizeof(RowMeta);
  const uint64_t cdf_bytes = static_cast<uint64_t>(total_entries) * 2u;  // multiple of 16
  const uint64_t base = static_cast<uint64_t>(h.off_cdf) + cdf_bytes;
  if (base + static_cast<uint64_t>(K) * 8u > (1ull << 31)) {
    cudaFree(row_off);
    set_error("cai_table_create: table too large (%llu bytes)", static_cast<unsigned long long>(base));
    return CAI_E_TOO_LARGE;
  }
  h.off_lut = static_cast<uint32_t>(base);
  h.enc_bytes = h.off_lut;
  // Shared memory budget: the coder kernels add 32 warps x 1280 B of staging, round the blob up to 128 B and own a
  // few bytes of static shared memory (mbarrier); 1 KB covers the latter two.
  const int64_t budget = static_cast<int64_t>(dp.max_smem_optin) - 32 * 1280 - 1024;
  int nb = 256;
  { const int v = knobs().lut_buckets; if (v >= 1 && v <= 256 && (v & (v - 1)) == 0) nb = v; }
  while (nb > 1 && static_cast<int64_t>(base) + static_cast<int64_t>(K) * nb * 8 > budget) nb >>= 1;
  const int in_smem = static_cast<int64_t>(base) + static_cast<int64_t>(K) * nb * 8 <= budget;
  if (!in_smem) nb = 256;  // tables live in L2; keep the LUT fine
  h.lut_buckets = nb;
  int sh = 16;
  for (int t = nb; t > 1; t >>= 1) --sh;
  h.lut_shift = sh;
  // padded to 16 bytes: the blob is staged with cp.async.bulk, whose size must be a multiple of 16
  h.total_bytes = (h.off_lut + static_cast<uint32_t>(K) * static_cast<uint32_t>(nb) * 8u + 15u) & ~15u;

  cai_table *t = new cai_table();
  e = cudaMalloc(&t->blob, h.total_bytes);
  if (e != cudaSuccess) {
    cudaFree(row_off);
    delete t;
    set_error("cai_table_create: cudaMalloc(%u) failed: %s", h.total_bytes, cudaGetErrorString(e));
    return CAI_E_CUDA;
  }
  table_pack_kernel<<<K, 256, 0, stream>>>(cdfs, cdf_len, offsets, row_off, h, t->blob);
  table_lut_kernel<<<K, 256, 0, stream>>>(h, t->blob);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(row_off);
  if (e != cudaSuccess) {
    cudaFree(t->blob);
    delete t;
    set_error("cai_table_create: pack failed: %s", cudaGetErrorString(e));
    return CAI_E_CUDA;
  }
  t->blob_bytes = h.total_bytes;
  t->n_cdf_entries = h.n_cdf_entries;
  t->enc_bytes = h.enc_bytes;
  t->K = K;
  t->Lmax = Lmax;
  t->lut_shift = h.lut_shift;
  t->lut_buckets = h.lut_buckets;
  // CAI_TABLE_SMEM_KB: tables larger than this stay in global memory (L1 / L2) even if they would fit the CTA's
  // shared memory -- a coder CTA that stages a big table keeps the transform kernels' CTAs off its SM.
  int64_t cap = budget;
  if (knobs().table_smem_kb >= 0) cap = static_cast<int64_t>(knobs().table_smem_kb) * 1024;
  t->in_smem = in_smem && static_cast<int64_t>(h.total_bytes) <= cap;
  t->enc_in_smem = static_cast<int64_t>(h.enc_bytes) <= budget && static_cast<int64_t>(h.enc_bytes) <= cap;
  t->device = dp.device;
  *out = t;
  return CAI_OK;
}

void cai_table_destroy(cai_table_t t) {
  if (!t) return;
  if (t->blob) cudaFree(t->blob);
  if (t->enc_params) cudaFree(t->enc_params);
  delete t;
}

int cai_table_info(cai_table_t t, int32_t *K, int64_t *blob_bytes, int32_t *lut_buckets, int32_t *in_smem) {
  CAI_CHECK_ARG(t != nullptr, "cai_table_info: NULL table");
  if (K) *K = t->K;
  if (blob_bytes) *blob_bytes = t->blob_bytes;
  if (lut_buckets) *lut_buckets = t->lut_buckets;
  if (in_smem) *in_smem = t->in_smem;
  return CAI_OK;
}

}  // extern "C"
