// common.cuh -- shared helpers for libcai_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/cai_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcai_b200 is written for sm_100a (B200) only"
#endif

namespace cai {

// ---- host error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define CAI_CHECK_ARG(cond, ...)      \
  do {                                \
    if (!(cond)) {                    \
      ::cai::set_error(__VA_ARGS__);  \
      return CAI_E_INVALID;           \
    }                                 \
  } while (0)

#define CAI_CUDA(call)                                                                    \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      ::cai::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                         \
      return CAI_E_CUDA;                                                                  \
    }                                                                                     \
  } while (0)

#define CAI_LAUNCH_CHECK() CAI_CUDA(cudaGetLastError())

struct DeviceProps {
  int device = -1;
  int sm_count = 0;
  int max_smem_optin = 0;
};
// Cached per-thread view of the current device (cudaGetDevice + attributes).
int get_device_props(DeviceProps *out);

// Opt `kernel` into the device's full dynamic shared memory (opt-in maximum minus the kernel's static shared memory)
// ONCE per (device, kernel), under a mutex.  cudaFuncAttributeMaxDynamicSharedMemorySize is per-function state
// shared by all host threads: setting it to each launch's own size let a concurrent caller LOWER it between another
// thread's set and launch ("invalid argument").  It is never lowered here, so the entry points are re-entrant.
// `max_dynamic` (optional) receives the limit that was set.  `carveout` >= 0 additionally sets the preferred
// shared-memory carve-out (percent), also once.
int optin_max_smem(const void *kernel, const DeviceProps &dp, int *max_dynamic = nullptr, int carveout = -1);

// Experiment knobs (environment), read ONCE at first use -- never per launch.
struct Knobs {
  int conv_stages = 0;      // CAI_CONV_STAGES   (0 = automatic)
  int conv_generic = 0;     // CAI_CONV_GENERIC  (force the all-runtime-flags instantiation)
  int conv_carveout = -1;   // CAI_CONV_CARVEOUT (percent, -1 = driver default)
  int conv_debug = 0;       // CAI_CONV_DEBUG    (only honoured by -DCAI_DEBUG_BUILD builds)
  int patch_generic = 0;    // CAI_PATCH_GENERIC (generic im2col / col2im kernels)
  int coder_warps = 0;      // CAI_CODER_WARPS   (0 = automatic)
  int lut_buckets = 0;      // CAI_LUT_BUCKETS   (0 = automatic)
  int table_smem_kb = -1;   // CAI_TABLE_SMEM_KB (-1 = no cap)
  int coder_lanes = 0;      // CAI_CODER_LANES   (N > 0: lane-per-string kernels from N strings per launch up; 0 = never, the default)
  int conv_persist = -1;    // CAI_CONV_PERSIST  (-1 = automatic)
  int coder_lut_adapt = -1; // CAI_LUT_ADAPT     (-1 = automatic)
  int tma_epi_warps = 0;    // CAI_TMA_EPI_WARPS (8 or 16 epilogue warps in the persistent transform kernel; 0 = 16)
};
const Knobs &knobs();

// ---- packed CDF table blob (see table.cu) ----------------------------------------------------------
constexpr uint32_t kBlobMagic = 0x43414931u;  // "CAI1"

struct BlobHeader {  // 64 bytes
  uint32_t magic;
  int32_t K;
  int32_t Lmax;
  uint32_t n_cdf_entries;  // uint16 entries in the cdf section (padded to a multiple of 8)
  uint32_t off_meta;       // byte offsets from the start of the blob, all multiples of 16
  uint32_t off_cdf;
  uint32_t off_lut;
  uint32_t total_bytes;
  int32_t lut_shift;    // bucket = cf >> lut_shift
  int32_t lut_buckets;  // per row, power of two
  uint32_t enc_bytes;   // header + meta + cdf: what the encoder stages (no LUT)
  uint32_t pad[5];
};
static_assert(sizeof(BlobHeader) == 64, "BlobHeader must be 64 bytes");

struct RowMeta {     // 16 bytes, read with one LDS.128
  uint32_t cdf_off;  // in uint16 units from the start of the cdf section
  int32_t len;       // _cdf_length[k] (>= 2 for a usable row)
  int32_t offset;    // _offset[k]
  uint32_t lut_off;  // in 8-byte LUT entries from the start of the LUT section
};
static_assert(sizeof(RowMeta) == 16, "RowMeta must be 16 bytes");

}  // namespace cai

struct cai_table {
  unsigned char *blob = nullptr;  // device
  int64_t blob_bytes = 0;
  uint32_t enc_bytes = 0;
  int32_t K = 0;
  int32_t Lmax = 0;
  int32_t lut_shift = 0;
  int32_t lut_buckets = 0;
  int in_smem = 0;      // whole blob (with LUT) fits the decoder's shared memory budget
  int enc_in_smem = 0;  // header + meta + cdf fits the encoder's budget
  int device = -1;
  uint32_t n_cdf_entries = 0;
  // per-(row, symbol) encoder parameters {reciprocal, bias | shift, freq << 15} for the lane-per-string encoder,
  // built on first use (rans.cu); 16 bytes per cdf entry, read through L1 / L2 one symbol ahead of the chain
  void *enc_params = nullptr;
};

// ---- device-side PTX wrappers ------------------------------------------------------------------------
#ifdef __CUDACC__
namespace cai {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(phase)
        : "memory");
  }
}
// TMA bulk (non-tensor) copy global -> shared, completion on an mbarrier.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Stage `bytes` (multiple of 16) of a blob into shared memory with TMA bulk copies issued by thread 0.
// Must be called by every thread of the CTA; returns after the data is visible to all of them.
__device__ __forceinline__ void stage_blob(unsigned char *dst_smem, const unsigned char *src, uint32_t bytes,
                                           uint64_t *bar) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, bytes);
    constexpr uint32_t kPiece = 32768;
    for (uint32_t o = 0; o < bytes; o += kPiece) {
      const uint32_t nb = (bytes - o < kPiece) ? (bytes - o) : kPiece;
      tma_bulk_g2s(dst_smem + o, src + o, nb, bar);
    }
  }
  mbar_wait(bar, 0);
}

}  // namespace cai
#endif
