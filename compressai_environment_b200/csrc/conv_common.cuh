// conv_common.cuh -- constants, launch parameters and PTX wrappers shared by the transform kernels
// (conv.cu: per-tile implicit GEMM with cp.async operands; conv_tma.cu: persistent TMA-fed variant).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace cai {

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kMaxGdnK = 8;       // GDN k-steps per tile (BN <= 256)
constexpr int kBK = 32;           // k elements per stage (2 x UMMA_K=16): small stages -> 2 CTAs per SM
constexpr int kProducerThreads = 64;   // warps 0-1 issue the A-tile cp.async copies (fewer mbarrier arrivals per k-step)
constexpr int kLoadIters = 8;          // rows per producer thread: kBM / (kProducerThreads / 4)
constexpr int kEpiWarps = 8;             // warps 0-7: two per TMEM lane quarter, taking alternate 32-column slabs
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kConvThreads = kEpiThreads + 32;  // + warp 8: MMA issuer, owns the TMEM allocation
constexpr int kMaxTaps = 32;
// A-tile chunk pitch (UMMA leading byte offset): 128 rows x 16 B + 32 B of padding so that a quarter warp that
// writes 2 pixels x 4 k-chunks touches 8 distinct 16-byte bank groups ((2c + p) mod 8) -> conflict-free LDGSTS
constexpr uint32_t kLboA = kBM * 16u + 32u;
constexpr uint32_t kSpinLimit = 1u << 28;


struct ConvKernelParams {
  const __nv_bfloat16 *a_hi, *a_lo;  // [N, H, W, Cin] bf16 planes
  const unsigned char *w_packed;     // [n_tiles][ksteps][hi | lo] each BN x kBK bf16, canonical layout
  const float *bias;                 // [Cout] or NULL
  const __nv_bfloat16 *aux_hi, *aux_lo;  // [M_out_pixels, Cout] planes for the GDN finalize, or NULL
  float *out_f32;                    // [N, Ho, Wo, Cout] or NULL
  __nv_bfloat16 *out_hi, *out_lo;    // split planes of the output, or NULL
  __nv_bfloat16 *sq_hi, *sq_lo;      // split planes of output^2, or NULL
  __nv_bfloat16 *abs_hi, *abs_lo;    // split planes of |output|, or NULL
  int N, H, W, Cin, Ho, Wo, Cout;
  int Hp, Wp;                        // phase grid
  int os, o0y, o0x;                  // output pixel = (i * os + o0y, j * os + o0x)
  int is;                            // input pixel  = (i * is + dy[t], j * is + dx[t])
  int ntaps;
  int kchunks;                       // ceil(Cin / kBK)
  int BN;                            // output channels per tile (multiple of 16, <= 256)
  int epilogue;                      // 0 linear, 1 relu, 2 leaky relu (0.01), 3 gdn (aux * rsqrt), 4 igdn (aux * sqrt)
  float clamp_lo, clamp_hi;          // applied when clamp_lo < clamp_hi
  int stages;
  // fused GDN / IGDN (second in-kernel GEMM): norm = gamma . out^2 + beta ; out = out * rsqrt(norm) (or * sqrt)
  const unsigned char *gdn_w;        // packed gamma: [kc][hi | lo][BN x kBK bf16], NULL = no fusion
  const float *gdn_beta;             // [Cout]
  int gdn_mode;                      // 1 GDN, 2 IGDN
  int debug;                         // experiment bitmask: 1 skip A loads, 2 skip B loads, 4 skip MMAs
  int8_t dy[kMaxTaps], dx[kMaxTaps];
  int8_t glen[kMaxTaps];             // tap group lengths (taps of a group share dy; dx differ by multiples of `is`)
};

__device__ __forceinline__ void spin_fail() {
  asm volatile("trap;");
}

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t phase) {
  uint32_t done = 0, spins = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(phase)
        : "memory");
    if (!done && ++spins > kSpinLimit) spin_fail();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: canonical ((8, n), 2) : ((16 B, SBO), LBO)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

struct Pack8 {
  uint4 hi, lo;
};
__device__ __forceinline__ Pack8 split8(const float *v) {
  __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(v[i], h[i], l[i]);
  Pack8 p;
  p.hi = *reinterpret_cast<uint4 *>(h);
  p.lo = *reinterpret_cast<uint4 *>(l);
  return p;
}


// conv_tma.cu (host): launch on the persistent TMA-fed kernel; 1 = layer not eligible
int launch_conv_tma(const cai_conv_desc *d, cudaStream_t st);
bool conv_tma_eligible(const cai_conv_desc *d);

}  // namespace cai
