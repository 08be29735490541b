// conv_tma.cu -- persistent, TMA-fed implicit-GEMM transform kernel for sm_100a (the wide layers of g_a / g_s).
//
// Same math and the same reference lines as conv.cu (compressai/models/utils.py:128-146 conv / deconv,
// compressai/layers/gdn.py:77-92 GDN / IGDN): D[pixel, cout] = sum over (tap, cin) of A * W with split-bf16 operands
// (three tcgen05.mma per k16), optional fused GDN / IGDN as a second in-kernel GEMM.  What differs is how the operands
// reach the tensor core and how tiles are scheduled:
//
//   * Tiles are ROW SEGMENTS: 128 accumulator rows = S <= 128 consecutive pixels of ONE row of the output (phase)
//     grid.  All taps that share dy and whose dx differ by multiples of the input stride ("tap group") then read the
//     same input row shifted by whole pixels, so ONE copy of that row segment (S + group length - 1 pixels) serves
//     every tap of the group: the tap is selected by the start address of the UMMA shared-memory descriptor (+16 B
//     per pixel in the K-major, no-swizzle canonical layout).  The copy is one TMA tensor load per plane
//     (cp.async.bulk.tensor.5d, SASS UTMALDG) over the view {8 channels, W, C/8, H, N} of the NHWC plane with box
//     {8, pixels, 4 chunks, 1, 1}: it lands as [chunk][pixel][16 B] -- exactly the canonical layout -- with the
//     conv stride as TMA element stride and image borders zero-filled by the hardware.  No LDGSTS, no L1, no
//     per-thread address arithmetic: one elected thread issues every load.
//   * The CTA is PERSISTENT (one per SM) with TWO accumulator sets in TMEM: while the eight epilogue warps drain
//     tile t (GDN operand, norm GEMM, output, copy-out), the producer and MMA threads already run tile t + 1's
//     main loop.  Operand rings are deep (A: row slots, B: one 16 KB weight slab per tap and 32-channel chunk),
//     gamma stays resident in shared memory for the kernel's lifetime, and the GDN products of tile t are issued
//     opportunistically between the k-steps of tile t + 1 as soon as their x^2 slabs are written.
//
// k-step order: channel chunk (32) -> tap group -> tap in group; weights are packed in that order by the host
// (transforms.pack_weights(order="chunk")).  Layers this kernel does not take (narrow grids, more than 128 output
// channels, GDN finalize with aux planes) stay on conv.cu's kernel: cai_conv_tma_eligible() is the single test.
#include <cuda.h>

#include <mutex>

#include "conv_common.cuh"

namespace cai {

// Epilogue warps per CTA: template parameter EW = 8 or 16 = one or two TEAMS of 8 warps.  A team drains one tile: two
// groups of 128 threads (thread = TMEM lane = pixel row; group h takes alternate 32-column slabs and owns one x^2 slab
// buffer).  A tile's epilogue is a dependent chain (x^2 slabs -> norm GEMM -> output math -> staging -> copy-out, with
// two named barriers): more warps on ONE tile do not shorten it (measured: 16 warps on one tile 1.273 ms vs 8 warps
// 1.295 ms for the first layer), so with EW = 16 the two teams take ALTERNATE tiles -- team t owns accumulator set t,
// slab buffers 2t / 2t+1, staging region t and named barrier 1 + t -- and one team's waits overlap the other's stores.
// Two more warps follow: warp EW = TMA producer, warp EW + 1 = MMA issuer (owns TMEM).
constexpr int kTmaMaxGroups = 4;
constexpr int kMaxRing = 8;

struct TmaConvParams {
  const unsigned char *w_packed;  // [kchunk][tap (grouped order)][hi | lo] each BN x 32 bf16, canonical layout
  const float *bias;
  float *out_f32;
  __nv_bfloat16 *out_hi, *out_lo, *abs_hi, *abs_lo;
  const unsigned char *gdn_w;
  const float *gdn_beta;
  int gdn_mode;
  int N, H, W, Cin, Ho, Wo, Cout;
  int Hp, Wp, os, o0y, o0x, is;
  int ntaps, ngroups, kchunks, BN, epilogue;
  float clamp_lo, clamp_hi;
  int S, segs, ppx;          // valid pixels per tile, tiles per grid row, pixels per A slot row
  int a_slots, b_slots;
  int ntiles;
  uint32_t a_slot_bytes, off_a, off_b, off_e, e_bytes;  // shared-memory carve-up (gamma at offset 0)
  int8_t dy[kMaxTaps], dx[kMaxTaps], glen[kMaxTaps];
};

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t phase) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
      "r"(smem_u32(bar))
      : "memory");
}

// KIND: 1 fused GDN / IGDN -> split planes; 2 linear / ReLU / LeakyReLU -> split planes; 3 -> fp32 (+ |.| planes, clamp)
template <int KIND, int EW>
__global__ void __launch_bounds__(EW * 32 + 64, 1)
conv_tma_kernel(const __grid_constant__ TmaConvParams p, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t a_full[kMaxRing], a_empty[kMaxRing], b_full[kMaxRing], b_empty[kMaxRing];
  __shared__ __align__(8) uint64_t acc_full[2], acc_free[2], slab_full[kTmaMaxGroups], slab_free[kTmaMaxGroups], norm_full[2], gamma_bar;
  constexpr int kTmaEpiWarps = EW;
  constexpr int kTmaEpiThreads = EW * 32;
  constexpr int kTeams = EW / 8;
  constexpr int kGroups = 2 * kTeams;  // slab buffers: two per team
  static_assert(kTeams == 1 || kTeams == 2, "EW must be 8 or 16");
  __shared__ uint32_t s_tmem_base;
  __shared__ int64_t s_opix[2][kBM];
  __shared__ __align__(16) float s_bias[128];
  __shared__ __align__(16) float s_beta[128];

  constexpr bool kGdn = KIND == 1;
  constexpr bool kF32 = KIND == 3;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  const uint32_t b_plane = static_cast<uint32_t>(BN) * kBK * 2;  // one bf16 plane of a weight slab: BN x 32
  const uint32_t a_plane = 4u * static_cast<uint32_t>(p.ppx) * 16u;  // one plane of an A slot: 4 chunks x ppx pixels
  const uint32_t slab_plane = (kBK / 8) * kLboA;                  // one plane of an x^2 slab (GDN operand)
  const int gk = kGdn ? (BN + kBK - 1) / kBK : 0;                 // GDN k-steps (32-column slabs) per tile
  uint32_t acc_cols = 32;
  while (acc_cols < static_cast<uint32_t>(BN)) acc_cols <<= 1;
  const uint32_t set_cols = kGdn ? 2 * acc_cols : acc_cols;       // main (+ norm) accumulator of one tile
  const uint32_t tmem_cols = 2 * set_cols;                        // two tiles in flight
  unsigned char *sm_gamma = smem;
  unsigned char *sm_a = smem + p.off_a;
  unsigned char *sm_b = smem + p.off_b;
  unsigned char *sm_e = smem + p.off_e;  // x^2 slabs (two) during the GDN phase, output staging afterwards

  if (tid == 0) {
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_free[s], 1);
      mbar_init(&norm_full[s], 1);
    }
    for (int s = 0; s < kGroups; ++s) {
      mbar_init(&slab_full[s], 128);
      mbar_init(&slab_free[s], 1);
    }
    mbar_init(&gamma_bar, 1);
    mbar_fence_init();
  }
  if (warp == kTmaEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < kTmaEpiWarps) {
    for (int i = tid; i < 128; i += kTmaEpiThreads) {
      const bool in = i < p.Cout && i < BN;
      s_bias[i] = (p.bias && in) ? __ldg(p.bias + i) : 0.f;
      s_beta[i] = (kGdn && in) ? __ldg(p.gdn_beta + i) : 1.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int per_row = p.segs;

  if (warp == kTmaEpiWarps) {
    // ===================== producer: one thread issues every operand load =====================
    if (lane == 0) {
      if (kGdn) {  // gamma: resident for the whole kernel
        const uint32_t gbytes = static_cast<uint32_t>(gk) * 2u * b_plane;
        mbar_expect_tx(&gamma_bar, gbytes);
        for (uint32_t o = 0; o < gbytes; o += 32768u)
          tma_bulk_g2s(sm_gamma + o, p.gdn_w + o, (gbytes - o < 32768u) ? (gbytes - o) : 32768u, &gamma_bar);
      }
      uint32_t a_it = 0, b_it = 0;
      int sa = 0, sb = 0;  // ring positions of a_it / b_it
      uint32_t a_pass = 0, b_pass = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int seg = tile % per_row;
        const int rowid = tile / per_row;
        const int gi = rowid % p.Hp, n = rowid / p.Hp;
        const int j0 = seg * p.S;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          int t0 = 0;
          for (int g = 0; g < p.ngroups; ++g) {
            if (a_pass > 0) mbar_wait_bounded(&a_empty[sa], (a_pass - 1) & 1);
            unsigned char *dst = sm_a + static_cast<uint32_t>(sa) * p.a_slot_bytes;
            mbar_expect_tx(&a_full[sa], 2u * a_plane);
            const int x0 = p.is * j0 + p.dx[t0];
            const int y = p.is * gi + p.dy[t0];
            tma_load_5d(dst, &map_hi, 0, x0, kc * 4, y, n, &a_full[sa]);
            tma_load_5d(dst + a_plane, &map_lo, 0, x0, kc * 4, y, n, &a_full[sa]);
            if (++sa == p.a_slots) {
              sa = 0;
              ++a_pass;
            }
            ++a_it;
            const int gl = p.glen[g];
            for (int u = 0; u < gl; ++u) {
              if (b_pass > 0) mbar_wait_bounded(&b_empty[sb], (b_pass - 1) & 1);
              mbar_expect_tx(&b_full[sb], 2u * b_plane);
              tma_bulk_g2s(sm_b + static_cast<uint32_t>(sb) * (2u * b_plane),
                           p.w_packed + static_cast<size_t>(kc * p.ntaps + t0 + u) * (2u * b_plane), 2u * b_plane,
                           &b_full[sb]);
              if (++sb == p.b_slots) {
                sb = 0;
                ++b_pass;
              }
              ++b_it;
            }
            t0 += gl;
          }
        }
      }
    }
  } else if (warp == kTmaEpiWarps + 1) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0) {
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                             (static_cast<uint32_t>(kBM >> 4) << 24);
      const uint32_t lbo_a = static_cast<uint32_t>(p.ppx) * 16u, lbo_b = static_cast<uint32_t>(BN) * 16u;
      int sa = 0, sb = 0;
      uint32_t a_par = 0, b_par = 0;
      // GDN products still owed to an earlier tile (its x^2 slabs are written by the epilogue warps)
      int g_next = gk, g_set = 0;
      uint32_t g_use[kGroups] = {};  // completed uses of each slab buffer (phase of slab_full)
      auto gdn_issue = [&](bool block) {
        while (g_next < gk) {
          const int gteam = (kTeams == 2) ? g_set : 0;  // the pending tile's team = its accumulator set
          const int buf = 2 * gteam + (g_next & 1);
          if (block) mbar_wait_bounded(&slab_full[buf], g_use[buf] & 1);
          else if (!mbar_test(&slab_full[buf], g_use[buf] & 1)) return;
          tc_fence_after();
          if (g_next == 0) {
            mbar_wait_bounded(&gamma_bar, 0);  // immediate after the first tile
          }
          const uint32_t d_norm = tmem_base + static_cast<uint32_t>(g_set) * set_cols + acc_cols;
          const uint32_t sl = smem_u32(sm_e) + static_cast<uint32_t>(gteam) * p.e_bytes +
                              static_cast<uint32_t>(g_next & 1) * (2u * slab_plane);
          const uint32_t gb = smem_u32(sm_gamma) + static_cast<uint32_t>(g_next) * (2u * b_plane);
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint64_t dah = make_smem_desc(sl + kk * 2 * kLboA, kLboA, 128);
            const uint64_t dal = make_smem_desc(sl + slab_plane + kk * 2 * kLboA, kLboA, 128);
            const uint64_t dbh = make_smem_desc(gb + kk * 2 * lbo_b, lbo_b, 128);
            const uint64_t dbl = make_smem_desc(gb + b_plane + kk * 2 * lbo_b, lbo_b, 128);
            umma_bf16(d_norm, dah, dbh, idesc, (g_next > 0 || kk > 0) ? 1u : 0u);
            umma_bf16(d_norm, dah, dbl, idesc, 1u);
            umma_bf16(d_norm, dal, dbh, idesc, 1u);
          }
          umma_commit(&slab_free[buf]);
          ++g_use[buf];
          if (++g_next == gk) umma_commit(&norm_full[g_set]);
        }
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int set = it & 1;
        if (it >= 2) {  // the epilogue of tile it - 2 must have drained this accumulator set
          const uint32_t par = static_cast<uint32_t>((it >> 1) - 1) & 1u;
          if (kGdn) {
            // keep serving the OTHER team's norm GEMM while this team finishes its output phase: blocking here would
            // chain the two teams' epilogues through this thread (measured: two teams then gain only 7 %)
            uint32_t spins = 0;
            while (!mbar_test(&acc_free[set], par)) {
              gdn_issue(false);
              if (++spins > kSpinLimit) spin_fail();
            }
          } else {
            mbar_wait_bounded(&acc_free[set], par);
          }
          tc_fence_after();
        }
        const uint32_t d_main = tmem_base + static_cast<uint32_t>(set) * set_cols;
        uint32_t accumulate = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int g = 0; g < p.ngroups; ++g) {
            mbar_wait_bounded(&a_full[sa], a_par);
            const uint32_t a_base = smem_u32(sm_a) + static_cast<uint32_t>(sa) * p.a_slot_bytes;
            const int gl = p.glen[g];
            for (int u = 0; u < gl; ++u) {
              mbar_wait_bounded(&b_full[sb], b_par);
              tc_fence_after();
              const uint32_t a_hi = a_base + static_cast<uint32_t>(u) * 16u, a_lo = a_hi + a_plane;
              const uint32_t b_hi = smem_u32(sm_b) + static_cast<uint32_t>(sb) * (2u * b_plane), b_lo = b_hi + b_plane;
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk) {
                const uint64_t dah = make_smem_desc(a_hi + kk * 2 * lbo_a, lbo_a, 128);
                const uint64_t dal = make_smem_desc(a_lo + kk * 2 * lbo_a, lbo_a, 128);
                const uint64_t dbh = make_smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, 128);
                const uint64_t dbl = make_smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, 128);
                umma_bf16(d_main, dah, dbh, idesc, accumulate);
                umma_bf16(d_main, dah, dbl, idesc, 1u);
                umma_bf16(d_main, dal, dbh, idesc, 1u);
                accumulate = 1u;
              }
              umma_commit(&b_empty[sb]);  // frees the weight slab when the MMAs have read it
              if (++sb == p.b_slots) {
                sb = 0;
                b_par ^= 1u;
              }
              if (kGdn) gdn_issue(false);
            }
            umma_commit(&a_empty[sa]);    // frees the row slot
            if (++sa == p.a_slots) {
              sa = 0;
              a_par ^= 1u;
            }
          }
        }
        umma_commit(&acc_full[set]);
        if (kGdn) {
          gdn_issue(true);  // whatever is still owed to the previous tile
          g_next = 0;       // this tile's norm GEMM becomes pending
          g_set = set;
        }
      }
      if (kGdn) gdn_issue(true);
    }
  } else {
    // ===================== epilogue warps: thread = TMEM lane = pixel row r, group h takes alternate slabs ==========
    const int team = (kTeams == 2) ? (tid >> 8) : 0;
    const int tt = tid & 255;  // thread within the team
    const int r = tt & (kBM - 1);
    const int h = tt >> 7;
    const int sbuf = 2 * team + h;  // this group's x^2 slab buffer
    unsigned char *sm_et = sm_e + static_cast<uint32_t>(team) * p.e_bytes;  // the team's slab / staging region
    int64_t *opix_t = s_opix[team];
    const uint32_t lane_base = (static_cast<uint32_t>(warp & 3) * 32u) << 16;
    const uint32_t row_off = (static_cast<uint32_t>(r) >> 3) * 128u + (static_cast<uint32_t>(r) & 7u) * 16u;
    const int epi = kGdn ? 0 : p.epilogue;
    const bool has_abs = kF32 && p.abs_hi != nullptr;
    const bool do_clamp = kF32 && (p.clamp_lo < p.clamp_hi);
    // staging geometry (fixed for the launch)
    const int n_pl = (kF32 ? 0 : 1) + (has_abs ? 1 : 0);
    int ncols = BN;
    while (ncols > 16 && kBM * ((kF32 ? (ncols * 4u + 16u) : 0u) + n_pl * 2u * (ncols * 2u + 16u)) > p.e_bytes) ncols -= 16;
    const uint32_t pitch_f = ncols * 4u + 16u, pitch_b = ncols * 2u + 16u;
    const uint32_t off_out = kF32 ? kBM * pitch_f : 0u;  // first plane pair (out planes, or |out| planes for KIND 3)
    uint32_t slab_uses = 0;  // times this group has filled its slab buffer (buffer index = h)
    int it = team;  // team t drains the tiles it = t, t + kTeams, ... of this CTA's sequence
    for (int tile = blockIdx.x + team * static_cast<int>(gridDim.x); tile < p.ntiles;
         tile += kTeams * static_cast<int>(gridDim.x), it += kTeams) {
      const int set = it & 1;
      const uint32_t use_par = static_cast<uint32_t>(it >> 1) & 1u;
      const int seg = tile % per_row;
      const int rowid = tile / per_row;
      const int gi = rowid % p.Hp, n = rowid / p.Hp;
      const int j = seg * p.S + r;
      const bool row_ok = r < p.S && j < p.Wp;
      int64_t opix = -1;
      if (row_ok) opix = (static_cast<int64_t>(n) * p.Ho + (gi * p.os + p.o0y)) * p.Wo + (j * p.os + p.o0x);
      const uint32_t t_main = tmem_base + static_cast<uint32_t>(set) * set_cols + lane_base;
      mbar_wait_bounded(&acc_full[set], use_par);
      tc_fence_after();
      if (kGdn) {
        // x = acc + bias; x^2 split into bf16 planes -> slab buffer h, one 32-column slab per GDN k-step
        uint32_t raw[32];
#pragma unroll 1
        for (int g = h; g < gk; g += 2) {
          if (slab_uses > 0) mbar_wait_bounded(&slab_free[sbuf], (slab_uses - 1) & 1);
          unsigned char *sl = sm_et + static_cast<uint32_t>(h) * (2u * slab_plane);
          const int col0 = g * kBK;
          tmem_ld32_nowait(t_main + static_cast<uint32_t>(col0), raw);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < kBK / 8; ++c) {
            const int col = col0 + c * 8;
            uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = make_uint4(0u, 0u, 0u, 0u);
            if (col < BN) {
              float sq[8];
              const float4 b0 = *reinterpret_cast<const float4 *>(s_bias + col);
              const float4 b1 = *reinterpret_cast<const float4 *>(s_bias + col + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a = __uint_as_float(raw[c * 8 + i]) + bb[i];
                sq[i] = row_ok ? a * a : 0.f;  // rows beyond the segment hold garbage: keep them finite
              }
              const Pack8 pk = split8(sq);
              vh = pk.hi;
              vl = pk.lo;
            }
            const uint32_t so = static_cast<uint32_t>(c) * kLboA + row_off;
            *reinterpret_cast<uint4 *>(sl + so) = vh;
            *reinterpret_cast<uint4 *>(sl + slab_plane + so) = vl;
          }
          fence_async_proxy();  // generic-proxy stores -> visible to the tensor core (async proxy)
          mbar_arrive(&slab_full[sbuf]);
          ++slab_uses;
        }
        mbar_wait_bounded(&norm_full[set], use_par);
        tc_fence_after();
      }
      if (h == 0) opix_t[r] = opix;
      // ---- output: TMEM -> registers -> math -> staging (thread = row), then cooperative full-line copy-out
      for (int cA = 0; cA < BN; cA += ncols) {
        const int cB = (cA + ncols < BN) ? cA + ncols : BN;
        const bool last_pass = cB == BN;
        uint32_t qa[16], qa2[16];
#pragma unroll 1
        for (int s0 = cA + 32 * h; s0 < cB; s0 += 64) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c0 = s0 + 16 * q;
            if (c0 >= cB) continue;
            tmem_ld16_nowait(t_main + static_cast<uint32_t>(c0), qa);
            if (kGdn) tmem_ld16_nowait(t_main + acc_cols + static_cast<uint32_t>(c0), qa2);
            tmem_wait_ld();
            if (!row_ok || c0 >= p.Cout) continue;
            float v[16];
            {
              const float4 *b4 = reinterpret_cast<const float4 *>(s_bias + c0);
#pragma unroll
              for (int w4 = 0; w4 < 4; ++w4) {
                const float4 bb = b4[w4];
                v[4 * w4] = __uint_as_float(qa[4 * w4]) + bb.x;
                v[4 * w4 + 1] = __uint_as_float(qa[4 * w4 + 1]) + bb.y;
                v[4 * w4 + 2] = __uint_as_float(qa[4 * w4 + 2]) + bb.z;
                v[4 * w4 + 3] = __uint_as_float(qa[4 * w4 + 3]) + bb.w;
              }
            }
            if (kGdn) {
              const float4 *g4 = reinterpret_cast<const float4 *>(s_beta + c0);
#pragma unroll
              for (int w4 = 0; w4 < 4; ++w4) {
                const float4 bb = g4[w4];
                const float be[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float nrm = __uint_as_float(qa2[4 * w4 + i]) + be[i];
                  const float rs = rsqrtf(nrm);
                  v[4 * w4 + i] *= (p.gdn_mode == 1) ? rs : nrm * rs;  // n^-1/2 or n^+1/2 = n * n^-1/2
                }
              }
            }
            if (epi == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (epi == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
            }
            if (do_clamp) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(v[i], p.clamp_lo), p.clamp_hi);
            }
            const uint32_t cc = static_cast<uint32_t>(c0 - cA);
            if (kF32) {
              float4 *o = reinterpret_cast<float4 *>(sm_et + static_cast<uint32_t>(r) * pitch_f + cc * 4u);
#pragma unroll
              for (int w4 = 0; w4 < 4; ++w4) o[w4] = make_float4(v[4 * w4], v[4 * w4 + 1], v[4 * w4 + 2], v[4 * w4 + 3]);
            }
            const uint32_t rb = static_cast<uint32_t>(r) * pitch_b + cc * 2u;
            if (!kF32) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const Pack8 pk = split8(v + 8 * hh);
                *reinterpret_cast<uint4 *>(sm_et + off_out + rb + hh * 16u) = pk.hi;
                *reinterpret_cast<uint4 *>(sm_et + off_out + kBM * pitch_b + rb + hh * 16u) = pk.lo;
              }
            } else if (has_abs) {
              float sv[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) sv[i] = fabsf(v[i]);
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const Pack8 pk = split8(sv + 8 * hh);
                *reinterpret_cast<uint4 *>(sm_et + off_out + rb + hh * 16u) = pk.hi;
                *reinterpret_cast<uint4 *>(sm_et + off_out + kBM * pitch_b + rb + hh * 16u) = pk.lo;
              }
            }
          }
        }
        if (last_pass) tc_fence_before();  // last TMEM reads of this accumulator set
        asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory");
        if (last_pass && tt == 0) mbar_arrive(&acc_free[set]);  // tile it + 2 may overwrite the set
        // ---- cooperative copy-out: 16-byte units, consecutive lanes along a row (full-line stores)
        const int cols_here = (cB - cA < p.Cout - cA) ? (cB - cA) : (p.Cout - cA);
        if (cols_here > 0) {
          const int nbuf = (kF32 ? 1 : 0) + 2 * n_pl;
#pragma unroll 1
          for (int bi = 0; bi < nbuf; ++bi) {
            unsigned char *gptr;
            uint32_t soff, esize;
            if (kF32 && bi == 0) {
              gptr = reinterpret_cast<unsigned char *>(p.out_f32);
              soff = 0;
              esize = 4u;
            } else {
              const int q = bi - (kF32 ? 1 : 0);  // 0: hi plane, 1: lo plane
              __nv_bfloat16 *pl = kF32 ? (q == 0 ? p.abs_hi : p.abs_lo) : (q == 0 ? p.out_hi : p.out_lo);
              gptr = reinterpret_cast<unsigned char *>(pl);
              soff = off_out + static_cast<uint32_t>(q) * kBM * pitch_b;
              esize = 2u;
            }
            const uint32_t pitch = (esize == 4u) ? pitch_f : pitch_b;
            const uint32_t units = static_cast<uint32_t>(cols_here) * esize / 16u;  // per row
            unsigned char *gbase = gptr + static_cast<int64_t>(cA) * esize;
            const int64_t row_stride = static_cast<int64_t>(p.Cout) * esize;
            const uint32_t total_u = kBM * units;
            for (uint32_t u = tt; u < total_u; u += 256u) {
              const uint32_t row = u / units, jj = u - row * units;
              const int64_t op = opix_t[row];
              if (op < 0) continue;
              const uint4 val = *reinterpret_cast<const uint4 *>(sm_et + soff + row * pitch + jj * 16u);
              *reinterpret_cast<uint4 *>(gbase + op * row_stride + jj * 16u) = val;
            }
          }
        }
        asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory");
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTmaEpiWarps + 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// epilogue warps of the persistent kernel: 16 unless CAI_TMA_EPI_WARPS=8 (experiments)
static int tma_epi_warps() { return knobs().tma_epi_warps == 8 ? 8 : 16; }

struct TmaPlan {
  int S, segs, ppx, a_slots, b_slots, ngroups, maxg;
  uint32_t a_slot_bytes, off_a, off_b, off_e, e_bytes, smem;
};

// The single eligibility test (also exported through cai_conv_tma_eligible): geometry the persistent kernel takes,
// and the shared-memory carve-up it would use.  `kind` as chosen by cai_conv_gemm (1, 2 or 3).
static bool plan_tma(const cai_conv_desc *d, int kind, int max_smem, TmaPlan *pl, int epi_warps) {
  if (knobs().conv_persist == 0) return false;
  if (kind < 1 || kind > 3) return false;
  if (d->BN != d->Cout || d->BN > 128 || d->BN % 16 || d->Cin % 8) return false;
  if (d->is != 1 && d->is != 2) return false;
  if (d->Wp < 64 || d->glen[0] == 0) return false;
  if ((reinterpret_cast<uintptr_t>(d->a_hi) | reinterpret_cast<uintptr_t>(d->a_lo)) & 15u) return false;
  if (static_cast<int64_t>(d->W) * d->Cin * 2 % 16) return false;
  // groups: same dy, dx ascending in steps of `is`
  int covered = 0, ng = 0, maxg = 1;
  for (int g = 0; g < kMaxTaps && covered < d->ntaps; ++g) {
    const int n = d->glen[g];
    if (n < 1 || covered + n > d->ntaps) return false;
    for (int t = covered + 1; t < covered + n; ++t)
      if (d->dy[t] != d->dy[covered] || d->dx[t] != d->dx[t - 1] + d->is) return false;
    if (n > maxg) maxg = n;
    covered += n;
    ++ng;
  }
  if (covered != d->ntaps) return false;
  int segs = (d->Wp + kBM - 1) / kBM;
  int S = (d->Wp + segs - 1) / segs;  // TMA box extent along W: pixels * stride <= 256
  // pixels per row slot, kept even so that both planes of a slot start on 128-byte boundaries (TMA destination)
  int ppx = (S + maxg) & ~1;
  while (ppx * d->is > 256) {
    ++segs;
    S = (d->Wp + segs - 1) / segs;
    ppx = (S + maxg) & ~1;
  }
  pl->S = S;
  pl->segs = segs;
  pl->maxg = maxg;
  pl->ngroups = ng;
  pl->ppx = ppx;
  const uint32_t a_plane = 4u * static_cast<uint32_t>(pl->ppx) * 16u;
  pl->a_slot_bytes = (2u * a_plane + 127u) & ~127u;
  const uint32_t b_slot = 2u * static_cast<uint32_t>(d->BN) * kBK * 2u;
  const uint32_t gamma = kind == 1 ? static_cast<uint32_t>((d->BN + kBK - 1) / kBK) * b_slot : 0u;
  const uint32_t teams = static_cast<uint32_t>(epi_warps / 8);  // each team has its own slab / staging region
  const uint32_t slabs = kind == 1 ? 2u * 2u * (kBK / 8) * kLboA : 0u;
  // staging: full tile when it fits, else column passes of >= 64 (planes) / 32 (fp32) columns
  uint32_t e_full;
  if (kind == 3)
    e_full = kBM * ((d->BN * 4u + 16u) + (d->abs_hi ? 2u * (d->BN * 2u + 16u) : 0u));
  else
    e_full = kBM * 2u * (d->BN * 2u + 16u);
  const uint32_t e_min = kind == 3 ? kBM * ((32u * 4u + 16u) + (d->abs_hi ? 2u * (32u * 2u + 16u) : 0u))
                                   : kBM * 2u * (64u * 2u + 16u);
  const int ksteps_tile = ((d->Cin + kBK - 1) / kBK) * d->ntaps;
  const uint32_t budget = static_cast<uint32_t>(max_smem) - 1024u;  // alignment slack
  // rings: as deep as the budget allows up to kMaxRing, at least 2 row slots and 3 weight slabs
  for (uint32_t e_bytes : {e_full, e_min}) {
    uint32_t eb = e_bytes > slabs ? e_bytes : slabs;
    eb = (eb + 127u) & ~127u;
    if (gamma + teams * eb + 2u * pl->a_slot_bytes + 3u * b_slot > budget) continue;
    uint32_t rest = budget - gamma - teams * eb;
    int a_slots = 2, b_slots = 3;
    rest -= 2u * pl->a_slot_bytes + 3u * b_slot;
    // grow the weight ring first (one slab per tap: the finest-grained consumer), then the row ring
    while (true) {
      bool grew = false;
      if (b_slots < kMaxRing && b_slots < ksteps_tile * 2 && rest >= b_slot) {
        ++b_slots;
        rest -= b_slot;
        grew = true;
      }
      if (a_slots < 4 && rest >= pl->a_slot_bytes && b_slots >= 5) {
        ++a_slots;
        rest -= pl->a_slot_bytes;
        grew = true;
      }
      if (!grew) break;
    }
    pl->a_slots = a_slots;
    pl->b_slots = b_slots;
    pl->off_a = (gamma + 127u) & ~127u;
    pl->off_b = pl->off_a + static_cast<uint32_t>(a_slots) * pl->a_slot_bytes;
    pl->off_e = pl->off_b + static_cast<uint32_t>(b_slots) * b_slot;
    pl->e_bytes = eb;
    pl->smem = pl->off_e + teams * eb;
    return pl->smem <= static_cast<uint32_t>(max_smem);
  }
  return false;
}

static int conv_kind(const cai_conv_desc *d) {
  const bool only_planes = d->out_hi && !d->out_f32 && !d->sq_hi && !d->abs_hi;
  const bool no_clamp = !(d->clamp_lo < d->clamp_hi);
  if (d->gdn_w && only_planes && no_clamp && d->epilogue == 0) return 1;
  if (!d->gdn_w && d->epilogue <= 2 && only_planes && no_clamp) return 2;
  if (!d->gdn_w && d->epilogue <= 2 && d->out_f32 && !d->out_hi && !d->sq_hi) return 3;
  return 0;
}

// Launch on the persistent kernel.  Returns CAI_OK, an error, or 1 when the layer is not eligible (caller falls back).
int launch_conv_tma(const cai_conv_desc *d, cudaStream_t st) {
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  const int kind = conv_kind(d);
  if (kind == 0) return 1;
  const int ew = tma_epi_warps();
  const void *fn8[3] = {reinterpret_cast<const void *>(conv_tma_kernel<1, 8>), reinterpret_cast<const void *>(conv_tma_kernel<2, 8>),
                        reinterpret_cast<const void *>(conv_tma_kernel<3, 8>)};
  const void *fn16[3] = {reinterpret_cast<const void *>(conv_tma_kernel<1, 16>), reinterpret_cast<const void *>(conv_tma_kernel<2, 16>),
                         reinterpret_cast<const void *>(conv_tma_kernel<3, 16>)};
  const void *fn = (ew == 16 ? fn16 : fn8)[kind - 1];
  int max_dyn = 0;
  rc = optin_max_smem(fn, dp, &max_dyn);
  if (rc != CAI_OK) return rc;
  TmaPlan pl;
  if (!plan_tma(d, kind, max_dyn, &pl, ew)) return 1;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return 1;

  CUtensorMap maps[2];
  const cuuint64_t gdim[5] = {8, static_cast<cuuint64_t>(d->W), static_cast<cuuint64_t>(d->Cin / 8),
                              static_cast<cuuint64_t>(d->H), static_cast<cuuint64_t>(d->N)};
  const cuuint64_t gstr[4] = {static_cast<cuuint64_t>(d->Cin) * 2, 16, static_cast<cuuint64_t>(d->W) * d->Cin * 2,
                              static_cast<cuuint64_t>(d->H) * d->W * d->Cin * 2};
  const cuuint32_t box[5] = {8, static_cast<cuuint32_t>(pl.ppx * d->is), 4, 1, 1};
  const cuuint32_t estr[5] = {1, static_cast<cuuint32_t>(d->is), 1, 1, 1};
  const void *planes[2] = {d->a_hi, d->a_lo};
  for (int q = 0; q < 2; ++q) {
    const CUresult r = enc(&maps[q], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void *>(planes[q]), gdim, gstr, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed (%d) for W=%d C=%d H=%d N=%d box=%u", static_cast<int>(r), d->W, d->Cin,
                d->H, d->N, box[1]);
      return CAI_E_CUDA;
    }
  }

  TmaConvParams p{};
  p.w_packed = static_cast<const unsigned char *>(d->w_packed);
  p.bias = d->bias;
  p.out_f32 = d->out_f32;
  p.out_hi = static_cast<__nv_bfloat16 *>(d->out_hi);
  p.out_lo = static_cast<__nv_bfloat16 *>(d->out_lo);
  p.abs_hi = static_cast<__nv_bfloat16 *>(d->abs_hi);
  p.abs_lo = static_cast<__nv_bfloat16 *>(d->abs_lo);
  p.gdn_w = static_cast<const unsigned char *>(d->gdn_w);
  p.gdn_beta = d->gdn_beta;
  p.gdn_mode = d->gdn_mode;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
  p.Hp = d->Hp; p.Wp = d->Wp; p.os = d->os; p.o0y = d->o0y; p.o0x = d->o0x; p.is = d->is;
  p.ntaps = d->ntaps;
  p.ngroups = pl.ngroups;
  p.kchunks = (d->Cin + kBK - 1) / kBK;
  p.BN = d->BN;
  p.epilogue = d->epilogue;
  p.clamp_lo = d->clamp_lo;
  p.clamp_hi = d->clamp_hi;
  p.S = pl.S; p.segs = pl.segs; p.ppx = pl.ppx;
  p.a_slots = pl.a_slots; p.b_slots = pl.b_slots;
  p.a_slot_bytes = pl.a_slot_bytes; p.off_a = pl.off_a; p.off_b = pl.off_b; p.off_e = pl.off_e; p.e_bytes = pl.e_bytes;
  const int64_t ntiles = static_cast<int64_t>(d->N) * d->Hp * pl.segs;
  CAI_CHECK_ARG(ntiles < (1ll << 31), "cai_conv_gemm: too many tiles");
  p.ntiles = static_cast<int>(ntiles);
  for (int t = 0; t < d->ntaps; ++t) {
    p.dy[t] = d->dy[t];
    p.dx[t] = d->dx[t];
  }
  for (int g = 0; g < pl.ngroups; ++g) p.glen[g] = d->glen[g];
  const int grid = ntiles < dp.sm_count ? static_cast<int>(ntiles) : dp.sm_count;
  const int threads = ew * 32 + 64;
  if (ew == 16) {
    switch (kind) {
      case 1: conv_tma_kernel<1, 16><<<grid, threads, pl.smem, st>>>(p, maps[0], maps[1]); break;
      case 2: conv_tma_kernel<2, 16><<<grid, threads, pl.smem, st>>>(p, maps[0], maps[1]); break;
      default: conv_tma_kernel<3, 16><<<grid, threads, pl.smem, st>>>(p, maps[0], maps[1]); break;
    }
  } else {
    switch (kind) {
      case 1: conv_tma_kernel<1, 8><<<grid, threads, pl.smem, st>>>(p, maps[0], maps[1]); break;
      case 2: conv_tma_kernel<2, 8><<<grid, threads, pl.smem, st>>>(p, maps[0], maps[1]); break;
      default: conv_tma_kernel<3, 8><<<grid, threads, pl.smem, st>>>(p, maps[0], maps[1]); break;
    }
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

bool conv_tma_eligible(const cai_conv_desc *d) {
  DeviceProps dp;
  if (get_device_props(&dp) != CAI_OK) return false;
  const int kind = conv_kind(d);
  if (kind == 0 || !encode_tiled_fn()) return false;
  TmaPlan pl;
  return plan_tma(d, kind, dp.max_smem_optin - 8192, &pl, tma_epi_warps());
}

}  // namespace cai
