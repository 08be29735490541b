// conv.cu -- tcgen05 implicit-GEMM convolution / transposed convolution / GDN for sm_100a.
//
// What it replaces (reference tree): the cuDNN / ATen kernels behind
//   compressai/models/utils.py:128-146   conv (k5 s2 p2, k3 s1 p1) and deconv (k5 s2 p2 op1)
//   compressai/layers/gdn.py:77-92       GDN / IGDN  (1x1 conv of x^2 with gamma, + beta, rsqrt / sqrt, * x)
//   compressai/models/google.py:134-152, :219-254, :339-353   the g_a / g_s / h_a / h_s stacks
//
// Formulation.  Every layer is D[m, n] = sum_k A[m, k] * B[n, k] with
//   m = output pixel of one output "phase" grid, n = output channel, k = (tap, input channel).
//   conv:    one phase, taps = all (ky, kx), input pixel = (i * stride + ky - pad, j * stride + kx - pad)
//   deconv:  4 phases (oy%2, ox%2); phase (py, px) uses the taps with ky = (py + pad) mod 2 (9/6/6/4 taps),
//            input pixel = (i + (py + pad - ky) / 2, j + ...) -- a stride-1 gather, no zero insertion
//   GDN:     1x1 "conv" of x^2 with gamma, epilogue out = x * rsqrt(acc + beta)   (IGDN: * sqrt)
//
// Precision.  Operands are fp32 values SPLIT into two bf16 planes (hi = bf16(x), lo = bf16(x - hi)); each
// k16 step issues three tcgen05.mma (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), which keeps ~16
// mantissa bits per operand: results agree with an fp32 reference to ~1e-5 relative, which the symbol
// rounding downstream needs.  Activations travel between layers as the two bf16 planes (same bytes as
// fp32), so the A operand is a pure copy: cp.async (LDGSTS, 16 B = 8 channels of one pixel, zero-filled
// for padding) straight into the UMMA canonical K-major layout; weights are pre-packed on the host in
// that same layout and arrive with one TMA bulk copy per stage.
//
// CTA = 5 warps, two CTAs per SM: warps 0-1 produce the A tile (4 lanes per pixel), warps 0-3 run the epilogue
// (thread = TMEM lane = pixel row), warp 4 issues the MMAs (one elected lane) and owns the TMEM allocation.  smem
// ring of 2-3 stages of BK = 32 (A hi/lo 2 x 8.3 KB + B hi/lo BN*128 B per stage), mbarrier full/empty per stage
// (producers arrive asynchronously with cp.async.mbarrier.arrive.noinc), accumulators of 128 lanes x BN columns of
// TMEM (main + GDN norm), drained with double-buffered tcgen05.ld.
//
// k-step order.  Taps come in groups (same dy, dx congruent modulo the input stride) whose input pixel sets are
// one-pixel shifts of each other; the kernel walks  group -> channel chunk -> tap in group  and the A copies are
// L1-allocating (cp.async.ca), so only the first tap of a group fetches its activation lines from L2 and the others
// hit L1.  This halves the L2 -> SM traffic of the 5x5 layers and was worth 2.85 -> 2.08 ms on the largest launch
// (an ablation had shown the A-operand path, not the tensor pipe or the weights, to be the bound).
#include <cstdlib>


#include "conv_common.cuh"

namespace cai {

// Ablation / trace instrumentation exists only in -DCAI_DEBUG_BUILD builds (`make DEBUG=1`); release kernels carry
// none of it: CAI_DBG() folds to false and CAI_TRACE() to nothing at compile time.
#ifdef CAI_DEBUG_BUILD
__device__ long long g_conv_trace[64 * 8];  // phase timestamps of the first 64 CTAs (debug bit 32)
#define CAI_DBG(mask) ((p.debug & (mask)) != 0)
#else
#define CAI_DBG(mask) (false)
#endif

#ifdef CAI_DEBUG_BUILD
#define CAI_TRACE(slot) do { if (CAI_DBG(32) && blockIdx.x < 64 && blockIdx.y == 0 && tid == 0) g_conv_trace[blockIdx.x * 8 + (slot)] = clock64(); } while (0)
#else
#define CAI_TRACE(slot) do { } while (0)
#endif

// KIND selects a specialised epilogue so that each instantiation stays small (the all-runtime-flags version is 33k
// SASS instructions and thrashes the instruction cache in its epilogue loops):
//   0 generic (every flag read at run time)        1 fused GDN / IGDN, split-plane output only
//   2 linear / ReLU / LeakyReLU, plane output only  3 fp32 output (+ optional |.| planes), activation, clamp
template <int KIND>
__global__ void __launch_bounds__(kConvThreads, 2) conv_gemm_kernel(const __grid_constant__ ConvKernelParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[4];
  __shared__ __align__(8) uint64_t empty_bar[4];
  __shared__ __align__(8) uint64_t acc_bar;
  __shared__ __align__(8) uint64_t acc2_bar;
  __shared__ __align__(8) uint64_t gfull_bar[8];  // GDN k-steps: each used once, 128 writer arrivals + gamma TMA
  __shared__ uint32_t s_tmem_base;
  __shared__ int64_t s_opix[kBM];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  CAI_TRACE(0);
  const int stages = p.stages;
  const uint32_t a_plane = (kBK / 8) * kLboA;      // one bf16 plane of the A tile: kBK/8 chunks of kLboA bytes
  const uint32_t b_plane = static_cast<uint32_t>(BN) * kBK * 2;
  const uint32_t stage_bytes = 2 * a_plane + 2 * b_plane;
  const int ksteps = p.ntaps * p.kchunks;
  const int64_t M_total = static_cast<int64_t>(p.N) * p.Hp * p.Wp;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kBM;
  const int n_tile = blockIdx.y;
  const int n0 = n_tile * BN;

  const bool fuse_gdn = (KIND == 0) ? (p.gdn_w != nullptr) : (KIND == 1);
  const bool has_f32 = (KIND == 0) ? (p.out_f32 != nullptr) : (KIND == 3);
  const bool has_out = (KIND == 0) ? (p.out_hi != nullptr) : (KIND == 1 || KIND == 2);
  const bool has_sq = (KIND == 0) ? (p.sq_hi != nullptr) : false;
  const bool has_abs = (KIND == 0 || KIND == 3) ? (p.abs_hi != nullptr) : false;
  const int epi = (KIND == 1) ? 0 : p.epilogue;
  const bool has_aux = (KIND == 0) ? (epi >= 3) : false;
  const bool do_clamp = (KIND == 0 || KIND == 3) ? (p.clamp_lo < p.clamp_hi) : false;
  const int gdn_ksteps = fuse_gdn ? (BN + kBK - 1) / kBK : 0;
  uint32_t acc_cols = 32;
  while (acc_cols < static_cast<uint32_t>(BN)) acc_cols <<= 1;
  const uint32_t tmem_cols = fuse_gdn ? 2 * acc_cols : acc_cols;  // second accumulator at column acc_cols

  __shared__ __align__(16) float s_bias[256];
  __shared__ __align__(16) float s_beta[256];
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], kProducerThreads + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    mbar_init(&acc2_bar, 1);
    for (int g = 0; g < 8; ++g) mbar_init(&gfull_bar[g], 128 + 1);
    mbar_fence_init();
  }
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  CAI_TRACE(1);

  if (warp >= 2 && warp < kEpiWarps) {
    // bias / beta of this N tile -> shared memory (warps 2-7 idle through the main loop).  The epilogue reads them
    // per column group, and a global load there -- L1 is being streamed through by the operand copies -- put an L2
    // round trip on every group (measured: ~40% of the epilogue).  Published by the bar.sync after the acc wait.
    for (int i = tid - 64; i < BN; i += kEpiThreads - 64) {
      const bool in = n0 + i < p.Cout;
      s_bias[i] = (p.bias && in) ? __ldg(p.bias + n0 + i) : 0.f;
      s_beta[i] = (fuse_gdn && in) ? __ldg(p.gdn_beta + n0 + i) : 1.f;
    }
  }
  if (warp < kEpiWarps) {
    // ===================== producers =====================
    // Load mapping: 4 lanes share one pixel (its kBK = 32 channels = 64 contiguous bytes per plane), a warp
    // instruction covers 8 pixels -> 8 L1 wavefronts per LDGSTS instead of 32 with a lane-per-pixel mapping.
    // Thread (warp w, lane l) loads chunk c = l & 3 of rows it * 32 + w * 8 + (l >> 2), it = 0..3.
    // The epilogue below keeps thread = TMEM lane = row `tid & 127`; the two warps that share a lane quarter
    // (w and w + 4, group h = tid >> 7) take alternate 32-column slabs of the tile.
    const int r = tid & (kBM - 1);
    const int h = tid >> 7;
    const int64_t m = m0 + r;
    const bool row_ok = m < M_total;
    int n_img = 0, pi = 0, pj = 0;
    const uint32_t per_img = static_cast<uint32_t>(p.Hp) * static_cast<uint32_t>(p.Wp);  // host checks M_total < 2^31
    if (row_ok) {
      const uint32_t mu = static_cast<uint32_t>(m);
      n_img = static_cast<int>(mu / per_img);
      const uint32_t rem = mu - static_cast<uint32_t>(n_img) * per_img;
      pi = static_cast<int>(rem / static_cast<uint32_t>(p.Wp));
      pj = static_cast<int>(rem - static_cast<uint32_t>(pi) * static_cast<uint32_t>(p.Wp));
    }
    const uint32_t row_off = (static_cast<uint32_t>(r) >> 3) * 128u + (static_cast<uint32_t>(r) & 7u) * 16u;
    const unsigned char *wbase = p.w_packed + static_cast<size_t>(n_tile) * ksteps * (2 * b_plane);

    const int lc = lane & 3;  // k-chunk handled by this lane
    const bool is_loader = warp < kProducerThreads / 32;
    int ld_iy[kLoadIters], ld_ix[kLoadIters];  // input coordinate of tap (0, 0) for the rows this thread loads
    int64_t ld_img[kLoadIters];                // image base pixel index
    uint32_t ld_so[kLoadIters];
#pragma unroll
    for (int it = 0; it < kLoadIters; ++it) {
      const int lr = it * (kProducerThreads / 4) + (warp & (kProducerThreads / 32 - 1)) * 8 + (lane >> 2);
      const int64_t lm = m0 + lr;
      ld_so[it] = static_cast<uint32_t>(lc) * kLboA + (static_cast<uint32_t>(lr) >> 3) * 128u +
                  (static_cast<uint32_t>(lr) & 7u) * 16u;
      if (lm < M_total) {
        const uint32_t lmu = static_cast<uint32_t>(lm);
        const uint32_t ni = lmu / per_img;
        const uint32_t rem = lmu - ni * per_img;
        const uint32_t li = rem / static_cast<uint32_t>(p.Wp);
        ld_iy[it] = static_cast<int>(li) * p.is;
        ld_ix[it] = static_cast<int>(rem - li * static_cast<uint32_t>(p.Wp)) * p.is;
        ld_img[it] = static_cast<int64_t>(ni) * p.H * p.W;
      } else {
        ld_iy[it] = -(1 << 28);
        ld_ix[it] = 0;
        ld_img[it] = 0;
      }
    }

    // Per-GROUP state (recomputed only when the tap group changes): element offset of the input pixel for
    // (dy of the group, dx = 0) for each of this thread's rows, and 0 / 16 bytes for rows whose input row is padding.
    // k-steps walk  group -> channel chunk -> tap in group: the taps of a group read the same activation lines
    // shifted by one pixel, so all but the first are L1 hits instead of L2 round trips.  Per k-step only
    // dx * Cin + kc * kBK is added and the column bound is checked.
    int64_t tap_off[kLoadIters];
    uint32_t tap_bytes[kLoadIters];
    int cur_t0 = 0, cur_g = 0, cur_gl = p.glen[0], cur_ti = 0, cur_kc = 0;
    auto set_group = [&](int t0) {
      const int dy = p.dy[t0];
#pragma unroll
      for (int it = 0; it < kLoadIters; ++it) {
        const int iy = ld_iy[it] + dy;
        const bool ok = iy >= 0 && iy < p.H;
        tap_off[it] = ok ? ((ld_img[it] + static_cast<int64_t>(iy) * p.W + ld_ix[it]) * p.Cin + lc * 8) : 0;
        tap_bytes[it] = ok ? 16u : 0u;
      }
    };
    set_group(0);

    auto issue = [&](int ks, int s) {  // must be called with ks = 0, 1, 2, ... in order; s = ks mod stages
      unsigned char *sa = smem + static_cast<uint32_t>(s) * stage_bytes;
      const int kbase = cur_kc * kBK;
      const bool k_ok = kbase + lc * 8 < p.Cin;
      const int dx = p.dx[cur_t0 + cur_ti];
      const int64_t koff = static_cast<int64_t>(dx) * p.Cin + kbase;
#pragma unroll
      for (int it = 0; it < kLoadIters; ++it) {
        if CAI_DBG(1) break;
        const bool x_ok = static_cast<uint32_t>(ld_ix[it] + dx) < static_cast<uint32_t>(p.W);
        const uint32_t nbytes = (k_ok && x_ok) ? tap_bytes[it] : 0u;
        const int64_t off = nbytes ? tap_off[it] + koff : 0;
        cp_async16(sa + ld_so[it], p.a_hi + off, nbytes);
        cp_async16(sa + a_plane + ld_so[it], p.a_lo + off, nbytes);
      }
      if (tid == 0) {
        if CAI_DBG(2) {
          mbar_arrive(&full_bar[s]);
        } else {
          mbar_expect_tx(&full_bar[s], 2 * b_plane);
          tma_bulk_g2s(sa + 2 * a_plane, wbase + static_cast<size_t>(ks) * (2 * b_plane), 2 * b_plane, &full_bar[s]);
        }
      }
      if (++cur_ti == cur_gl) {
        cur_ti = 0;
        if (++cur_kc == p.kchunks) {
          cur_kc = 0;
          cur_t0 += cur_gl;
          ++cur_g;
          if (cur_t0 < p.ntaps) {
            cur_gl = p.glen[cur_g];
            set_group(cur_t0);
          }
        }
      }
    };

    // Producer loop.  Each thread's cp.async copies of a k-step are tied to the stage's "full" mbarrier with
    // cp.async.mbarrier.arrive.noinc: the hardware performs this thread's arrival when its copies have landed, so
    // the thread never waits for its own loads and runs up to `stages` k-steps ahead of the MMA warp (the only
    // blocking point is the "empty" barrier of the stage being refilled).  Ring positions and parities are carried
    // incrementally.
    int ps = 0, ppass = 0;  // stage of k-step ks, number of completed passes over the ring
    if (!is_loader) {       // warps 2-7 only take part in the epilogue: fast-forward their ring position
      ppass = ksteps / stages;
      ps = ksteps - ppass * stages;
    }
    for (int ks = 0; is_loader && ks < ksteps; ++ks) {
      if (ppass > 0) mbar_wait_bounded(&empty_bar[ps], (ppass - 1) & 1);
      issue(ks, ps);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[ps])) : "memory");
      if (++ps == stages) {
        ps = 0;
        ++ppass;
      }
    }

    // ===================== epilogue: thread = TMEM lane = pixel row =====================
    CAI_TRACE(2);
    // gamma prefetch: the ring stages the GDN GEMM will use free up in MMA order while the main loop drains, so
    // thread 0 claims the first min(stages, gdn_ksteps) of them as they free and starts their TMA loads now --
    // the copies land while the remaining threads are still waiting for the accumulator.
    const int gdn_pre = fuse_gdn ? (gdn_ksteps < stages ? gdn_ksteps : stages) : 0;
    if (tid == 0) {
      int s = ps, pass = ppass;
      for (int g = 0; g < gdn_pre; ++g) {
        if (pass > 0) mbar_wait_bounded(&empty_bar[s], (pass - 1) & 1);
        mbar_expect_tx(&gfull_bar[g], 2 * b_plane);
        tma_bulk_g2s(smem + static_cast<uint32_t>(s) * stage_bytes + 2 * a_plane,
                     p.gdn_w + static_cast<size_t>(g) * (2 * b_plane), 2 * b_plane, &gfull_bar[g]);
        if (++s == stages) {
          s = 0;
          ++pass;
        }
      }
    }
    mbar_wait_bounded(&acc_bar, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // s_bias / s_beta visible to all epilogue warps
    CAI_TRACE(3);
    const uint32_t lane_base = (static_cast<uint32_t>(warp & 3) * 32u) << 16;
    if (fuse_gdn) {
      // ---- fused GDN, part 1: turn the accumulator into the A operand of the second GEMM.
      // x = acc + bias; x^2 is split into bf16 planes and written, kBK channels (one k-step) at a time, into the
      // A area of the next ring stage; gamma's matching K chunk arrives in the B area by TMA.
      // The TMEM load of slab g+1 is in flight while slab g is processed (two register buffers).
      // Group h writes the k-steps g = h, h + 2, ...: the two groups fill two ring stages concurrently.
      uint32_t raw[32];
#pragma unroll 1
      for (int g = h; g < gdn_ksteps; g += 2) {
        int s = ps + g, pass = ppass;  // ring stage / pass of GDN k-step g (k-steps continue the main loop's ring walk)
        while (s >= stages) {
          s -= stages;
          ++pass;
        }
        if (pass > 0) mbar_wait_bounded(&empty_bar[s], (pass - 1) & 1);
        unsigned char *sa = smem + static_cast<uint32_t>(s) * stage_bytes;
        if (r == 0 && g >= gdn_pre) {
          mbar_expect_tx(&gfull_bar[g], 2 * b_plane);
          tma_bulk_g2s(sa + 2 * a_plane, p.gdn_w + static_cast<size_t>(g) * (2 * b_plane), 2 * b_plane, &gfull_bar[g]);
        }
        static_assert(kBK == 32, "the GDN operand phase loads one 32-column TMEM slab per k-step");
        const int col0 = g * kBK;
        const bool any = col0 < BN;  // warp-uniform (BN is a multiple of 16: a k-step may be half empty)
        if (any && !CAI_DBG(128)) tmem_ld32_nowait(tmem_base + lane_base + static_cast<uint32_t>(col0), raw);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < kBK / 8; ++c) {
          const int col = col0 + c * 8;
          uint4 vh = make_uint4(0u, 0u, 0u, 0u), vl = make_uint4(0u, 0u, 0u, 0u);
          if (any && col < BN) {
            float sq[8];
            const float4 b0 = *reinterpret_cast<const float4 *>(s_bias + col);
            const float4 b1 = *reinterpret_cast<const float4 *>(s_bias + col + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float a = __uint_as_float(raw[c * 8 + i]) + bb[i];
              sq[i] = a * a;
            }
            const Pack8 pk = split8(sq);
            vh = pk.hi;
            vl = pk.lo;
          }
          const uint32_t so = static_cast<uint32_t>(c) * kLboA + row_off;
          *reinterpret_cast<uint4 *>(sa + so) = vh;
          *reinterpret_cast<uint4 *>(sa + a_plane + so) = vl;
        }
        if (!CAI_DBG(64)) fence_async_proxy();  // generic-proxy stores -> visible to the tensor core (async proxy)
        mbar_arrive(&gfull_bar[g]);   // every writer arrives: no CTA-wide barrier needed
      }
      CAI_TRACE(4);
      mbar_wait_bounded(&acc2_bar, 0);
      tc_fence_after();
      CAI_TRACE(5);
    }
    // ---- output epilogue.  Each thread owns one pixel row of the accumulator; writing that row straight to
    // global memory makes every warp store hit 32 different rows with 16-byte pieces (measured: ~38k cycles per
    // tile).  Instead the tile is staged in the (now idle) operand ring with a padded row pitch and then copied
    // out cooperatively, consecutive lanes covering consecutive 16-byte units of a row (full-line stores).
    int64_t opix = -1;
    if (row_ok) {
      const int oy = pi * p.os + p.o0y, ox = pj * p.os + p.o0x;
      opix = (static_cast<int64_t>(n_img) * p.Ho + oy) * p.Wo + ox;
    }
    if (h == 0) s_opix[r] = opix;
    // staged buffers: fp32 tile and/or bf16 plane pairs (out, out^2, |out|)
    struct StageBuf {
      unsigned char *g;   // global base of the tensor (element (pixel 0, channel 0))
      uint32_t off;       // byte offset of the staging buffer inside the ring
      uint32_t esize;     // bytes per element
    };
    StageBuf bufs[7];
    int nbuf = 0;
    const int n_f32 = has_f32 ? 1 : 0;
    const int n_pl = (has_out ? 1 : 0) + (has_sq ? 1 : 0) + (has_abs ? 1 : 0);
    const uint32_t ring_bytes = static_cast<uint32_t>(stages) * stage_bytes;
    int ncols = BN;  // columns per pass: as many as fit the ring
    while (ncols > 16 && kBM * (n_f32 * (ncols * 4u + 16u) + n_pl * 2u * (ncols * 2u + 16u)) > ring_bytes) ncols -= 16;
    const uint32_t pitch_f = ncols * 4u + 16u, pitch_b = ncols * 2u + 16u;
    {
      uint32_t off = 0;
      if (has_f32) {
        bufs[nbuf++] = {reinterpret_cast<unsigned char *>(p.out_f32), off, 4u};
        off += kBM * pitch_f;
      }
      __nv_bfloat16 *pl[6] = {p.out_hi, p.out_lo, p.sq_hi, p.sq_lo, p.abs_hi, p.abs_lo};
      const bool use[6] = {has_out, has_out, has_sq, has_sq, has_abs, has_abs};
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        if (use[q]) {
          bufs[nbuf++] = {reinterpret_cast<unsigned char *>(pl[q]), off, 2u};
          off += kBM * pitch_b;
        }
      }
    }
    const uint32_t off_out = n_f32 * kBM * pitch_f;                   // first plane buffer
    const uint32_t off_sq = off_out + (has_out ? 2u : 0u) * kBM * pitch_b;
    const uint32_t off_abs = off_sq + (has_sq ? 2u : 0u) * kBM * pitch_b;

    for (int cA = 0; cA < BN; cA += ncols) {
      const int cB = (cA + ncols < BN) ? cA + ncols : BN;
      // ---- phase 1: TMEM -> registers -> epilogue math -> staging (thread = row)
      // The TMEM loads of column group c0+16 are in flight while group c0 is processed (two register buffers).
      auto issue = [&](int c0, uint32_t (&raw)[16], uint32_t (&raw2)[16]) {
        tmem_ld16_nowait(tmem_base + lane_base + static_cast<uint32_t>(c0), raw);
        if (fuse_gdn) tmem_ld16_nowait(tmem_base + lane_base + acc_cols + static_cast<uint32_t>(c0), raw2);
      };
      auto process = [&](int c0, const uint32_t (&raw)[16], const uint32_t (&raw2)[16]) {
        if (!row_ok || CAI_DBG(8)) return;
        const int cg = n0 + c0;  // global output channel of raw[0]
        if (cg >= p.Cout) return;
        float v[16];
        if (p.bias) {
          const float4 *b4 = reinterpret_cast<const float4 *>(s_bias + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bb = b4[q];
            v[4 * q] = __uint_as_float(raw[4 * q]) + bb.x;
            v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + bb.y;
            v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + bb.z;
            v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + bb.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
        }
        if (fuse_gdn) {
          const float4 *g4 = reinterpret_cast<const float4 *>(s_beta + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bb = g4[q];
            const float be[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float nrm = __uint_as_float(raw2[4 * q + i]) + be[i];
              const float rs = rsqrtf(nrm);
              v[4 * q + i] *= (p.gdn_mode == 1) ? rs : nrm * rs;  // n^-1/2 or n^+1/2 = n * n^-1/2
            }
          }
        }
        if (epi == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (epi == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
        } else if (has_aux) {
          const int64_t obase = opix * p.Cout + cg;
          const uint4 *ah = reinterpret_cast<const uint4 *>(p.aux_hi + obase);
          const uint4 *al = reinterpret_cast<const uint4 *>(p.aux_lo + obase);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 qh = __ldg(ah + h), ql = __ldg(al + h);
            const __nv_bfloat16 *bh = reinterpret_cast<const __nv_bfloat16 *>(&qh);
            const __nv_bfloat16 *bl = reinterpret_cast<const __nv_bfloat16 *>(&ql);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x = __bfloat162float(bh[i]) + __bfloat162float(bl[i]);
              const float nrm = v[h * 8 + i];
              const float rs = rsqrtf(nrm);
              v[h * 8 + i] = x * ((epi == 3) ? rs : nrm * rs);
            }
          }
        }
        if (do_clamp) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(v[i], p.clamp_lo), p.clamp_hi);
        }
        const uint32_t cc = static_cast<uint32_t>(c0 - cA);
        if (has_f32) {
          float4 *o = reinterpret_cast<float4 *>(smem + static_cast<uint32_t>(r) * pitch_f + cc * 4u);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        const uint32_t rb = static_cast<uint32_t>(r) * pitch_b + cc * 2u;
        if (has_out) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const Pack8 pk = split8(v + 8 * h);
            *reinterpret_cast<uint4 *>(smem + off_out + rb + h * 16u) = pk.hi;
            *reinterpret_cast<uint4 *>(smem + off_out + kBM * pitch_b + rb + h * 16u) = pk.lo;
          }
        }
        if (has_sq) {
          float sv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) sv[i] = v[i] * v[i];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const Pack8 pk = split8(sv + 8 * h);
            *reinterpret_cast<uint4 *>(smem + off_sq + rb + h * 16u) = pk.hi;
            *reinterpret_cast<uint4 *>(smem + off_sq + kBM * pitch_b + rb + h * 16u) = pk.lo;
          }
        }
        if (has_abs) {
          float sv[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) sv[i] = fabsf(v[i]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const Pack8 pk = split8(sv + 8 * h);
            *reinterpret_cast<uint4 *>(smem + off_abs + rb + h * 16u) = pk.hi;
            *reinterpret_cast<uint4 *>(smem + off_abs + kBM * pitch_b + rb + h * 16u) = pk.lo;
          }
        }
      };
      {
        // group h takes the 32-column slabs h, h + 2, ... of this pass, 16 columns per TMEM load (the other warps
        // resident on the scheduler -- four epilogue warps per scheduler with two CTAs per SM -- hide the load)
        uint32_t qa[16], qa2[16];
#pragma unroll 1
        for (int s0 = cA + 32 * h; s0 < cB; s0 += 64) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c0 = s0 + 16 * q;
            if (c0 < cB) {
              issue(c0, qa, qa2);
              tmem_wait_ld();
              process(c0, qa, qa2);
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (cA == 0) CAI_TRACE(6);
      // ---- phase 2: cooperative copy-out, 16-byte units, consecutive lanes along a row
      const int cols_here = (cB - cA < p.Cout - (n0 + cA)) ? (cB - cA) : (p.Cout - (n0 + cA));
      if (cols_here > 0 && !CAI_DBG(8)) {
#pragma unroll 1
        for (int bi = 0; bi < nbuf; ++bi) {
          const StageBuf sb = bufs[bi];
          const uint32_t pitch = (sb.esize == 4u) ? pitch_f : pitch_b;
          const uint32_t units = static_cast<uint32_t>(cols_here) * sb.esize / 16u;  // per row
          unsigned char *gbase = sb.g + static_cast<int64_t>(n0 + cA) * sb.esize;
          const int64_t row_stride = static_cast<int64_t>(p.Cout) * sb.esize;
          if ((units & (units - 1u)) == 0u && units <= static_cast<uint32_t>(kEpiThreads)) {
            // power-of-two units per row (the common case): shift / mask indexing, rows_per_iter rows per sweep
            const uint32_t j = tid & (units - 1u);
            const uint32_t rows_per_iter = static_cast<uint32_t>(kEpiThreads) / units;
            uint32_t row = tid / units;  // tid >> log2(units); done once
            const unsigned char *sp = smem + sb.off + row * pitch + j * 16u;
            const uint32_t sp_step = rows_per_iter * pitch;
#pragma unroll 4
            for (; row < kBM; row += rows_per_iter, sp += sp_step) {
              const int64_t op = s_opix[row];
              if (op >= 0) *reinterpret_cast<uint4 *>(gbase + op * row_stride + j * 16u) = *reinterpret_cast<const uint4 *>(sp);
            }
          } else {
            const uint32_t total_u = kBM * units;
            for (uint32_t u = tid; u < total_u; u += static_cast<uint32_t>(kEpiThreads)) {
              const uint32_t row = u / units, j = u - row * units;
              const int64_t op = s_opix[row];
              if (op < 0) continue;
              const uint4 val = *reinterpret_cast<const uint4 *>(smem + sb.off + row * pitch + j * 16u);
              *reinterpret_cast<uint4 *>(gbase + op * row_stride + j * 16u) = val;
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
  } else {
    // ===================== MMA issuer (warp 8, one elected lane) =====================
    // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                           (static_cast<uint32_t>(kBM >> 4) << 24);
    const uint32_t lbo_a = kLboA, lbo_b = static_cast<uint32_t>(BN) * 16u;
    int s = 0;
    uint32_t mphase = 0;
    for (int ks = 0; ks < ksteps + gdn_ksteps; ++ks) {
      if (ks < ksteps) mbar_wait_bounded(&full_bar[s], mphase);
      else mbar_wait_bounded(&gfull_bar[ks - ksteps], 0);
      tc_fence_after();
      if (lane == 0) {
        const bool second = ks >= ksteps;  // GDN GEMM: A = x^2 planes written by the epilogue warps, B = gamma
        const uint32_t d_tmem = second ? tmem_base + acc_cols : tmem_base;
        const int first_ks = second ? ksteps : 0;
        const uint32_t sa = smem_u32(smem) + static_cast<uint32_t>(s) * stage_bytes;
        const uint32_t a_hi = sa, a_lo = sa + a_plane, b_hi = sa + 2 * a_plane, b_lo = b_hi + b_plane;
#pragma unroll
        for (int kk = 0; kk < kBK / 16; ++kk) {
          if CAI_DBG(4) break;
          const uint64_t dah = make_smem_desc(a_hi + kk * 2 * lbo_a, lbo_a, 128);
          const uint64_t dal = make_smem_desc(a_lo + kk * 2 * lbo_a, lbo_a, 128);
          const uint64_t dbh = make_smem_desc(b_hi + kk * 2 * lbo_b, lbo_b, 128);
          const uint64_t dbl = make_smem_desc(b_lo + kk * 2 * lbo_b, lbo_b, 128);
          umma_bf16(d_tmem, dah, dbh, idesc, (ks > first_ks || kk > 0) ? 1u : 0u);
          umma_bf16(d_tmem, dah, dbl, idesc, 1u);
          umma_bf16(d_tmem, dal, dbh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);                         // frees the smem stage when the MMAs have read it
        if (ks == ksteps - 1) umma_commit(&acc_bar);        // main accumulator complete
        if (second && ks == ksteps + gdn_ksteps - 1) umma_commit(&acc2_bar);  // norm accumulator complete
      }
      __syncwarp();
      if (++s == stages) {
        s = 0;
        mphase ^= 1u;
      }
    }
  }

  CAI_TRACE(7);
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- helper kernels ------------------------------------------------------------------------------------

// fp32 (NCHW or NHWC) -> split bf16 NHWC planes, optionally padding channels to Cpad with zeros
__global__ void __launch_bounds__(256)
split_planes_kernel(const float *__restrict__ x, int layout, int64_t N, int64_t C, int64_t HW, int64_t Cpad,
                    __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
  const int64_t total = N * HW * Cpad;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t c = i % Cpad, pix = i / Cpad;
    float v = 0.f;
    if (c < C) {
      const int64_t n = pix / HW, hw = pix - n * HW;
      v = (layout == CAI_LAYOUT_NHWC) ? x[pix * C + c] : x[(n * C + c) * HW + hw];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// im2col for tiny Cin (first layer, Cin = 3): [N, H, W, C] fp32 (any layout) -> split planes [N*Ho*Wo, Kpad]
// with k = (ky * ks + kx) * C + c, zero padded to Kpad.
// One thread = one output pixel x one group of 8 k-values -> one 16-byte store per plane (coalesced along k).
// CT / KT > 0 fix the channel count and kernel size at compile time (the codecs' first layer is C = 3, k = 5:
// every division below becomes a multiply); 0 = runtime values.
template <int CT, int KT>
__global__ void __launch_bounds__(256)
im2col_split_kernel(const float *__restrict__ x, int layout, int N, int C_, int H, int W, int Ho, int Wo, int ksz_,
                    int stride, int pad, int Kpad, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
  const int C = CT > 0 ? CT : C_;
  const int ksz = KT > 0 ? KT : ksz_;
  const int groups = Kpad >> 3;
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * groups;
  const int64_t gs = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int K = ksz * ksz * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += gs) {
    const int g = static_cast<int>(i % groups);
    const int64_t pix = i / groups;
    const int ox = static_cast<int>(pix % Wo);
    const int64_t r = pix / Wo;
    const int oy = static_cast<int>(r % Ho), n = static_cast<int>(r / Ho);
    const int iy0 = oy * stride - pad, ix0 = ox * stride - pad;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < K) {
        const int t = k / C, c = k - t * C;
        const int ky = t / ksz, kx = t - ky * ksz;
        const int iy = iy0 + ky, ix = ix0 + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          val = (layout == CAI_LAYOUT_NHWC) ? __ldg(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * C + c)
                                            : __ldg(x + ((static_cast<int64_t>(n) * C + c) * H + iy) * W + ix);
        }
      }
      v[j] = val;
    }
    const Pack8 pk = split8(v);
    *reinterpret_cast<uint4 *>(hi + pix * Kpad + g * 8) = pk.hi;
    *reinterpret_cast<uint4 *>(lo + pix * Kpad + g * 8) = pk.lo;
  }
}

// Tiled im2col for the codecs' first layer (C = 3, k = 5, s = 2, p = 2).  The block stages the fp32 input patch of a
// 4 x 64 output tile (11 x 131 x 3 floats, read once with coalesced row segments) in shared memory and builds the
// split-plane rows from there, so the 75 taps of a pixel cost shared-memory reads instead of 75 scattered global
// loads; stores stay one 16-byte unit per (pixel, k-group), contiguous across the tile row.
constexpr int kI2cTY = 4, kI2cTX = 64;
constexpr int kI2cPR = 2 * kI2cTY + 3, kI2cPC = 2 * kI2cTX + 3, kI2cPitch = kI2cPC + 1;
__global__ void __launch_bounds__(256)
im2col_k5s2_c3_kernel(const float *__restrict__ x, int layout, int H, int W, int Ho, int Wo, int Kpad,
                      __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
  __shared__ float s_patch[3][kI2cPR][kI2cPitch];
  __shared__ int s_koff[80];  // k -> offset of tap (ky, kx), channel c inside the patch; -1 for the K padding
  if (threadIdx.x < 80) {
    const int k = threadIdx.x, t = k / 3, c = k - 3 * t, ky = t / 5, kx = t - 5 * ky;
    s_koff[k] = k < 75 ? (c * kI2cPR + ky) * kI2cPitch + kx : -1;
  }
  const int n = blockIdx.z, oy0 = blockIdx.y * kI2cTY, ox0 = blockIdx.x * kI2cTX;
  const int iy0 = 2 * oy0 - 2, ix0 = 2 * ox0 - 2;
  constexpr int kPatch = 3 * kI2cPR * kI2cPC;
  for (int u0 = threadIdx.x; u0 < kPatch; u0 += 6 * 256) {  // six loads in flight per thread
    float v[6];
    int cc[6], rr[6], col[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const int u = u0 + e * 256;
      v[e] = 0.f;
      cc[e] = -1;
      if (u < kPatch) {
        const int c = u / (kI2cPR * kI2cPC), rem = u - c * (kI2cPR * kI2cPC);
        const int r = rem / kI2cPC, cl = rem - r * kI2cPC;
        const int iy = iy0 + r, ix = ix0 + cl;
        cc[e] = c, rr[e] = r, col[e] = cl;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          v[e] = (layout == CAI_LAYOUT_NHWC) ? __ldg(x + ((static_cast<int64_t>(n) * H + iy) * W + ix) * 3 + c)
                                             : __ldg(x + ((static_cast<int64_t>(n) * 3 + c) * H + iy) * W + ix);
      }
    }
#pragma unroll
    for (int e = 0; e < 6; ++e)
      if (cc[e] >= 0) s_patch[cc[e]][rr[e]][col[e]] = v[e];
  }
  __syncthreads();
  const int groups = Kpad >> 3;
  for (int i = threadIdx.x; i < kI2cTY * kI2cTX * groups; i += 256) {
    const int pix = i / groups, g = i - pix * groups;
    const int py = pix / kI2cTX, px = pix - py * kI2cTX;
    const int oy = oy0 + py, ox = ox0 + px;
    if (oy >= Ho || ox >= Wo) continue;
    float v[8];
    const float *base = &s_patch[0][0][0] + (2 * py) * kI2cPitch + 2 * px;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int off = (g * 8 + j < 80) ? s_koff[g * 8 + j] : -1;
      v[j] = off >= 0 ? base[off] : 0.f;
    }
    const Pack8 pk = split8(v);
    const int64_t o = ((static_cast<int64_t>(n) * Ho + oy) * Wo + ox) * Kpad + g * 8;
    *reinterpret_cast<uint4 *>(hi + o) = pk.hi;
    *reinterpret_cast<uint4 *>(lo + o) = pk.lo;
  }
}

// Tiled col2im for the codecs' last layer (Cout = 3, k = 5, s = 2, p = 2, output_padding = 1 -> Ho = 2H, Wo = 2W).
// The generic gather below reads each 320-byte cols row from ~25 different threads spread over the grid (measured:
// 3.5x the algorithmic DRAM traffic).  Here a block stages the cols rows of the (4+2) x (32+2) input pixels that
// feed an 8 x 64 output tile (float4 loads, whole rows) into shared memory with an odd pixel pitch, and each thread
// produces two horizontally adjacent output pixels (even ox: kx = 0, 2, 4; odd ox: kx = 1, 3) for all three channels.
constexpr int kC2iTY = 8, kC2iTX = 64;
constexpr int kC2iIY = kC2iTY / 2 + 2, kC2iIX = kC2iTX / 2 + 2;
constexpr int kC2iPitch = 77;  // words per staged pixel: 75 used, odd -> lanes on consecutive pixels hit distinct banks
constexpr int kC2iSmem = kC2iIY * kC2iIX * kC2iPitch * 4;
__global__ void __launch_bounds__(256)
col2im_k5s2_c3_kernel(const float *__restrict__ cols, const float *__restrict__ bias, int H, int W, int Ho, int Wo,
                      int Npad, int out_layout, float clamp_lo, float clamp_hi, float *__restrict__ out) {
  extern __shared__ float s_cols[];
  const int n = blockIdx.z, oy0 = blockIdx.y * kC2iTY, ox0 = blockIdx.x * kC2iTX;
  const int iy0 = oy0 / 2 - 1, ix0 = ox0 / 2 - 1;
  // staging: kC2iIY * kC2iIX * 19 float4 units, four loads in flight per thread before the first store (the loop as a
  // load -> store chain left one 16-byte request per thread outstanding: 2.5 TB/s)
  constexpr int kUnits = kC2iIY * kC2iIX * 19;
  for (int u0 = threadIdx.x; u0 < kUnits; u0 += 4 * 256) {
    float4 v[4];
    int dsto[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int u = u0 + e * 256;
      v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      dsto[e] = -1;
      if (u < kUnits) {
        const int pix = u / 19, q = u - pix * 19;
        const int ly = pix / kC2iIX, lx = pix - ly * kC2iIX;
        const int iy = iy0 + ly, ix = ix0 + lx;
        dsto[e] = pix * kC2iPitch + q * 4;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          v[e] = __ldcs(reinterpret_cast<const float4 *>(cols + ((static_cast<int64_t>(n) * H + iy) * W + ix) * Npad) + q);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (dsto[e] >= 0) {
        float *d = s_cols + dsto[e];
        d[0] = v[e].x; d[1] = v[e].y; d[2] = v[e].z; d[3] = v[e].w;
      }
    }
  }
  __syncthreads();
  const int ty = threadIdx.x >> 5, jx = threadIdx.x & 31;
  const int oy = oy0 + ty, ox = ox0 + 2 * jx;
  if (oy >= Ho || ox >= Wo) return;
  float acc[2][3];
#pragma unroll
  for (int co = 0; co < 3; ++co) acc[0][co] = acc[1][co] = bias ? __ldg(bias + co) : 0.f;
  for (int ky = ty & 1; ky < 5; ky += 2) {            // (oy + 2) % 2 == ty % 2 because oy0 is even
    const int ly = (ty + 2 - ky) / 2 + 1;
    const float *row = s_cols + (ly * kC2iIX + jx) * kC2iPitch + ky * 15;
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      acc[0][co] += row[2 * kC2iPitch + co];          // kx = 0 -> input pixel jx + 2
      acc[0][co] += row[1 * kC2iPitch + 6 + co];      // kx = 2 -> jx + 1
      acc[0][co] += row[12 + co];                     // kx = 4 -> jx
      acc[1][co] += row[2 * kC2iPitch + 3 + co];      // kx = 1 -> jx + 2
      acc[1][co] += row[1 * kC2iPitch + 9 + co];      // kx = 3 -> jx + 1
    }
  }
  if (clamp_lo < clamp_hi) {
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int co = 0; co < 3; ++co) acc[e][co] = fminf(fmaxf(acc[e][co], clamp_lo), clamp_hi);
  }
  const bool two = ox + 1 < Wo;
  if (out_layout == CAI_LAYOUT_NHWC) {
    float *o = out + ((static_cast<int64_t>(n) * Ho + oy) * Wo + ox) * 3;
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      o[co] = acc[0][co];
      if (two) o[3 + co] = acc[1][co];
    }
  } else {
#pragma unroll
    for (int co = 0; co < 3; ++co) {
      float *o = out + ((static_cast<int64_t>(n) * 3 + co) * Ho + oy) * Wo + ox;
      if (two && (Wo & 1) == 0) {
        *reinterpret_cast<float2 *>(o) = make_float2(acc[0][co], acc[1][co]);
      } else {
        o[0] = acc[0][co];
        if (two) o[1] = acc[1][co];
      }
    }
  }
}

// col2im gather for tiny Cout (last deconv, Cout = 3): cols fp32 [N*H*W, Npad] with n = (ky*ks+kx)*Cout + co
// -> out fp32 NCHW or NHWC [N, Cout, Ho, Wo], out(oy, ox) = bias + sum over taps with (oy + pad - ky) % stride == 0
__global__ void __launch_bounds__(256)
col2im_kernel(const float *__restrict__ cols, const float *__restrict__ bias, int N, int Cout, int H, int W, int Ho,
              int Wo, int ksz, int stride, int pad, int Npad, int out_layout, float clamp_lo, float clamp_hi,
              float *__restrict__ out) {
  const int64_t total = static_cast<int64_t>(N) * Ho * Wo * Cout;
  const int64_t gs = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += gs) {
    int co, ox, oy, n;
    if (out_layout == CAI_LAYOUT_NHWC) {
      co = static_cast<int>(i % Cout);
      int64_t r = i / Cout;
      ox = static_cast<int>(r % Wo);
      r /= Wo;
      oy = static_cast<int>(r % Ho);
      n = static_cast<int>(r / Ho);
    } else {
      ox = static_cast<int>(i % Wo);
      int64_t r = i / Wo;
      oy = static_cast<int>(r % Ho);
      r /= Ho;
      co = static_cast<int>(r % Cout);
      n = static_cast<int>(r / Cout);
    }
    float acc = bias ? bias[co] : 0.f;
    for (int ky = (oy + pad) % stride; ky < ksz; ky += stride) {
      const int iy = (oy + pad - ky) / stride;
      if (iy < 0 || iy >= H) continue;
      for (int kx = (ox + pad) % stride; kx < ksz; kx += stride) {
        const int ix = (ox + pad - kx) / stride;
        if (ix < 0 || ix >= W) continue;
        acc += cols[((static_cast<int64_t>(n) * H + iy) * W + ix) * Npad + (ky * ksz + kx) * Cout + co];
      }
    }
    if (clamp_lo < clamp_hi) acc = fminf(fmaxf(acc, clamp_lo), clamp_hi);
    out[i] = acc;
  }
}

static int ew_grid2(const DeviceProps &dp, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(dp.sm_count) * 32;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace cai

using namespace cai;

extern "C" {

int cai_conv_gemm(const cai_conv_desc *d, cai_stream_t stream_) {
  CAI_CHECK_ARG(d != nullptr, "cai_conv_gemm: NULL descriptor");
  CAI_CHECK_ARG(d->a_hi && d->a_lo && d->w_packed, "cai_conv_gemm: NULL operand");
  CAI_CHECK_ARG(d->N >= 1 && d->H >= 1 && d->W >= 1 && d->Hp >= 1 && d->Wp >= 1, "cai_conv_gemm: bad geometry");
  CAI_CHECK_ARG(d->Cin >= 8 && d->Cin % 8 == 0, "cai_conv_gemm: Cin=%d must be a multiple of 8", d->Cin);
  CAI_CHECK_ARG(d->BN >= 16 && d->BN <= 256 && d->BN % 16 == 0, "cai_conv_gemm: BN=%d must be a multiple of 16 <= 256",
                d->BN);
  CAI_CHECK_ARG(d->Cout >= 16 && d->Cout % 16 == 0, "cai_conv_gemm: Cout=%d must be a multiple of 16", d->Cout);
  CAI_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= kMaxTaps, "cai_conv_gemm: ntaps=%d out of range", d->ntaps);
  CAI_CHECK_ARG(d->epilogue >= 0 && d->epilogue <= 4, "cai_conv_gemm: bad epilogue");
  CAI_CHECK_ARG(d->epilogue < 3 || (d->aux_hi && d->aux_lo), "cai_conv_gemm: GDN epilogue needs aux planes");
  CAI_CHECK_ARG(d->out_f32 || d->out_hi, "cai_conv_gemm: no output");
  CAI_CHECK_ARG((d->out_hi == nullptr) == (d->out_lo == nullptr) && (d->sq_hi == nullptr) == (d->sq_lo == nullptr) &&
                    (d->abs_hi == nullptr) == (d->abs_lo == nullptr),
                "cai_conv_gemm: split planes come in pairs");
  CAI_CHECK_ARG(d->mode == 0 || d->mode == 1, "cai_conv_gemm: mode=%d", d->mode);
  if (d->mode == 1) {  // persistent TMA-fed kernel (weights packed chunk-major by the caller)
    CAI_CHECK_ARG(!d->aux_hi && !d->sq_hi, "cai_conv_gemm: mode 1 has no aux / squared-plane outputs");
    const int rc1 = launch_conv_tma(d, static_cast<cudaStream_t>(stream_));
    CAI_CHECK_ARG(rc1 != 1, "cai_conv_gemm: mode 1 requested for a layer cai_conv_tma_eligible() rejects");
    return rc1;
  }
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  ConvKernelParams p{};
  p.a_hi = static_cast<const __nv_bfloat16 *>(d->a_hi);
  p.a_lo = static_cast<const __nv_bfloat16 *>(d->a_lo);
  p.w_packed = static_cast<const unsigned char *>(d->w_packed);
  p.bias = d->bias;
  p.aux_hi = static_cast<const __nv_bfloat16 *>(d->aux_hi);
  p.aux_lo = static_cast<const __nv_bfloat16 *>(d->aux_lo);
  p.out_f32 = d->out_f32;
  p.out_hi = static_cast<__nv_bfloat16 *>(d->out_hi);
  p.out_lo = static_cast<__nv_bfloat16 *>(d->out_lo);
  p.sq_hi = static_cast<__nv_bfloat16 *>(d->sq_hi);
  p.sq_lo = static_cast<__nv_bfloat16 *>(d->sq_lo);
  p.abs_hi = static_cast<__nv_bfloat16 *>(d->abs_hi);
  p.abs_lo = static_cast<__nv_bfloat16 *>(d->abs_lo);
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Ho = d->Ho; p.Wo = d->Wo; p.Cout = d->Cout;
  p.Hp = d->Hp; p.Wp = d->Wp; p.os = d->os; p.o0y = d->o0y; p.o0x = d->o0x; p.is = d->is;
  p.ntaps = d->ntaps;
  p.kchunks = (d->Cin + kBK - 1) / kBK;
  p.BN = d->BN;
  p.epilogue = d->epilogue;
  p.clamp_lo = d->clamp_lo;
  p.clamp_hi = d->clamp_hi;
  for (int t = 0; t < d->ntaps; ++t) {
    p.dy[t] = d->dy[t];
    p.dx[t] = d->dx[t];
  }
  {  // tap groups: all zero = one tap per group; otherwise lengths must cover ntaps and a group shares dy
    int covered = 0;
    for (int g = 0; g < kMaxTaps; ++g) p.glen[g] = 0;
    if (d->glen[0] == 0) {
      for (int t = 0; t < d->ntaps; ++t) p.glen[t] = 1;
      covered = d->ntaps;
    } else {
      for (int g = 0; g < kMaxTaps && covered < d->ntaps; ++g) {
        const int n = d->glen[g];
        CAI_CHECK_ARG(n >= 1 && covered + n <= d->ntaps, "cai_conv_gemm: bad tap group length glen[%d]=%d", g, n);
        for (int t = covered + 1; t < covered + n; ++t)
          CAI_CHECK_ARG(d->dy[t] == d->dy[covered], "cai_conv_gemm: taps of group %d do not share dy", g);
        p.glen[g] = static_cast<int8_t>(n);
        covered += n;
      }
    }
    CAI_CHECK_ARG(covered == d->ntaps, "cai_conv_gemm: tap groups cover %d of %d taps", covered, d->ntaps);
  }
  const size_t stage_bytes = 2 * static_cast<size_t>((kBK / 8) * kLboA) + 2 * static_cast<size_t>(d->BN) * kBK * 2;
  // two CTAs per SM (one CTA's epilogue overlaps the other's main loop): each gets half of the shared memory
  int stages = static_cast<int>((static_cast<size_t>(dp.max_smem_optin) / 2 - 2048) / stage_bytes);
  if (stages > 3) stages = 3;  // measured on the short-K last-layer GEMM: 0.89 ms with 3 stages, 1.01 with 4
  // Multi-tap layers re-read each activation line from L1 for the later taps of a group: a shallower ring leaves
  // more of the unified L1/shared memory to the cache, which is worth more than the third stage (measured on the
  // 128->128 k5 s2 layer: 2.19 ms with 3 stages, 2.08 ms with 2).
  // (Round 2: a per-layer rule that kept the third stage on stride-1 inputs and small grids made the layers 1.5 %
  // faster alone -- 9.10 -> 8.97 ms per micro-batch -- and the overlapped step 3 % SLOWER, 104.4 vs 100.3-101.8 ms: the
  // bigger CTAs co-reside worse with the kernels of the other streams.  Two stages stay.)
  if (d->ntaps > 1 && stages > 2) stages = 2;
  if (stages < 2) stages = static_cast<int>((static_cast<size_t>(dp.max_smem_optin) - 2048) / stage_bytes) >= 2 ? 2 : stages;
  const Knobs &kn = knobs();
  if (kn.conv_stages >= 2) {  // tuning override (experiments only)
    const int cap = static_cast<int>((static_cast<size_t>(dp.max_smem_optin) - 2048) / stage_bytes);
    stages = kn.conv_stages < cap ? kn.conv_stages : cap;
    if (stages > 4) stages = 4;
  }
  const int ksteps = p.ntaps * p.kchunks;
  if (stages > ksteps) stages = ksteps < 2 ? 2 : ksteps;
  CAI_CHECK_ARG(stages >= 2, "cai_conv_gemm: tile does not fit shared memory");
  p.stages = stages;
  p.gdn_w = static_cast<const unsigned char *>(d->gdn_w);
  p.gdn_beta = d->gdn_beta;
  p.gdn_mode = d->gdn_mode;
#ifdef CAI_DEBUG_BUILD
  p.debug = kn.conv_debug;
#endif
  if (p.gdn_w) {
    CAI_CHECK_ARG(d->BN == d->Cout && d->BN <= 256, "cai_conv_gemm: fused GDN needs all channels in one tile (Cout <= 256)");
    CAI_CHECK_ARG(d->gdn_beta && (d->gdn_mode == 1 || d->gdn_mode == 2), "cai_conv_gemm: fused GDN needs beta and mode 1|2");
    CAI_CHECK_ARG(d->epilogue == 0, "cai_conv_gemm: fused GDN excludes another epilogue");
  }
  const size_t smem = stages * stage_bytes;
  const int64_t M_total = static_cast<int64_t>(d->N) * d->Hp * d->Wp;
  const int64_t mt = (M_total + kBM - 1) / kBM;
  const int nt = (d->Cout + d->BN - 1) / d->BN;
  CAI_CHECK_ARG(mt <= 0x7fffffff && nt <= 65535, "cai_conv_gemm: grid too large");
  CAI_CHECK_ARG(M_total < (1ll << 31), "cai_conv_gemm: more than 2^31 output pixels per launch");
  // pick the specialised epilogue
  int kind = 0;
  const bool only_planes = d->out_hi && !d->out_f32 && !d->sq_hi && !d->abs_hi;
  const bool no_clamp = !(d->clamp_lo < d->clamp_hi);
  if (p.gdn_w && only_planes && no_clamp) kind = 1;
  else if (!p.gdn_w && d->epilogue <= 2 && only_planes && no_clamp) kind = 2;
  else if (!p.gdn_w && d->epilogue <= 2 && d->out_f32 && !d->out_hi && !d->sq_hi) kind = 3;
  if (kn.conv_generic) kind = 0;
  const dim3 grid(static_cast<unsigned>(mt), static_cast<unsigned>(nt));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  // the kernel's dynamic shared memory limit (and the optional carve-out knob) is set once per device, never per launch
#define CAI_LAUNCH_KIND(K)                                                                                       \
  do {                                                                                                           \
    rc = optin_max_smem(reinterpret_cast<const void *>(conv_gemm_kernel<K>), dp, nullptr, kn.conv_carveout);     \
    if (rc != CAI_OK) return rc;                                                                                 \
    conv_gemm_kernel<K><<<grid, kConvThreads, smem, st>>>(p);                                                   \
  } while (0)
  switch (kind) {
    case 1: CAI_LAUNCH_KIND(1); break;
    case 2: CAI_LAUNCH_KIND(2); break;
    case 3: CAI_LAUNCH_KIND(3); break;
    default: CAI_LAUNCH_KIND(0); break;
  }
#undef CAI_LAUNCH_KIND
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

#ifdef CAI_DEBUG_BUILD
// experiment helper of debug builds only (tools/conv_probe.py); not part of the ABI in include/cai_b200.h
__attribute__((visibility("default"))) int cai_debug_conv_trace(long long *out_host) {
  CAI_CUDA(cudaMemcpyFromSymbol(out_host, g_conv_trace, sizeof(long long) * 64 * 8));
  return CAI_OK;
}
#endif

int cai_conv_tma_eligible(const cai_conv_desc *d) {
  if (!d || !d->a_hi || !d->a_lo) return 0;
  return conv_tma_eligible(d) ? 1 : 0;
}

int cai_split_planes(const float *x, int32_t layout, int64_t N, int64_t C, int64_t HW, int64_t Cpad, void *hi, void *lo,
                     cai_stream_t stream_) {
  CAI_CHECK_ARG(x && hi && lo && Cpad >= C && N >= 0 && C >= 1 && HW >= 0, "cai_split_planes: bad arguments");
  if (N * HW == 0) return CAI_OK;
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  split_planes_kernel<<<ew_grid2(dp, N * HW * Cpad), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      x, layout, N, C, HW, Cpad, static_cast<__nv_bfloat16 *>(hi), static_cast<__nv_bfloat16 *>(lo));
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_im2col_split(const float *x, int32_t layout, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                     int32_t ksize, int32_t stride, int32_t pad, int32_t Kpad, void *hi, void *lo, cai_stream_t stream_) {
  CAI_CHECK_ARG(x && hi && lo && Kpad >= ksize * ksize * C && Kpad % 8 == 0, "cai_im2col_split: bad arguments");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  const int grid = ew_grid2(dp, static_cast<int64_t>(N) * Ho * Wo * (Kpad / 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  __nv_bfloat16 *h = static_cast<__nv_bfloat16 *>(hi), *l = static_cast<__nv_bfloat16 *>(lo);
  if (C == 3 && ksize == 5 && stride == 2 && pad == 2 && Kpad >= 75 && N <= 65535 && !knobs().patch_generic) {
    dim3 tg((Wo + kI2cTX - 1) / kI2cTX, (Ho + kI2cTY - 1) / kI2cTY, N);
    im2col_k5s2_c3_kernel<<<tg, 256, 0, st>>>(x, layout, H, W, Ho, Wo, Kpad, h, l);
  } else if (C == 3 && ksize == 5)
    im2col_split_kernel<3, 5><<<grid, 256, 0, st>>>(x, layout, N, C, H, W, Ho, Wo, ksize, stride, pad, Kpad, h, l);
  else
    im2col_split_kernel<0, 0><<<grid, 256, 0, st>>>(x, layout, N, C, H, W, Ho, Wo, ksize, stride, pad, Kpad, h, l);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_col2im(const float *cols, const float *bias, int32_t N, int32_t Cout, int32_t H, int32_t W, int32_t Ho,
               int32_t Wo, int32_t ksize, int32_t stride, int32_t pad, int32_t Npad, int32_t out_layout, float clamp_lo,
               float clamp_hi, float *out, cai_stream_t stream_) {
  CAI_CHECK_ARG(cols && out && Npad >= ksize * ksize * Cout, "cai_col2im: bad arguments");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  if (Cout == 3 && ksize == 5 && stride == 2 && pad == 2 && Ho == 2 * H && Wo == 2 * W && Npad % 4 == 0 && Npad >= 76 &&
      N <= 65535 && (reinterpret_cast<uintptr_t>(cols) & 15u) == 0 && !knobs().patch_generic) {
    rc = optin_max_smem(reinterpret_cast<const void *>(col2im_k5s2_c3_kernel), dp);
    if (rc != CAI_OK) return rc;
    dim3 tg((Wo + kC2iTX - 1) / kC2iTX, (Ho + kC2iTY - 1) / kC2iTY, N);
    col2im_k5s2_c3_kernel<<<tg, 256, kC2iSmem, static_cast<cudaStream_t>(stream_)>>>(
        cols, bias, H, W, Ho, Wo, Npad, out_layout, clamp_lo, clamp_hi, out);
  } else {
    col2im_kernel<<<ew_grid2(dp, static_cast<int64_t>(N) * Ho * Wo * Cout), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        cols, bias, N, Cout, H, W, Ho, Wo, ksize, stride, pad, Npad, out_layout, clamp_lo, clamp_hi, out);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

}  // extern "C"
