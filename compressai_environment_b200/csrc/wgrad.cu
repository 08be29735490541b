// wgrad.cu -- training-mode weight gradient of the conv / transposed-conv layers as a tcgen05 GEMM over pixels.
//
// Reference: autograd of nn.Conv2d / nn.ConvTranspose2d as built by compressai/models/utils.py:128-146 and driven by
// examples/train.py:132-165 (the reference leaves this to cuDNN).  Both layer kinds reduce to ONE contraction between
// the tensor on the coarse grid ("small": dY of a conv, X of a transposed conv) and the tensor on the fine grid
// ("big": X of a conv, dY of a transposed conv):
//
//   grad[cs][cb][ky][kx] = sum over (n, i, j) of  small[n, cs, i, j] * big[n, cb, stride*i + ky - pad, stride*j + kx - pad]
//
// which is exactly dW[cout][cin][ky][kx] of the conv and dW[cin][cout][ky][kx] of the transposed conv.
//
// GEMM view, per tap (ky, kx):  D[cs, cb] = A[cs, K] * B[cb, K]^T with K = pixels of the coarse grid.  NCHW storage
// makes the pixel index the contiguous one, so both operands are K-major and arrive by TMA tensor loads straight into
// the 128-byte-swizzled UMMA layout (64 pixels = 128 B per channel row, SASS UTMALDG):
//   * `small` is split into bf16 hi / lo planes [N][Cs][Hs][Wsp] (rows padded to 8 pixels = 16 B);
//   * `big` is split into bf16 hi / lo COLUMN-GATHERED planes [N][Cb][stride * ksize][Hh][Wsp], variant
//     v = (row phase py, kx):  plane[v][yh][j] = big[stride*yh + py][stride*j + kx - pad]  (zero outside).  TMA can
//     neither stride the contiguous dimension nor start a box row at an address that is not a multiple of 16 bytes,
//     so the horizontal part of a tap (stride and the kx shift) is resolved once by this pre-pass; a tap (ky, kx) then
//     reads the box at column j0 of variant ((ky - pad) mod stride, kx), shifted VERTICALLY by floor((ky - pad) /
//     stride) rows -- rows are a free TMA coordinate, and rows above / below the plane are the tensor map's
//     out-of-bounds zero fill (the vertical conv padding).
// A K block is a (bx x by) box of the coarse grid with bx * by = 64 (bx = 64, 32, 16 or 8 for narrow grids).  The tensor
// maps order their dimensions (x, channel, y, ...) so that a box lands as `by` sub-tiles [channel][bx pixels]: each is
// a K-major tile with a row pitch of 128 / 64 / 32 / 16 bytes, i.e. the SWIZZLE_128B / 64B / 32B / no-swizzle canonical
// layouts, and the UMMA descriptors of the four k16 steps of a K block walk the sub-tiles.  One CTA owns a group of taps (as many fp32
// accumulators as fit the 512 TMEM columns: 4 for 128 channels, all 25 for the 3-channel layers), one 128-row tile of
// cs, one tile of cb and one slice of the K blocks (split-K over the grid); the A box of a K block is loaded once and
// used by every tap of the group.  Products are split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) like the
// forward kernels.  Partial sums go to a workspace and a second kernel reduces the K slices in a fixed order
// (deterministic) while transposing to the weight layout.
#include <cuda.h>

#include <cstdlib>

#include "conv_common.cuh"

namespace cai {

constexpr int kWgThreads = 192;  // warps 0-3: epilogue (TMEM lanes), warp 4: TMA producer, warp 5: MMA issuer
constexpr int kWgKB = 64;        // pixels per K block (128 bytes of bf16: one swizzle row)
constexpr int kWgASlots = 2;
constexpr int kWgMaxBSlots = 6;
constexpr uint32_t kWgAPlane = 128u * 128u;  // 128 cs rows x 128 B

struct WgradParams {
  float *partial;       // [ksplit][ntaps][Mrows][Ncols] fp32
  int ntaps, tpc;       // taps in the layer, taps per CTA
  int tgroups, mtiles, ntiles, ksplit;
  int Nt;               // cb rows per tile (multiple of 16, <= 256)
  int Mrows, Ncols;     // mtiles * 128, ntiles * Nt
  int kbx, kby, bx, by; // K blocks per image row / column, box extents
  int kblocks;          // N * kby * kbx
  int b_slots;
  int mode;            // operand layout by box width: 0 bx = 64, 1 bx = 32, 2 bx = 16, 3 bx = 8 (see wg_desc)
  int debug;           // CAI_WGRAD_DEBUG (debug builds of the experiment knob): 1 skip MMAs, 2 skip TMA loads
  uint32_t b_plane;     // Nt * 128 bytes
  int8_t sy[kMaxTaps], variant[kMaxTaps];  // vertical window shift and column-gathered plane of each tap
};

// UMMA shared-memory descriptor of k16 step `kk` (0..3) of one operand plane of a K block.  The plane holds 64 / bx
// sub-tiles [rows][bx pixels] back to back (mode 0: bx = 64, one SWIZZLE_128B tile; 1: bx = 32, two SWIZZLE_64B tiles;
// 2: bx = 16, four SWIZZLE_32B tiles; 3: bx = 8, eight no-swizzle tiles = core-matrix columns, two per k16 step).
__device__ __forceinline__ uint64_t wg_desc(uint32_t base, uint32_t rows, int kk, int mode) {
  uint32_t addr, lbo = 16u, sbo;
  uint64_t type;
  if (mode == 0) {
    addr = base + static_cast<uint32_t>(kk) * 32u;  sbo = 1024u;  type = 2;   // SWIZZLE_128B
  } else if (mode == 1) {
    addr = base + static_cast<uint32_t>(kk >> 1) * rows * 64u + static_cast<uint32_t>(kk & 1) * 32u;  sbo = 512u;  type = 4;
  } else if (mode == 2) {
    addr = base + static_cast<uint32_t>(kk) * rows * 32u;  sbo = 256u;  type = 6;
  } else {
    addr = base + static_cast<uint32_t>(2 * kk) * rows * 16u;  lbo = rows * 16u;  sbo = 128u;  type = 0;
  }
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= type << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_a(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4,
                                              uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ CUtensorMap map_s_hi,
             const __grid_constant__ CUtensorMap map_s_lo, const __grid_constant__ CUtensorMap map_b_hi,
             const __grid_constant__ CUtensorMap map_b_lo) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kWgASlots], a_empty[kWgASlots], b_full[kWgMaxBSlots], b_empty[kWgMaxBSlots];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sm_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  const uint32_t sm_a = sm_base;
  const uint32_t sm_b = sm_base + kWgASlots * 2u * kWgAPlane;
  const uint32_t b_slot_bytes = 2u * p.b_plane;

  // work decomposition: blockIdx.x -> (tap group, m tile, n tile, k slice)
  int w = blockIdx.x;
  const int ks = w % p.ksplit;  w /= p.ksplit;
  const int nt = w % p.ntiles;  w /= p.ntiles;
  const int mt = w % p.mtiles;  w /= p.mtiles;
  const int tg = w;
  const int t0 = tg * p.tpc;
  const int nt_here = (p.ntaps - t0 < p.tpc) ? (p.ntaps - t0) : p.tpc;
  const int kb0 = static_cast<int>(static_cast<int64_t>(p.kblocks) * ks / p.ksplit);
  const int kb1 = static_cast<int>(static_cast<int64_t>(p.kblocks) * (ks + 1) / p.ksplit);

  if (tid == 0) {
    for (int s = 0; s < kWgASlots; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kWgMaxBSlots; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int per_img = p.kby * p.kbx;

  if (warp == 4) {
    if (lane == 0 && kb1 > kb0) {
      int sa = 0, sb = 0;
      uint32_t a_pass = 0, b_pass = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int n = kb / per_img, rem = kb - n * per_img;
        const int i0 = (rem / p.kbx) * p.by, j0 = (rem % p.kbx) * p.bx;
        if (a_pass > 0) mbar_wait_bounded(&a_empty[sa], (a_pass - 1) & 1);
        const uint32_t adst = sm_a + static_cast<uint32_t>(sa) * 2u * kWgAPlane;
        if (p.debug & (2 | 4)) {
          mbar_arrive(&a_full[sa]);
        } else {
          mbar_expect_tx(&a_full[sa], 2u * kWgAPlane);
          tma_load_4d(adst, &map_s_hi, j0, mt * 128, i0, n, &a_full[sa]);
          tma_load_4d(adst + kWgAPlane, &map_s_lo, j0, mt * 128, i0, n, &a_full[sa]);
        }
        if (++sa == kWgASlots) {
          sa = 0;
          ++a_pass;
        }
        for (int u = 0; u < nt_here; ++u) {
          const int t = t0 + u;
          if (b_pass > 0) mbar_wait_bounded(&b_empty[sb], (b_pass - 1) & 1);
          const uint32_t bdst = sm_b + static_cast<uint32_t>(sb) * b_slot_bytes;
          if (p.debug & (2 | 8)) {
            mbar_arrive(&b_full[sb]);
          } else {
            mbar_expect_tx(&b_full[sb], b_slot_bytes);
            tma_load_5d_a(bdst, &map_b_hi, j0, nt * p.Nt, i0 + p.sy[t], p.variant[t], n, &b_full[sb]);
            tma_load_5d_a(bdst + p.b_plane, &map_b_lo, j0, nt * p.Nt, i0 + p.sy[t], p.variant[t], n, &b_full[sb]);
          }
          if (++sb == p.b_slots) {
            sb = 0;
            ++b_pass;
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && kb1 > kb0) {
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = Nt, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.Nt >> 3) << 17) |
                             (static_cast<uint32_t>(128 >> 4) << 24);
      int sa = 0, sb = 0;
      uint32_t a_par = 0, b_par = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_bounded(&a_full[sa], a_par);
        const uint32_t a_hi = sm_a + static_cast<uint32_t>(sa) * 2u * kWgAPlane, a_lo = a_hi + kWgAPlane;
        for (int u = 0; u < nt_here; ++u) {
          mbar_wait_bounded(&b_full[sb], b_par);
          tc_fence_after();
          const uint32_t b_hi = sm_b + static_cast<uint32_t>(sb) * b_slot_bytes, b_lo = b_hi + p.b_plane;
          const uint32_t d = tmem_base + static_cast<uint32_t>(u * p.Nt);
          const uint32_t acc0 = (kb > kb0) ? 1u : 0u;
#pragma unroll
          for (int kk = 0; kk < kWgKB / 16; ++kk) {
            if (p.debug & 1) break;
            const uint64_t dah = wg_desc(a_hi, 128u, kk, p.mode), dal = wg_desc(a_lo, 128u, kk, p.mode);
            const uint64_t dbh = wg_desc(b_hi, static_cast<uint32_t>(p.Nt), kk, p.mode);
            const uint64_t dbl = wg_desc(b_lo, static_cast<uint32_t>(p.Nt), kk, p.mode);
            umma_bf16(d, dah, dbh, idesc, (kk > 0) ? 1u : acc0);
            umma_bf16(d, dah, dbl, idesc, 1u);
            umma_bf16(d, dal, dbh, idesc, 1u);
          }
          umma_commit(&b_empty[sb]);
          if (++sb == p.b_slots) {
            sb = 0;
            b_par ^= 1u;
          }
        }
        umma_commit(&a_empty[sa]);
        if (++sa == kWgASlots) {
          sa = 0;
          a_par ^= 1u;
        }
      }
      umma_commit(&acc_full);
    }
  } else {
    // epilogue: thread = TMEM lane = cs row of the tile; fp32 partial sums -> workspace
    const int row = tid;  // 0..127
    const uint32_t lane_base = (static_cast<uint32_t>(warp) * 32u) << 16;
    const bool have = kb1 > kb0;
    if (have) {
      mbar_wait_bounded(&acc_full, 0);
      tc_fence_after();
    }
    for (int u = 0; u < nt_here; ++u) {
      float *dst = p.partial + ((static_cast<int64_t>(ks) * p.ntaps + (t0 + u)) * p.Mrows + mt * 128 + row) * p.Ncols +
                   static_cast<int64_t>(nt) * p.Nt;
      for (int c0 = 0; c0 < p.Nt; c0 += 16) {
        uint32_t v[16];
        if (have) {
          tmem_ld16(tmem_base + lane_base + static_cast<uint32_t>(u * p.Nt + c0), v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        float4 *o = reinterpret_cast<float4 *>(dst + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                             __uint_as_float(v[4 * q + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// fp32 NCHW [NC][H][W] -> bf16 hi / lo planes [NC][nv][Hh][Wp]:  out[v][yh][j] = x[s*yh + v / kv][s*j + v % kv - pad]
// (zero outside the image).  `small`: nv = 1, s = 1, kv = 1, pad = 0 (plain planes with rows padded to Wp).
__global__ void wgrad_split_kernel(const float *__restrict__ x, int64_t NC, int H, int W, int s, int kv, int pad, int nv,
                                   int Hh, int Wj, int Wp, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo) {
  const int64_t total = NC * nv * Hh * Wp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += stride) {
    int64_t r = o;
    const int j = static_cast<int>(r % Wp);   r /= Wp;
    const int yh = static_cast<int>(r % Hh);  r /= Hh;
    const int v = static_cast<int>(r % nv);   r /= nv;
    const int y = yh * s + v / kv, xx = j * s + v % kv - pad;
    float val = 0.f;
    if (j < Wj && y < H && xx >= 0 && xx < W) val = __ldg(x + (r * H + y) * W + xx);
    __nv_bfloat16 h, l;
    split_bf16(val, h, l);
    hi[o] = h;
    lo[o] = l;
  }
}

// grad[cs][cb][tap] = sum over k slices of partial[ks][tap][cs][cb]   (fixed order: deterministic)
__global__ void wgrad_reduce_kernel(const float *__restrict__ partial, int ksplit, int ntaps, int Mrows, int Ncols, int Cs,
                                    int Cb, float *__restrict__ grad) {
  const int64_t total = static_cast<int64_t>(ntaps) * Cs * Cb;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t slice = static_cast<int64_t>(ntaps) * Mrows * Ncols;
  for (int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += stride) {
    const int cb = static_cast<int>(o % Cb);
    const int cs = static_cast<int>((o / Cb) % Cs);
    const int t = static_cast<int>(o / (static_cast<int64_t>(Cb) * Cs));
    const float *src = partial + (static_cast<int64_t>(t) * Mrows + cs) * Ncols + cb;
    float acc = 0.f;
    for (int k = 0; k < ksplit; ++k) acc += __ldg(src + k * slice);
    grad[(static_cast<int64_t>(cs) * Cb + cb) * ntaps + t] = acc;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn wg_encode_fn() {
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

struct WgradPlan {
  int Wsp, Hh, nv;
  int bx, by, kbx, kby, kblocks;
  int Nt, ntiles, mtiles, tpc, tgroups, ksplit, ntaps, b_slots;
  int64_t off_s_hi, off_s_lo, off_b_hi, off_b_lo, off_partial, total;
  uint32_t smem;
};

static int64_t up256(int64_t v) { return (v + 255) & ~static_cast<int64_t>(255); }

static bool plan_wgrad(int64_t N, int Cs, int Hs, int Ws, int Cb, int Hb, int Wb, int ksize, int stride, int sm_count,
                       WgradPlan *pl) {
  if (N < 1 || Cs < 1 || Hs < 1 || Ws < 1 || Cb < 1 || Hb < 1 || Wb < 1) return false;
  if (ksize < 1 || ksize * ksize > kMaxTaps || stride < 1 || stride > 2) return false;
  pl->ntaps = ksize * ksize;
  pl->nv = stride * ksize;
  pl->Wsp = (Ws + 7) & ~7;
  pl->Hh = (Hb + stride - 1) / stride;
  int bx = 8;
  while (bx < 64 && bx < Ws) bx <<= 1;
  pl->bx = bx;
  pl->by = kWgKB / bx;
  pl->kbx = (Ws + bx - 1) / bx;
  pl->kby = (Hs + pl->by - 1) / pl->by;
  const int64_t kblocks = N * pl->kby * pl->kbx;
  if (kblocks >= (1ll << 31)) return false;
  pl->kblocks = static_cast<int>(kblocks);
  const int cb16 = (Cb + 15) & ~15;
  pl->ntiles = (cb16 + 255) / 256;
  pl->Nt = (((cb16 + pl->ntiles - 1) / pl->ntiles) + 15) & ~15;
  pl->mtiles = (Cs + 127) / 128;
  pl->tpc = 512 / pl->Nt;
  if (pl->tpc > pl->ntaps) pl->tpc = pl->ntaps;
  pl->tgroups = (pl->ntaps + pl->tpc - 1) / pl->tpc;
  const int units = pl->tgroups * pl->mtiles * pl->ntiles;
  int ksplit = sm_count / units;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > pl->kblocks) ksplit = pl->kblocks;
  pl->ksplit = ksplit;
  const uint32_t b_slot = 2u * static_cast<uint32_t>(pl->Nt) * 128u;
  int b_slots = static_cast<int>((160u * 1024u) / b_slot);
  if (b_slots > kWgMaxBSlots) b_slots = kWgMaxBSlots;
  if (b_slots < 2) b_slots = 2;
  pl->b_slots = b_slots;
  pl->smem = 1024u + kWgASlots * 2u * kWgAPlane + static_cast<uint32_t>(b_slots) * b_slot;
  int64_t off = 0;
  const int64_t s_plane = N * Cs * static_cast<int64_t>(Hs) * pl->Wsp * 2;
  const int64_t b_plane = N * Cb * static_cast<int64_t>(pl->nv) * pl->Hh * pl->Wsp * 2;
  pl->off_s_hi = off;  off += up256(s_plane);
  pl->off_s_lo = off;  off += up256(s_plane);
  pl->off_b_hi = off;  off += up256(b_plane);
  pl->off_b_lo = off;  off += up256(b_plane);
  pl->off_partial = off;
  off += up256(static_cast<int64_t>(ksplit) * pl->ntaps * pl->mtiles * 128 * pl->ntiles * pl->Nt * 4);
  pl->total = off;
  return true;
}

}  // namespace cai

using namespace cai;

extern "C" {

int64_t cai_conv_wgrad_workspace(int64_t N, int32_t Cs, int32_t Hs, int32_t Ws, int32_t Cb, int32_t Hb, int32_t Wb,
                                 int32_t ksize, int32_t stride) {
  DeviceProps dp;
  if (get_device_props(&dp) != CAI_OK) return -1;
  WgradPlan pl;
  if (!plan_wgrad(N, Cs, Hs, Ws, Cb, Hb, Wb, ksize, stride, dp.sm_count, &pl)) {
    set_error("cai_conv_wgrad_workspace: unsupported geometry (ksize^2 <= %d, stride 1 or 2)", kMaxTaps);
    return -1;
  }
  return pl.total;
}

int cai_conv_wgrad(const float *small, const float *big, int64_t N, int32_t Cs, int32_t Hs, int32_t Ws, int32_t Cb,
                   int32_t Hb, int32_t Wb, int32_t ksize, int32_t stride, int32_t pad, float *grad, void *workspace,
                   int64_t workspace_bytes, cai_stream_t stream_) {
  CAI_CHECK_ARG(small && big && grad && workspace, "cai_conv_wgrad: NULL pointer");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  WgradPlan pl;
  CAI_CHECK_ARG(plan_wgrad(N, Cs, Hs, Ws, Cb, Hb, Wb, ksize, stride, dp.sm_count, &pl),
                "cai_conv_wgrad: unsupported geometry (ksize^2 <= %d, stride 1 or 2)", kMaxTaps);
  CAI_CHECK_ARG(workspace_bytes >= pl.total, "cai_conv_wgrad: workspace too small (%lld < %lld)",
                static_cast<long long>(workspace_bytes), static_cast<long long>(pl.total));
  CAI_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, "cai_conv_wgrad: workspace must be 256-byte aligned");
  EncodeTiledFn enc = wg_encode_fn();
  CAI_CHECK_ARG(enc != nullptr, "cai_conv_wgrad: cuTensorMapEncodeTiled is not available in this driver");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  unsigned char *ws = static_cast<unsigned char *>(workspace);
  __nv_bfloat16 *s_hi = reinterpret_cast<__nv_bfloat16 *>(ws + pl.off_s_hi);
  __nv_bfloat16 *s_lo = reinterpret_cast<__nv_bfloat16 *>(ws + pl.off_s_lo);
  __nv_bfloat16 *b_hi = reinterpret_cast<__nv_bfloat16 *>(ws + pl.off_b_hi);
  __nv_bfloat16 *b_lo = reinterpret_cast<__nv_bfloat16 *>(ws + pl.off_b_lo);
  float *partial = reinterpret_cast<float *>(ws + pl.off_partial);

  {
    const int64_t tot_s = N * Cs * static_cast<int64_t>(Hs) * pl.Wsp;
    const int64_t tot_b = N * Cb * static_cast<int64_t>(pl.nv) * pl.Hh * pl.Wsp;
    const int64_t cap = static_cast<int64_t>(dp.sm_count) * 16;
    int64_t g1 = (tot_s + 255) / 256, g2 = (tot_b + 255) / 256;
    if (g1 > cap) g1 = cap;
    if (g2 > cap) g2 = cap;
    wgrad_split_kernel<<<static_cast<int>(g1), 256, 0, st>>>(small, N * Cs, Hs, Ws, 1, 1, 0, 1, Hs, Ws, pl.Wsp, s_hi, s_lo);
    CAI_LAUNCH_CHECK();
    wgrad_split_kernel<<<static_cast<int>(g2), 256, 0, st>>>(big, N * Cb, Hb, Wb, stride, ksize, pad, pl.nv, pl.Hh, Ws,
                                                              pl.Wsp, b_hi, b_lo);
    CAI_LAUNCH_CHECK();
  }

  static const int dbg = [] { const char *v = getenv("CAI_WGRAD_DEBUG"); return (v && *v) ? atoi(v) : 0; }();
  const int mode = pl.bx == 64 ? 0 : pl.bx == 32 ? 1 : pl.bx == 16 ? 2 : 3;
  const CUtensorMapSwizzle swz = mode == 0   ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : mode == 2 ? CU_TENSOR_MAP_SWIZZLE_32B
                                             : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap maps[4];
  {
    // dimensions (x, channel, y, n): the box lands as `by` sub-tiles [128 channels][bx pixels]
    const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(Ws), static_cast<cuuint64_t>(Cs), static_cast<cuuint64_t>(Hs),
                                static_cast<cuuint64_t>(N)};
    const cuuint64_t row = static_cast<cuuint64_t>(pl.Wsp) * 2;
    const cuuint64_t gstr[3] = {row * Hs, row, row * Hs * Cs};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(pl.bx), 128, static_cast<cuuint32_t>(pl.by), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    void *planes[2] = {s_hi, s_lo};
    for (int q = 0; q < 2; ++q) {
      const CUresult r = enc(&maps[q], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, planes[q], gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cai_conv_wgrad: cuTensorMapEncodeTiled (small) failed (%d)", static_cast<int>(r));
        return CAI_E_CUDA;
      }
    }
  }
  {
    // dimensions (x, channel, y, variant, n)
    const cuuint64_t gdim[5] = {static_cast<cuuint64_t>(Ws), static_cast<cuuint64_t>(Cb), static_cast<cuuint64_t>(pl.Hh),
                                static_cast<cuuint64_t>(pl.nv), static_cast<cuuint64_t>(N)};
    const cuuint64_t row = static_cast<cuuint64_t>(pl.Wsp) * 2;
    const cuuint64_t gstr[4] = {row * pl.Hh * pl.nv, row, row * pl.Hh, row * pl.Hh * pl.nv * Cb};
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(pl.bx), static_cast<cuuint32_t>(pl.Nt),
                               static_cast<cuuint32_t>(pl.by), 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void *planes[2] = {b_hi, b_lo};
    for (int q = 0; q < 2; ++q) {
      const CUresult r = enc(&maps[2 + q], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, planes[q], gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cai_conv_wgrad: cuTensorMapEncodeTiled (big) failed (%d)", static_cast<int>(r));
        return CAI_E_CUDA;
      }
    }
  }

  WgradParams p{};
  p.partial = partial;
  p.ntaps = pl.ntaps;
  p.tpc = pl.tpc;
  p.tgroups = pl.tgroups;
  p.mtiles = pl.mtiles;
  p.ntiles = pl.ntiles;
  p.ksplit = pl.ksplit;
  p.Nt = pl.Nt;
  p.Mrows = pl.mtiles * 128;
  p.Ncols = pl.ntiles * pl.Nt;
  p.kbx = pl.kbx;
  p.kby = pl.kby;
  p.bx = pl.bx;
  p.by = pl.by;
  p.kblocks = pl.kblocks;
  p.b_slots = pl.b_slots;
  p.b_plane = static_cast<uint32_t>(pl.Nt) * 128u;
  p.debug = dbg;
  p.mode = mode;
  for (int ky = 0; ky < ksize; ++ky)
    for (int kx = 0; kx < ksize; ++kx) {
      const int t = ky * ksize + kx;
      const int oy = ky - pad;
      const int py = ((oy % stride) + stride) % stride;
      p.sy[t] = static_cast<int8_t>((oy - py) / stride);
      p.variant[t] = static_cast<int8_t>(py * ksize + kx);
    }
  int max_dyn = 0;
  rc = optin_max_smem(reinterpret_cast<const void *>(wgrad_kernel), dp, &max_dyn);
  if (rc != CAI_OK) return rc;
  CAI_CHECK_ARG(static_cast<int>(pl.smem) <= max_dyn, "cai_conv_wgrad: %u bytes of shared memory exceed the limit %d",
                pl.smem, max_dyn);
  const int grid = pl.tgroups * pl.mtiles * pl.ntiles * pl.ksplit;
  wgrad_kernel<<<grid, kWgThreads, pl.smem, st>>>(p, maps[0], maps[1], maps[2], maps[3]);
  CAI_LAUNCH_CHECK();
  {
    const int64_t tot = static_cast<int64_t>(pl.ntaps) * Cs * Cb;
    int64_t g = (tot + 255) / 256;
    const int64_t cap = static_cast<int64_t>(dp.sm_count) * 16;
    if (g > cap) g = cap;
    wgrad_reduce_kernel<<<static_cast<int>(g), 256, 0, st>>>(partial, pl.ksplit, pl.ntaps, p.Mrows, p.Ncols, Cs, Cb, grad);
    CAI_LAUNCH_CHECK();
  }
  return CAI_OK;
}

}  // extern "C"
