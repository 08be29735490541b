// ar.cu -- autoregressive context-model scan for the joint autoregressive + hierarchical prior models.
//
// What it replaces (reference tree):
//   compressai/models/google.py:535-577   JointAutoregressiveHierarchicalPriors._compress_ar
//   compressai/models/google.py:620-661   JointAutoregressiveHierarchicalPriors._decompress_ar
//   (inherited unchanged by Cheng2020Anchor / Cheng2020Attention, compressai/models/waseda.py:44-153)
//
// The reference walks the latent grid in raster order in Python; per pixel it runs a masked 5x5 convolution on a
// crop of y_hat (context_prediction, layers.py:52-78), three 1x1 convolutions (entropy_parameters) on
// cat(params, ctx), build_indexes, and then either quantises y around the predicted mean (encoder) or decodes M
// symbols from the image's single rANS stream (decoder).  Pixel (h, w) needs y_hat of (h, w-1), so both directions
// are strictly sequential inside an image; the parallelism that exists is
//   * across images,
//   * across the output rows of the four matrix-vector products of one pixel (7.6 MB of fp32 weights for M = 192).
//
// Design: ONE THREAD-BLOCK CLUSTER per group of G <= R images, persistent over the whole grid.  The R CTAs of the
// cluster split the output rows of every layer; a warp computes whole rows (fixed summation order: the result of a
// row never depends on R, G or the launch shape -- the encoder and the decoder MUST compute bit-identical Gaussian
// parameters), streaming the weights from L2 with 16-byte loads, and scatters the row's result into the activation
// vectors of all R CTAs through distributed shared memory; barrier.cluster separates the layers.  G images share
// one pass over the weights (G accumulators per row).  Image g of the group is owned by CTA rank g: it holds the
// packed CDF rows in shared memory (one TMA bulk copy), and one of its warps runs the image's rANS chain (32-ary
// cooperative search in the uint16 row, words prefetched one per lane); the other ranks wait at the cluster barrier.
// y_hat lives in HBM as a zero-padded NHWC tensor; readers fetch the 12 causal taps with L2 loads (ld.global.cg)
// after the barrier that follows the owner's store.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace cai {

constexpr int kArThreads = 1024;
constexpr int kArWarps = kArThreads / 32;
constexpr int kArMaxCluster = 8;

struct ArKernelParams {
  const float *w_ctx, *b_ctx, *w1, *b1, *w2, *b2, *w3, *b3;
  const float *params;  // [B, H, W, P]
  const float *y;       // encoder: [B, H, W, M]
  float *y_hat;         // [B, H + 2 pad, W + 2 pad, M], border zero
  int32_t *sym, *idx;   // [B, H * W * M] pixel-major, channel-minor (the order the reference pushes them)
  const float *scale_table;
  const unsigned char *blob;  // packed CDF table (decoder)
  const uint32_t *words;
  const int64_t *word_begin;  // [B + 1]
  int32_t *status;            // [B]
  float bound, slope;
  uint32_t enc_bytes;    // header + row metadata + CDF rows of the blob
  uint32_t stage_bytes;  // what the decoder stages: enc_bytes, or the whole blob (with the decode LUT) when it fits
  int T, B, H, W, M, P, n_ctx, n1, n2, n3, ksize, pad;
  int K0, K1p, K2p, K3p;  // padded reduction lengths (multiples of 4)
  int R;
};

__host__ __device__ inline int pad4(int v) { return (v + 3) & ~3; }

// One layer for G images: rows split across the R CTAs, one warp per row, result scattered through DSMEM.
//   in  : shared, [G][ldin] (zero padded up to Kp)        out : shared, [G][ldout], element out_off + row
//   owner_only: row results of image g go to CTA g only (last layer), otherwise to every CTA of the cluster.
template <int G>
__device__ __forceinline__ void ar_layer(cg::cluster_group &cluster, const float *__restrict__ Wm,
                                         const float *__restrict__ bias, int rows, int Kp, const float *in, int ldin,
                                         float *out, int ldout, int out_off, bool leaky, float slope, bool owner_only,
                                         int R, int rank) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (rows + R - 1) / R;
  const int r0 = rank * per;
  const int r1 = min(rows, r0 + per);
  const int k4n = Kp >> 2;
  for (int r = r0 + warp; r < r1; r += kArWarps) {
    float acc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g] = 0.0f;
    const float4 *wr = reinterpret_cast<const float4 *>(Wm + static_cast<size_t>(r) * Kp);
#pragma unroll 6
    for (int k4 = lane; k4 < k4n; k4 += 32) {
      const float4 w = __ldg(wr + k4);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 x = *reinterpret_cast<const float4 *>(in + g * ldin + 4 * k4);
        acc[g] = fmaf(w.x, x.x, acc[g]);
        acc[g] = fmaf(w.y, x.y, acc[g]);
        acc[g] = fmaf(w.z, x.z, acc[g]);
        acc[g] = fmaf(w.w, x.w, acc[g]);
      }
    }
    const float b = __ldg(bias + r);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float v = acc[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      v += b;
      if (leaky) v = (v > 0.0f) ? v : v * slope;
      float *dst = out + g * ldout + out_off + r;
      if (owner_only) {
        if (lane == 0) *cluster.map_shared_rank(dst, g) = v;
      } else if (lane < R) {
        *cluster.map_shared_rank(dst, lane) = v;
      }
    }
  }
}

// GaussianConditional.build_indexes for one value (entropy_models.py:684-689; same arithmetic as quantize.cu).
__device__ __forceinline__ int32_t ar_index_of(float s, float bound, const float *t, int T) {
  s = (s < bound) ? bound : s;
  int lo = 0, hi = T - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s <= t[mid])
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ int32_t ar_quant_sym(float y, float mean) {
  const float r = rintf(__fsub_rn(y, mean));
  if (!(r >= -2147483648.0f && r < 2147483648.0f)) return static_cast<int32_t>(0x80000000u);
  return static_cast<int32_t>(r);
}

// Streaming rANS decoder state of one image, replicated in every lane of the decoding warp
// (rans_interface.cpp:286-359: set_stream + decode_stream; rans64.h:104-142).
struct ArDec {
  uint64_t x;
  int64_t pos, n_words;  // next word, words in the string
  int64_t base;          // word index held by lane 0 of `buf`
  const uint32_t *w;
  uint32_t buf;
  int lane;
  bool truncated;
  __device__ __forceinline__ void fill(int64_t b) {
    base = b;
    const int64_t i = b + lane;
    buf = (i < n_words) ? __ldcg(w + i) : 0u;
  }
  __device__ __forceinline__ uint32_t next_word() {
    if (pos >= base + 32) fill(pos);
    if (pos >= n_words) truncated = true;
    const uint32_t v = __shfl_sync(0xffffffffu, buf, static_cast<int>(pos - base));
    pos += 1;
    return v;
  }
  __device__ __forceinline__ void init(const uint32_t *words, int64_t n, int ln) {
    w = words;
    n_words = n;
    lane = ln;
    truncated = false;
    pos = 0;
    fill(0);
    const uint64_t lo = next_word(), hi = next_word();
    x = lo | (hi << 32);
  }
  __device__ __forceinline__ void renorm() {
    if (x < (1ull << 31)) x = (x << 32) | next_word();
  }
  __device__ __forceinline__ uint32_t get_bits4() {
    const uint32_t v = static_cast<uint32_t>(x) & 15u;
    x >>= 4;
    renorm();
    return v;
  }
};

// One symbol with CDF row metadata `m`; `cdf` / `lut` are the shared-memory copies of the packed table.  With the
// decode LUT staged (lut != nullptr; same entry format as rans.cu: {start | freq << 16, s0}, freq == 0 = "search forward
// from s0" or, with bit 31 of the second word, "whole bucket inside the escape symbol") the common symbol costs one
// LDS; without it the row is searched 32 ways at a time.
__device__ __forceinline__ int32_t ar_decode_symbol(ArDec &d, const RowMeta m, const uint16_t *cdf, const uint2 *lut,
                                                    int lut_shift) {
  const uint16_t *row = cdf + m.cdf_off;
  const uint32_t cf = static_cast<uint32_t>(d.x) & 0xffffu;
  const int32_t max_value = m.len - 2;
  int s;
  uint32_t start, freq;
  if (lut != nullptr) {
    const uint2 e = lut[m.lut_off + (cf >> lut_shift)];
    start = e.x & 0xffffu;
    freq = e.x >> 16;
    s = static_cast<int>(e.y);
    if (freq == 0u && (e.y & 0x80000000u)) {
      s = max_value;
      start = e.y & 0xffffu;
      freq = 0x10000u - start;
    } else if (freq == 0u) {
      for (;;) {  // lane i tests symbol s + i; exactly one lane can hit
        const int cand = s + d.lane;
        uint32_t packed = 0;
        if (cand <= max_value) {
          const uint32_t c_lo = row[cand];
          const uint32_t c_hi = (cand == max_value) ? 0x10000u : static_cast<uint32_t>(row[cand + 1]);
          if (c_lo <= cf && cf < c_hi) packed = ((c_hi - c_lo) << 16) | c_lo;
        }
        const uint32_t hit = __ballot_sync(0xffffffffu, packed != 0u);
        if (hit) {
          const int src = __ffs(hit) - 1;
          packed = __shfl_sync(0xffffffffu, packed, src);
          s += src;
          start = packed & 0xffffu;
          freq = packed >> 16;
          break;
        }
        s += 32;
        if (s > max_value) {  // malformed table: stay in bounds, keep the chain defined
          s = max_value < 0 ? 0 : max_value;
          start = row[s];
          freq = 1u;
          break;
        }
      }
    }
  } else {
    // s = #{ j < len - 1 : row[j] <= cf } - 1  (row[0] = 0; the terminal 65536 is stored as 0 and never compared)
    int lo = 0, cnt = m.len - 1;
    while (cnt > 32) {
      const int stride = (cnt + 31) >> 5;
      const int off = d.lane * stride;
      const bool le = off < cnt && row[lo + off] <= cf;
      const int c = __popc(__ballot_sync(0xffffffffu, le));  // >= 1: row[lo] <= cf by induction
      lo += (c - 1) * stride;
      cnt = min(stride, cnt - (c - 1) * stride);
    }
    const bool le = d.lane < cnt && row[lo + d.lane] <= cf;
    s = lo + __popc(__ballot_sync(0xffffffffu, le)) - 1;
    start = row[s];
    freq = (static_cast<uint32_t>(row[s + 1]) - start) & 0xffffu;
    if (freq == 0u) freq = 65536u;
  }
  d.x = static_cast<uint64_t>(freq) * (d.x >> 16) + cf - start;
  d.renorm();
  int32_t value = s;
  if (value == max_value) {  // bypass-coded tail (rans_interface.cpp:256-279)
    uint32_t val = d.get_bits4();
    int32_t nb = static_cast<int32_t>(val);
    while (val == 15u) {
      val = d.get_bits4();
      nb += static_cast<int32_t>(val);
    }
    uint32_t raw = 0;
    for (int32_t j = 0; j < nb; ++j) {
      val = d.get_bits4();
      if (j < 8) raw |= val << (4 * j);
    }
    const int32_t sraw = static_cast<int32_t>(raw);
    value = sraw >> 1;
    value = (sraw & 1) ? (-value - 1) : (value + max_value);
  }
  return value + m.offset;
}

template <int G, bool kDecode>
__global__ void __launch_bounds__(kArThreads, 1) ar_scan_kernel(const ArKernelParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  cg::cluster_group cluster = cg::this_cluster();
  const int R = p.R;
  const int rank = static_cast<int>(cluster.block_rank());
  const int group = blockIdx.x / R;
  const int b0 = group * G;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // shared memory: activation vectors of the G images | scale table | symbols, indexes | packed CDF rows
  const int ld0 = p.K0, ld1 = p.K1p, ld2 = p.K2p, ld3 = p.K3p, ld4 = pad4(p.n3);
  float *v0 = reinterpret_cast<float *>(smem_raw);
  float *cat = v0 + G * ld0;
  float *h1 = cat + G * ld1;
  float *h2 = h1 + G * ld2;
  float *gp = h2 + G * ld3;
  float *s_tab = gp + G * ld4;
  int32_t *s_idx = reinterpret_cast<int32_t *>(s_tab + pad4(p.T));
  int32_t *s_sym = s_idx + pad4(p.M);
  unsigned char *s_blob = reinterpret_cast<unsigned char *>(s_sym + pad4(p.M));

  for (int i = tid; i < G * (ld1 + ld2 + ld3 + ld4); i += kArThreads) cat[i] = 0.0f;  // zero the K padding once
  for (int i = tid; i < p.T; i += kArThreads) s_tab[i] = p.scale_table[i];
  const RowMeta *meta = nullptr;
  const uint16_t *cdf = nullptr;
  const uint2 *lut = nullptr;  // staged only when the whole blob fits beside the activation vectors (p.stage_bytes)
  int lut_shift = 0;
  if (kDecode) {
    stage_blob(s_blob, p.blob, p.stage_bytes, &s_bar);
    const BlobHeader *hdr = reinterpret_cast<const BlobHeader *>(s_blob);
    meta = reinterpret_cast<const RowMeta *>(s_blob + hdr->off_meta);
    cdf = reinterpret_cast<const uint16_t *>(s_blob + hdr->off_cdf);
    if (p.stage_bytes > p.enc_bytes) {
      lut = reinterpret_cast<const uint2 *>(s_blob + hdr->off_lut);
      lut_shift = hdr->lut_shift;
    }
  }
  const int my_b = b0 + rank;  // image owned by this CTA (if rank < G and inside the batch)
  const bool owner = rank < G && my_b < p.B;
  ArDec dec;
  if (kDecode && owner && warp == 0) {
    const int64_t wb = p.word_begin[my_b], we = p.word_begin[my_b + 1];
    dec.init(p.words + wb, we - wb, lane);
  }
  cluster.sync();

  const int Hp = p.H + 2 * p.pad, Wp = p.W + 2 * p.pad;
  const int half = p.ksize >> 1;
  const int rowlen = p.ksize * p.M;  // one full kernel row of taps is contiguous in NHWC
  for (int h = 0; h < p.H; ++h) {
    for (int w = 0; w < p.W; ++w) {
      // A. gather: the causal taps of y_hat (rows h-half .. h-1 complete, row h up to w-1) and this pixel's params
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int b = min(b0 + g, p.B - 1);
        const float *yh = p.y_hat + ((static_cast<size_t>(b) * Hp + h) * Wp + w) * p.M;
        for (int i = tid * 4; i < p.K0; i += kArThreads * 4) {
          const int ky = i / rowlen, rem = i - ky * rowlen;
          const float4 v = __ldcg(reinterpret_cast<const float4 *>(yh + static_cast<size_t>(ky) * Wp * p.M + rem));
          *reinterpret_cast<float4 *>(v0 + g * ld0 + i) = v;
        }
        const float *pp = p.params + ((static_cast<size_t>(b) * p.H + h) * p.W + w) * p.P;
        for (int i = tid; i < p.P; i += kArThreads) cat[g * ld1 + i] = __ldg(pp + i);
      }
      __syncthreads();
      // B. masked context convolution -> cat[P : P + n_ctx]
      ar_layer<G>(cluster, p.w_ctx, p.b_ctx, p.n_ctx, p.K0, v0, ld0, cat, ld1, p.P, false, 0.f, false, R, rank);
      cluster.sync();
      // C-E. entropy_parameters
      ar_layer<G>(cluster, p.w1, p.b1, p.n1, p.K1p, cat, ld1, h1, ld2, 0, true, p.slope, false, R, rank);
      cluster.sync();
      ar_layer<G>(cluster, p.w2, p.b2, p.n2, p.K2p, h1, ld2, h2, ld3, 0, true, p.slope, false, R, rank);
      cluster.sync();
      ar_layer<G>(cluster, p.w3, p.b3, p.n3, p.K3p, h2, ld3, gp, ld4, 0, false, 0.f, true, R, rank);
      cluster.sync();
      // F. owner: indexes, then quantise (encoder) or decode (decoder); y_hat = symbol + mean
      if (owner) {
        const float *mine = gp + rank * ld4;  // chunk(2, 1): scales | means
        const size_t o = ((static_cast<size_t>(my_b) * p.H + h) * p.W + w) * p.M;
        float *yo = p.y_hat + ((static_cast<size_t>(my_b) * Hp + h + p.pad) * Wp + w + p.pad) * p.M;
        if (!kDecode) {
          for (int c = tid; c < p.M; c += kArThreads) {
            const float mean = mine[p.M + c];
            const int32_t k = ar_index_of(mine[c], p.bound, s_tab, p.T);
            const int32_t s = ar_quant_sym(p.y[o + c], mean);
            p.idx[o + c] = k;
            p.sym[o + c] = s;
            yo[c] = __fadd_rn(static_cast<float>(s), mean);
          }
        } else {
          for (int c = tid; c < p.M; c += kArThreads) s_idx[c] = ar_index_of(mine[c], p.bound, s_tab, p.T);
          __syncthreads();
          if (warp == 0) {
            RowMeta m_next = meta[s_idx[0]];
            for (int c = 0; c < p.M; ++c) {
              const RowMeta m = m_next;
              if (c + 1 < p.M) m_next = meta[s_idx[c + 1]];  // off the chain: next symbol's row while this one decodes
              const int32_t s = ar_decode_symbol(dec, m, cdf, lut, lut_shift);
              if (lane == 0) s_sym[c] = s;
            }
          }
          __syncthreads();
          for (int c = tid; c < p.M; c += kArThreads) {
            const int32_t s = s_sym[c];
            yo[c] = __fadd_rn(static_cast<float>(s), mine[p.M + c]);
            if (p.sym) p.sym[o + c] = s;
          }
        }
      }
      cluster.sync();  // y_hat(h, w) is visible (L2) to every CTA of the cluster
    }
  }
  if (kDecode && owner && warp == 0 && lane == 0) p.status[my_b] = dec.truncated ? CAI_S_TRUNCATED : CAI_S_OK;
}

template <int G, bool kDecode>
static int launch_ar(const ArKernelParams &p, size_t smem, cudaStream_t stream, const DeviceProps &dp) {
  auto kern = ar_scan_kernel<G, kDecode>;
  int lim = 0;
  int rc = optin_max_smem(reinterpret_cast<const void *>(kern), dp, &lim);
  if (rc != CAI_OK) return rc;
  CAI_CHECK_ARG(static_cast<int>(smem) <= lim, "ar scan: %zu bytes of shared memory needed, %d available", smem, lim);
  const int groups = (p.B + G - 1) / G;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(groups * p.R));
  cfg.blockDim = dim3(kArThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(p.R);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CAI_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  return CAI_OK;
}

static size_t ar_smem_bytes(const ArKernelParams &p, int G, bool decode) {
  size_t fl = static_cast<size_t>(G) * (p.K0 + p.K1p + p.K2p + p.K3p + pad4(p.n3)) + pad4(p.T) + 2 * pad4(p.M);
  size_t bytes = fl * 4;
  if (decode) bytes += (p.stage_bytes + 15u) & ~15u;
  return bytes;
}

static int ar_run(const cai_ar_desc *d, bool decode, cai_table_t t, const float *y, float *y_hat, int32_t *sym,
                  int32_t *idx, const uint32_t *words, const int64_t *word_begin, int32_t *status,
                  cudaStream_t stream) {
  CAI_CHECK_ARG(d != nullptr, "ar scan: null descriptor");
  CAI_CHECK_ARG(d->B >= 0 && d->H > 0 && d->W > 0 && d->M > 0, "ar scan: bad shape");
  CAI_CHECK_ARG(d->M % 4 == 0, "ar scan: the latent channel count must be a multiple of 4 (got %d)", d->M);
  CAI_CHECK_ARG(d->ksize >= 3 && (d->ksize & 1), "ar scan: odd kernel size >= 3 expected");
  CAI_CHECK_ARG(d->n3 == 2 * d->M, "ar scan: entropy_parameters must end with 2 * M channels (scales | means)");
  CAI_CHECK_ARG(d->T >= 1 && d->T <= 4096, "ar scan: scale table size");
  CAI_CHECK_ARG(d->w_ctx && d->b_ctx && d->w1 && d->b1 && d->w2 && d->b2 && d->w3 && d->b3 && d->params &&
                    d->scale_table && y_hat,
                "ar scan: null pointer");
  if (d->B == 0) return CAI_OK;
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  ArKernelParams p = {};
  p.w_ctx = d->w_ctx, p.b_ctx = d->b_ctx, p.w1 = d->w1, p.b1 = d->b1, p.w2 = d->w2, p.b2 = d->b2, p.w3 = d->w3,
  p.b3 = d->b3;
  p.params = d->params, p.y = y, p.y_hat = y_hat, p.sym = sym, p.idx = idx, p.scale_table = d->scale_table;
  p.words = words, p.word_begin = word_begin, p.status = status;
  p.bound = d->scale_bound, p.slope = d->slope;
  p.T = d->T, p.B = d->B, p.H = d->H, p.W = d->W, p.M = d->M, p.P = d->P, p.n_ctx = d->n_ctx, p.n1 = d->n1,
  p.n2 = d->n2, p.n3 = d->n3, p.ksize = d->ksize, p.pad = d->ksize / 2;
  const int ntaps = (d->ksize / 2) * d->ksize + d->ksize / 2;
  p.K0 = ntaps * d->M;
  p.K1p = pad4(d->P + d->n_ctx), p.K2p = pad4(d->n1), p.K3p = pad4(d->n2);
  if (decode) {
    CAI_CHECK_ARG(t && t->blob && words && word_begin && status, "ar decode: null pointer");
    CAI_CHECK_ARG(t->device == dp.device, "ar decode: table lives on device %d, current device is %d", t->device,
                  dp.device);
    p.blob = t->blob, p.enc_bytes = (t->enc_bytes + 15u) & ~15u;
    CAI_CHECK_ARG(p.enc_bytes <= static_cast<uint32_t>(t->blob_bytes), "ar decode: table blob too small");
    p.stage_bytes = p.enc_bytes;
  } else {
    CAI_CHECK_ARG(y && sym && idx, "ar encode: null pointer");
  }
  // Launch shape.  Measured on B200 (tools/ar_bench.py, M = 192, 32 x 48 latent pixels): one SM streams the 7.6 MB of
  // weights from L2 at ~77 GB/s, so a pixel step costs ~95 us / R plus the serial part (barriers; decoder: the M-symbol
  // rANS chain, ~50 us); clusters of all images read the same weight lines at the same time, and beyond ~64-128 CTAs
  // that contention eats the gain of a wider split (B = 16: R = 4 beats R = 8; B = 64: R = 2 beats R = 1 and G = R = 8).
  // Hence: small batches get the widest clusters, larger ones narrower clusters until B x R fills the GPU once; beyond
  // one wave, images share a cluster (G = R) so that one pass over the weights serves G images.
  int R, G;
  if (d->cluster > 0) {
    R = d->cluster;
    G = d->group > 0 ? d->group : 1;
    const int clusters_resident = dp.sm_count / R > 0 ? dp.sm_count / R : 1;
    if (d->group <= 0)
      while (G < R && (p.B + G - 1) / G > clusters_resident) G *= 2;
  } else {
    if (p.B <= 8) R = 8, G = 1;
    else if (p.B <= dp.sm_count / 4) R = 4, G = 1;
    else if (p.B <= dp.sm_count / 2) R = 2, G = 1;
    else if (p.B <= dp.sm_count) R = 2, G = 2;
    else R = 4, G = 4;
    if (d->group > 0) G = d->group;
  }
  CAI_CHECK_ARG(R == 1 || R == 2 || R == 4 || R == 8, "ar scan: cluster size must be 1, 2, 4 or 8");
  p.R = R;
  CAI_CHECK_ARG(G == 1 || G == 2 || G == 4 || G == 8, "ar scan: images per cluster must be 1, 2, 4 or 8");
  CAI_CHECK_ARG(G <= R, "ar scan: images per cluster (%d) cannot exceed the cluster size (%d)", G, R);
  while (G > 1 && ar_smem_bytes(p, G, decode) > static_cast<size_t>(dp.max_smem_optin) - 1024) G /= 2;
  if (decode && (t->blob_bytes & 15) == 0 && (d->flags & 1)) {
    // On request, stage the decode LUT too when the whole blob fits beside the G activation sets.  Off by default:
    // measured on B200 the 187 KB blob leaves ~50 KB of L1 and the weight streaming of the four layers (the bound of
    // the pixel step) slows down by more than the LUT saves on the chain (B = 16, R = 4: decode scan 125 -> 174 ms).
    ArKernelParams q = p;
    q.stage_bytes = static_cast<uint32_t>(t->blob_bytes);
    if (q.stage_bytes > p.enc_bytes && ar_smem_bytes(q, G, true) <= static_cast<size_t>(dp.max_smem_optin) - 1024)
      p.stage_bytes = q.stage_bytes;
  }
  const size_t smem = ar_smem_bytes(p, G, decode);
#define CAI_AR_LAUNCH(GG)                                                   \
  case GG:                                                                  \
    return decode ? launch_ar<GG, true>(p, smem, stream, dp) : launch_ar<GG, false>(p, smem, stream, dp);
  switch (G) {
    CAI_AR_LAUNCH(1)
    CAI_AR_LAUNCH(2)
    CAI_AR_LAUNCH(4)
    CAI_AR_LAUNCH(8)
  }
#undef CAI_AR_LAUNCH
  return CAI_E_INVALID;
}

}  // namespace cai

extern "C" {

__attribute__((visibility("default"))) int cai_ar_encode(const cai_ar_desc *d, const float *y, float *y_hat,
                                                         int32_t *sym, int32_t *idx, cai_stream_t stream) {
  return cai::ar_run(d, false, nullptr, y, y_hat, sym, idx, nullptr, nullptr, nullptr,
                     static_cast<cudaStream_t>(stream));
}

__attribute__((visibility("default"))) int cai_ar_decode(const cai_ar_desc *d, cai_table_t t, const uint32_t *words,
                                                         const int64_t *word_begin, float *y_hat, int32_t *sym,
                                                         int32_t *status, cai_stream_t stream) {
  return cai::ar_run(d, true, t, nullptr, y_hat, sym, nullptr, words, word_begin, status,
                     static_cast<cudaStream_t>(stream));
}

}  // extern "C"
