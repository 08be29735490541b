// quantize.cu -- fused quantise / index / dequantise kernels (HBM-bound, single pass).
//
// What they replace (reference tree, compressai/entropy_models/entropy_models.py):
//   :155-180  EntropyModel.quantize(mode="symbols")   round-half-even(x - mean) -> int32
//   :188-197  EntropyModel.dequantize                 float(sym) + mean
//   :518-541  EntropyBottleneck._build_indexes / compress front end (channel index, medians)
//   :684-689  GaussianConditional.build_indexes       63 sequential compare+subtract passes -> ONE pass
//
// Layout: inputs are logical (N, C, HW) latents stored NCHW or NHWC (channels-last, what the conv
// kernels produce); integer outputs are always in coder order (N, C, HW) so that string b is a
// contiguous run.  The NHWC variants transpose tiles through shared memory so that both the
// loads (along C) and the stores (along HW) are full 128-byte lines (measured: 3.5-4.0 TB/s, 53-61 % of the
// 6.55 TB/s copy peak; the pure elementwise NCHW variant 4.4-5.3 TB/s).
//
// Algorithmic bytes per element: GC 16 B (20 B with means), EB 12 B, dequantize 8-12 B.
#include "common.cuh"

namespace cai {

constexpr int kTabRep = 32;       // scale table replicated per bank -> conflict-free binary search
constexpr int kMaxRepT = 128;     // replicate when T <= 128 (16 KB); otherwise a single copy
constexpr int kMaxT = 4096;

struct ScaleTab {
  const float *t;  // shared memory
  int T;
  int rep;  // 32 or 1
  int lane;
  // (T-1) - #{ j < T-1 : s <= t[j] }  ==  first j in [0, T-1) with s <= t[j], else T-1.  NaN -> T-1.
  __device__ __forceinline__ int32_t index_of(float s, float bound) const {
    s = (s < bound) ? bound : s;  // torch.max(x, bound): NaN propagates (LowerBound, bound_ops.py:36-37)
    int lo = 0, hi = T - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s <= t[mid * rep + (rep == 1 ? 0 : lane)])
        hi = mid;
      else
        lo = mid + 1;
    }
    return lo;
  }
};

__device__ __forceinline__ void load_table(float *s_tab, const float *__restrict__ table, int T, int rep) {
  for (int i = threadIdx.x; i < T * rep; i += blockDim.x) s_tab[i] = table[i / rep];
  __syncthreads();
}

__device__ __forceinline__ int32_t quant_sym(float y, float mean) {
  // clone(); sub_(means); round(); .int()   (entropy_models.py:167-180).  rintf = round-half-even.
  const float r = rintf(__fsub_rn(y, mean));
  // float -> int32 as the x86 reference does (cvttss2si): out-of-range / NaN -> INT_MIN
  if (!(r >= -2147483648.0f && r < 2147483648.0f)) return static_cast<int32_t>(0x80000000u);
  return static_cast<int32_t>(r);
}

// ---- NCHW (pure elementwise, 4 elements per thread) ---------------------------------------------------
__global__ void __launch_bounds__(256)
gc_qi_nchw_kernel(const float *__restrict__ y, const float *__restrict__ scales, const float *__restrict__ means,
                  const float *__restrict__ table, int T, float bound, int64_t total, int vec,
                  int32_t *__restrict__ sym, int32_t *__restrict__ idx) {
  extern __shared__ float s_tab[];
  const int rep = (T <= kMaxRepT) ? kTabRep : 1;
  if (scales) load_table(s_tab, table, T, rep);
  ScaleTab st{s_tab, T, rep, static_cast<int>(threadIdx.x & 31)};
  const int64_t nvec = vec ? (total >> 2) : 0;  // vector path needs 16-byte aligned bases
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    if (y) {
      const float4 a = __ldcs(reinterpret_cast<const float4 *>(y) + v);
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
      if (means) m = __ldcs(reinterpret_cast<const float4 *>(means) + v);
      int4 o;
      o.x = quant_sym(a.x, m.x);
      o.y = quant_sym(a.y, m.y);
      o.z = quant_sym(a.z, m.z);
      o.w = quant_sym(a.w, m.w);
      reinterpret_cast<int4 *>(sym)[v] = o;
    }
    if (scales) {
      const float4 s = __ldcs(reinterpret_cast<const float4 *>(scales) + v);
      int4 o;
      o.x = st.index_of(s.x, bound);
      o.y = st.index_of(s.y, bound);
      o.z = st.index_of(s.z, bound);
      o.w = st.index_of(s.w, bound);
      reinterpret_cast<int4 *>(idx)[v] = o;
    }
  }
  // tail (total % 4)
  const int64_t t0 = nvec << 2;
  for (int64_t i = t0 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    if (y) sym[i] = quant_sym(y[i], means ? means[i] : 0.f);
    if (scales) idx[i] = st.index_of(scales[i], bound);
  }
}

// ---- NHWC -> coder order (tile transpose) -------------------------------------------------------------
// Tile = 32 pixels (hw) x 64 channels.  Loads: float4 along the channel axis (16 threads cover the 256
// contiguous bytes of one pixel's 64 channels; 256 threads = 16 pixels per pass, 2 passes).  Stores: int4 along
// the hw axis (8 threads cover the 128 contiguous bytes of one channel's 32 pixels; 32 channels per pass, 2 passes).
// Both directions move full 128-byte lines; the int32 tiles are transposed through padded shared memory.
// grid: x = hw tiles, y = c tiles, z = n ; block 256
constexpr int kTH = 32, kTC = 64;

__global__ void __launch_bounds__(256)
gc_qi_nhwc_kernel(const float *__restrict__ y, const float *__restrict__ scales, const float *__restrict__ means,
                  const float *__restrict__ table, int T, float bound, int64_t C, int64_t HW,
                  int32_t *__restrict__ sym, int32_t *__restrict__ idx) {
  extern __shared__ float s_tab[];
  __shared__ int32_t t_sym[kTC][kTH + 1];
  __shared__ int32_t t_idx[kTC][kTH + 1];
  const int rep = (T <= kMaxRepT) ? kTabRep : 1;
  if (scales) load_table(s_tab, table, T, rep);
  ScaleTab st{s_tab, T, rep, static_cast<int>(threadIdx.x & 31)};
  const int64_t n = blockIdx.z;
  const int64_t hw0 = static_cast<int64_t>(blockIdx.x) * kTH, c0 = static_cast<int64_t>(blockIdx.y) * kTC;
  const int tid = threadIdx.x;
  const bool vec_ok = (C & 3) == 0;  // float4 loads need C % 4 == 0 (bases are 16-byte aligned by the host check)
  {
    const int cq = tid & 15, hr = tid >> 4;  // 16 channel quads x 16 pixels per pass
#pragma unroll
    for (int pss = 0; pss < kTH / 16; ++pss) {
      const int hl = hr + 16 * pss;
      const int64_t hw = hw0 + hl, c = c0 + 4 * cq;
      if (hw >= HW || c >= C) continue;
      const int64_t base = (n * HW + hw) * C + c;
      float yv[4] = {0.f, 0.f, 0.f, 0.f}, mv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec_ok && c + 3 < C) {
        if (y) {
          const float4 a = __ldcs(reinterpret_cast<const float4 *>(y + base));
          yv[0] = a.x; yv[1] = a.y; yv[2] = a.z; yv[3] = a.w;
          if (means) {
            const float4 m = __ldcs(reinterpret_cast<const float4 *>(means + base));
            mv[0] = m.x; mv[1] = m.y; mv[2] = m.z; mv[3] = m.w;
          }
        }
        if (scales) {
          const float4 q = __ldcs(reinterpret_cast<const float4 *>(scales + base));
          sv[0] = q.x; sv[1] = q.y; sv[2] = q.z; sv[3] = q.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c + j < C) {
            if (y) yv[j] = __ldcs(y + base + j);
            if (y && means) mv[j] = __ldcs(means + base + j);
            if (scales) sv[j] = __ldcs(scales + base + j);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (y) t_sym[4 * cq + j][hl] = quant_sym(yv[j], mv[j]);
        if (scales) t_idx[4 * cq + j][hl] = st.index_of(sv[j], bound);
      }
    }
  }
  __syncthreads();
  {
    const int hq = tid & 7, cr = tid >> 3;  // 8 pixel quads x 32 channels per pass
    const bool out_vec = (HW & 3) == 0;
#pragma unroll
    for (int pss = 0; pss < kTC / 32; ++pss) {
      const int cl = cr + 32 * pss;
      const int64_t c = c0 + cl, hw = hw0 + 4 * hq;
      if (c >= C || hw >= HW) continue;
      const int64_t o = (n * C + c) * HW + hw;
      if (out_vec && hw + 3 < HW) {
        if (y) *reinterpret_cast<int4 *>(sym + o) = make_int4(t_sym[cl][4 * hq], t_sym[cl][4 * hq + 1], t_sym[cl][4 * hq + 2], t_sym[cl][4 * hq + 3]);
        if (scales) *reinterpret_cast<int4 *>(idx + o) = make_int4(t_idx[cl][4 * hq], t_idx[cl][4 * hq + 1], t_idx[cl][4 * hq + 2], t_idx[cl][4 * hq + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (hw + j < HW) {
            if (y) sym[o + j] = t_sym[cl][4 * hq + j];
            if (scales) idx[o + j] = t_idx[cl][4 * hq + j];
          }
        }
      }
    }
  }
}

// ---- entropy bottleneck front end ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eb_qi_nchw_kernel(const float *__restrict__ x, const float *__restrict__ medians, int64_t C, int64_t HW,
                  int64_t total, int32_t *__restrict__ sym, int32_t *__restrict__ idx) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int32_t c = static_cast<int32_t>((i / HW) % C);
    if (sym) sym[i] = quant_sym(__ldcs(x + i), __ldg(medians + c));
    if (idx) idx[i] = c;
  }
}

__global__ void __launch_bounds__(256)
eb_qi_nhwc_kernel(const float *__restrict__ x, const float *__restrict__ medians, int64_t C, int64_t HW,
                  int32_t *__restrict__ sym, int32_t *__restrict__ idx) {
  __shared__ int32_t t_sym[32][33];
  const int64_t n = blockIdx.z;
  const int64_t hw0 = static_cast<int64_t>(blockIdx.x) * 32, c0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (sym) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int64_t hw = hw0 + ty + 8 * r, c = c0 + tx;
      if (hw < HW && c < C) t_sym[ty + 8 * r][tx] = quant_sym(__ldcs(x + (n * HW + hw) * C + c), __ldg(medians + c));
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t c = c0 + ty + 8 * r, hw = hw0 + tx;
    if (hw < HW && c < C) {
      const int64_t o = (n * C + c) * HW + hw;
      if (sym) sym[o] = t_sym[tx][ty + 8 * r];
      if (idx) idx[o] = static_cast<int32_t>(c);
    }
  }
}

// ---- dequantise ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
deq_nchw_kernel(const int32_t *__restrict__ sym, const float *__restrict__ means, const float *__restrict__ medians,
                int64_t C, int64_t HW, int64_t total, float *__restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    float v = static_cast<float>(__ldcs(sym + i));
    if (means) v = __fadd_rn(v, __ldcs(means + i));
    if (medians) v = __fadd_rn(v, __ldg(medians + (i / HW) % C));
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256)
deq_nhwc_kernel(const int32_t *__restrict__ sym, const float *__restrict__ means, const float *__restrict__ medians,
                int64_t C, int64_t HW, float *__restrict__ out) {
  __shared__ int32_t t_sym[32][33];
  const int64_t n = blockIdx.z;
  const int64_t hw0 = static_cast<int64_t>(blockIdx.x) * 32, c0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t c = c0 + ty + 8 * r, hw = hw0 + tx;
    if (hw < HW && c < C) t_sym[ty + 8 * r][tx] = __ldcs(sym + (n * C + c) * HW + hw);  // [c_local][hw_local]
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t hw = hw0 + ty + 8 * r, c = c0 + tx;
    if (hw < HW && c < C) {
      const int64_t o = (n * HW + hw) * C + c;
      float v = static_cast<float>(t_sym[tx][ty + 8 * r]);
      if (means) v = __fadd_rn(v, __ldcs(means + o));
      if (medians) v = __fadd_rn(v, __ldg(medians + c));
      out[o] = v;
    }
  }
}

static int elementwise_grid(const DeviceProps &dp, int64_t work_items, int threads) {
  int64_t g = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(dp.sm_count) * 16;  // multiple of the SM count, grid-stride beyond
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}


// 8-bit pixels <-> unit-range floats (the ToTensor() convention of the reference's image pipeline,
// examples/codec.py:112-128 / utils/eval_model/__main__.py:70-74: x = u8 / 255).  Shipping uint8 over PCIe
// and converting on the device moves a quarter of the bytes of the fp32 tensor the reference's API takes.
__global__ void u8_to_unit_kernel(const uint8_t *__restrict__ in, int64_t n, float *__restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n16 = n >> 4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4 *o = reinterpret_cast<float4 *>(out) + 4 * i;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      o[q] = make_float4(static_cast<float>(w[q] & 0xFFu) / 255.f, static_cast<float>((w[q] >> 8) & 0xFFu) / 255.f,
                         static_cast<float>((w[q] >> 16) & 0xFFu) / 255.f, static_cast<float>(w[q] >> 24) / 255.f);
  }
  for (int64_t i = (n16 << 4) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = static_cast<float>(in[i]) / 255.f;
}

__device__ __forceinline__ uint32_t unit_to_u8(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);                      // NaN -> 0 (fmaxf returns the non-NaN operand)
  return static_cast<uint32_t>(__float2int_rn(v * 255.f));  // round half to even, like torch.round
}

__global__ void unit_to_u8_kernel(const float *__restrict__ in, int64_t n, uint8_t *__restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n16 = n >> 4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 f = __ldg(reinterpret_cast<const float4 *>(in) + 4 * i + q);
      w[q] = unit_to_u8(f.x) | (unit_to_u8(f.y) << 8) | (unit_to_u8(f.z) << 16) | (unit_to_u8(f.w) << 24);
    }
    reinterpret_cast<uint4 *>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int64_t i = (n16 << 4) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = static_cast<uint8_t>(unit_to_u8(in[i]));
}

}  // namespace cai

using namespace cai;

extern "C" {

int cai_gc_quantize_index(const float *y, const float *scales, const float *means, const float *scale_table,
                          int32_t T, float scale_bound, int32_t layout, int64_t N, int64_t C, int64_t HW,
                          int32_t *sym, int32_t *idx, cai_stream_t stream_) {
  CAI_CHECK_ARG(N >= 0 && C >= 0 && HW >= 0, "cai_gc_quantize_index: negative size");
  CAI_CHECK_ARG(layout == CAI_LAYOUT_NCHW || layout == CAI_LAYOUT_NHWC, "cai_gc_quantize_index: bad layout");
  CAI_CHECK_ARG(!y || sym, "cai_gc_quantize_index: y given without sym output");
  CAI_CHECK_ARG(!scales || (idx && scale_table && T >= 1 && T <= kMaxT),
                "cai_gc_quantize_index: scales need idx, scale_table and 1 <= T <= %d", kMaxT);
  CAI_CHECK_ARG(!means || y, "cai_gc_quantize_index: means given without y");
  const int64_t total = N * C * HW;
  if (total == 0 || (!y && !scales)) return CAI_OK;
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rep = (T <= kMaxRepT) ? kTabRep : 1;
  const size_t smem = scales ? sizeof(float) * static_cast<size_t>(T) * rep : 0;
  if (layout == CAI_LAYOUT_NCHW || C == 1 || HW == 1) {
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(scales) |
                           reinterpret_cast<uintptr_t>(means) | reinterpret_cast<uintptr_t>(sym) |
                           reinterpret_cast<uintptr_t>(idx)) & 15u) == 0;
    const int grid = elementwise_grid(dp, aligned ? (total + 3) / 4 : total, 256);
    gc_qi_nchw_kernel<<<grid, 256, smem, stream>>>(y, scales, means, scale_table, T, scale_bound, total,
                                                   aligned ? 1 : 0, sym, idx);
  } else {
    CAI_CHECK_ARG(N <= 65535 && (C + kTC - 1) / kTC <= 65535, "cai_gc_quantize_index: N or C too large for the grid");
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(scales) |
                           reinterpret_cast<uintptr_t>(means) | reinterpret_cast<uintptr_t>(sym) |
                           reinterpret_cast<uintptr_t>(idx)) & 15u) == 0;
    CAI_CHECK_ARG(aligned, "cai_gc_quantize_index: NHWC tensors must be 16-byte aligned");
    dim3 grid(static_cast<unsigned>((HW + kTH - 1) / kTH), static_cast<unsigned>((C + kTC - 1) / kTC),
              static_cast<unsigned>(N));
    gc_qi_nhwc_kernel<<<grid, 256, smem, stream>>>(y, scales, means, scale_table, T, scale_bound, C, HW, sym, idx);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_eb_quantize_index(const float *x, const float *medians, int32_t layout, int64_t N, int64_t C, int64_t HW,
                          int32_t *sym, int32_t *idx, cai_stream_t stream_) {
  CAI_CHECK_ARG(N >= 0 && C >= 0 && HW >= 0, "cai_eb_quantize_index: negative size");
  CAI_CHECK_ARG(layout == CAI_LAYOUT_NCHW || layout == CAI_LAYOUT_NHWC, "cai_eb_quantize_index: bad layout");
  CAI_CHECK_ARG(!sym || (x && medians), "cai_eb_quantize_index: sym needs x and medians");
  const int64_t total = N * C * HW;
  if (total == 0 || (!sym && !idx)) return CAI_OK;
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (layout == CAI_LAYOUT_NCHW || C == 1 || HW == 1) {
    eb_qi_nchw_kernel<<<elementwise_grid(dp, total, 256), 256, 0, stream>>>(x, medians, C, HW, total, sym, idx);
  } else {
    CAI_CHECK_ARG(N <= 65535 && (C + 31) / 32 <= 65535, "cai_eb_quantize_index: N or C too large for the grid");
    dim3 grid(static_cast<unsigned>((HW + 31) / 32), static_cast<unsigned>((C + 31) / 32), static_cast<unsigned>(N));
    eb_qi_nhwc_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, medians, C, HW, sym, idx);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_dequantize(const int32_t *sym, const float *means, const float *medians, int32_t layout, int64_t N,
                   int64_t C, int64_t HW, float *out, cai_stream_t stream_) {
  CAI_CHECK_ARG(N >= 0 && C >= 0 && HW >= 0, "cai_dequantize: negative size");
  CAI_CHECK_ARG(layout == CAI_LAYOUT_NCHW || layout == CAI_LAYOUT_NHWC, "cai_dequantize: bad layout");
  CAI_CHECK_ARG(!(means && medians), "cai_dequantize: give means or medians, not both");
  const int64_t total = N * C * HW;
  if (total == 0) return CAI_OK;
  CAI_CHECK_ARG(sym && out, "cai_dequantize: NULL pointer");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (layout == CAI_LAYOUT_NCHW || C == 1 || HW == 1) {
    deq_nchw_kernel<<<elementwise_grid(dp, total, 256), 256, 0, stream>>>(sym, means, medians, C, HW, total, out);
  } else {
    CAI_CHECK_ARG(N <= 65535 && (C + 31) / 32 <= 65535, "cai_dequantize: N or C too large for the grid");
    dim3 grid(static_cast<unsigned>((HW + 31) / 32), static_cast<unsigned>((C + 31) / 32), static_cast<unsigned>(N));
    deq_nhwc_kernel<<<grid, dim3(32, 8), 0, stream>>>(sym, means, medians, C, HW, out);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_pixels_u8_to_f32(const uint8_t *in, int64_t n, float *out, cai_stream_t stream_) {
  CAI_CHECK_ARG(n >= 0, "cai_pixels_u8_to_f32: n < 0");
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(in && out, "cai_pixels_u8_to_f32: NULL pointer");
  CAI_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                "cai_pixels_u8_to_f32: buffers must be 16-byte aligned");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  u8_to_unit_kernel<<<elementwise_grid(dp, (n + 15) / 16, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(in, n, out);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_pixels_f32_to_u8(const float *in, int64_t n, uint8_t *out, cai_stream_t stream_) {
  CAI_CHECK_ARG(n >= 0, "cai_pixels_f32_to_u8: n < 0");
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(in && out, "cai_pixels_f32_to_u8: NULL pointer");
  CAI_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                "cai_pixels_f32_to_u8: buffers must be 16-byte aligned");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  unit_to_u8_kernel<<<elementwise_grid(dp, (n + 15) / 16, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(in, n, out);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

}  // extern "C"
