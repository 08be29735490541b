// likelihood.cu -- fused likelihood kernels (forward + backward) for the two entropy models.
//
// What they replace (reference tree, compressai/entropy_models/entropy_models.py):
//   :650-682  GaussianConditional._likelihood / forward  (~12 elementwise ATen kernels)  -> 1 kernel
//   :436-469  EntropyBottleneck._logits_cumulative / _likelihood (5 bmm + ~20 elementwise) -> 1 kernel
//   :471-516  EntropyBottleneck.forward (permute / quantize / likelihood / bound / permute back)
//   :431-434  EntropyBottleneck.loss (logits at the quantiles)
//   compressai/ops/bound_ops.py:36-42  LowerBound forward / custom backward (folded in)
// Backward formulas: SURVEY.md Appendix D.2 / D.3 (checked there against the reference autograd).
//
// All kernels are elementwise over HBM (roofline: HBM bandwidth; algorithmic bytes per element are in
// DESIGN.md).  EB parameters are passed already transformed ("tparams": softplus(matrix), bias,
// tanh(factor) packed per channel); the tiny C x P transform and its chain rule stay in the host layer.
#include "common.cuh"

namespace cai {

constexpr int kEbMaxLayers = 8;   // len(filters) + 1
constexpr int kEbMaxP = 64;       // packed parameters per channel supported by the backward kernel

struct EbSpec {
  int nlayers;                     // len(filters) + 1
  int width[kEbMaxLayers + 1];     // (1, f1, ..., fk, 1)
  int P;                           // packed floats per channel
  int stride;                      // smem stride per channel (odd -> conflict free)
};

__device__ __forceinline__ float phi_cdf(float t) {  // _standardized_cumulative (entropy_models.py:604-608)
  return 0.5f * erfcf(-0.70710678118654752440f * t);
}
__device__ __forceinline__ float phi_pdf(float t) { return 0.39894228040143267794f * expf(-0.5f * t * t); }
__device__ __forceinline__ float sigmoidf_(float a) { return 1.f / (1.f + expf(-a)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float signf_(float a) { return (a > 0.f) ? 1.f : ((a < 0.f) ? -1.f : 0.f); }

// ---- Gaussian conditional ------------------------------------------------------------------------------
// mode 0: y_hat = y + noise (training, means ignored by quantize: entropy_models.py:161-165)
// mode 1: y_hat = rint(y - mu) + mu (eval)            mode 2: y_hat = y (likelihood of given values)
__global__ void __launch_bounds__(256)
gc_forward_kernel(const float *__restrict__ y, const float *__restrict__ scales, const float *__restrict__ means,
                  const float *__restrict__ noise, int mode, float bound_scale, float bound_lik, int64_t n,
                  float *__restrict__ y_hat, float *__restrict__ lik) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float mu = means ? __ldcs(means + i) : 0.f;
    float v = __ldcs(y + i);
    if (mode == 0)
      v = __fadd_rn(v, __ldcs(noise + i));
    else if (mode == 1)
      v = __fadd_rn(rintf(__fsub_rn(v, mu)), mu);
    if (y_hat) y_hat[i] = v;
    if (lik) {
      const float a = fabsf(__fsub_rn(v, mu));
      float s = __ldcs(scales + i);
      s = (s < bound_scale) ? bound_scale : s;
      const float u = __fdiv_rn(0.5f - a, s), l = __fdiv_rn(-0.5f - a, s);
      float L = phi_cdf(u) - phi_cdf(l);
      if (bound_lik > 0.f) L = (L < bound_lik) ? bound_lik : L;
      lik[i] = L;
    }
  }
}

__global__ void __launch_bounds__(256)
gc_backward_kernel(const float *__restrict__ y_hat, const float *__restrict__ scales, const float *__restrict__ means,
                   const float *__restrict__ g_lik, float bound_scale, float bound_lik, int64_t n,
                   float *__restrict__ g_y, float *__restrict__ g_scales, float *__restrict__ g_means) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float mu = means ? __ldcs(means + i) : 0.f;
    const float d = __fsub_rn(__ldcs(y_hat + i), mu);
    const float a = fabsf(d);
    const float sig = __ldcs(scales + i);
    const float s = (sig < bound_scale) ? bound_scale : sig;
    const float u = __fdiv_rn(0.5f - a, s), l = __fdiv_rn(-0.5f - a, s);
    const float L = phi_cdf(u) - phi_cdf(l);
    float g = __ldcs(g_lik + i);
    if (bound_lik > 0.f && !(L >= bound_lik || g < 0.f)) g = 0.f;  // LowerBound backward gate
    const float pu = phi_pdf(u), pl = phi_pdf(l);
    const float inv = 1.f / s;
    const float dy = g * signf_(d) * (pl - pu) * inv;
    float ds = g * (l * pl - u * pu) * inv;
    if (!(sig >= bound_scale || ds < 0.f)) ds = 0.f;
    if (g_y) g_y[i] = dy;
    if (g_means) g_means[i] = -dy;
    if (g_scales) g_scales[i] = ds;
  }
}

// ---- entropy bottleneck ----------------------------------------------------------------------------------
// Per channel packed layout: for layer i: W_i (f_{i+1} x f_i, row major), b_i (f_{i+1}), a_i (f_{i+1}, i < last).
template <int MAXF>
struct EbNet {
  const float *p;  // this channel's transformed params (shared or global)
  const EbSpec &sp;
  __device__ __forceinline__ EbNet(const float *p_, const EbSpec &s_) : p(p_), sp(s_) {}

  // forward; optionally records h_i (input of layer i) and t_i = tanh(z_i) for the backward pass
  template <bool kSave>
  __device__ __forceinline__ float fwd(float x, float (*hs)[MAXF], float (*ts)[MAXF]) const {
    float h[MAXF];
    h[0] = x;
    const float *q = p;
#pragma unroll 1
    for (int i = 0; i < sp.nlayers; ++i) {
      const int fi = sp.width[i], fo = sp.width[i + 1];
      const bool last = (i == sp.nlayers - 1);
      float z[MAXF];
      if (kSave) {
#pragma unroll
        for (int c = 0; c < MAXF; ++c) hs[i][c] = (c < fi) ? h[c] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < MAXF; ++r) {
        if (r < fo) {
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < MAXF; ++c)
            if (c < fi) acc = fmaf(q[r * fi + c], h[c], acc);
          z[r] = acc + q[fo * fi + r];
        }
      }
#pragma unroll
      for (int r = 0; r < MAXF; ++r) {
        if (r < fo) {
          if (!last) {
            const float t = tanhf(z[r]);
            if (kSave) ts[i][r] = t;
            h[r] = fmaf(q[fo * fi + fo + r], t, z[r]);
          } else {
            h[r] = z[r];
          }
        }
      }
      q += fo * fi + fo + (last ? 0 : fo);
    }
    return h[0];
  }

  // backward of one path: upstream g on the scalar output; accumulates parameter grads into gp[] (if non-null)
  // and returns d(out)/dx * g
  __device__ __forceinline__ float bwd(float g, float (*hs)[MAXF], float (*ts)[MAXF], float *gp) const {
    // offsets of each layer
    int off[kEbMaxLayers];
    int o = 0;
    for (int i = 0; i < sp.nlayers; ++i) {
      off[i] = o;
      const int fi = sp.width[i], fo = sp.width[i + 1];
      o += fo * fi + fo + ((i == sp.nlayers - 1) ? 0 : fo);
    }
    float gh[MAXF];
#pragma unroll
    for (int c = 0; c < MAXF; ++c) gh[c] = 0.f;
    gh[0] = g;
#pragma unroll 1
    for (int i = sp.nlayers - 1; i >= 0; --i) {
      const int fi = sp.width[i], fo = sp.width[i + 1];
      const bool last = (i == sp.nlayers - 1);
      const float *q = p + off[i];
      float gz[MAXF];
#pragma unroll
      for (int r = 0; r < MAXF; ++r) {
        gz[r] = 0.f;
        if (r < fo) {
          if (!last) {
            const float t = ts[i][r];
            const float a = q[fo * fi + fo + r];
            if (gp) gp[off[i] + fo * fi + fo + r] += gh[r] * t;  // d/da
            gz[r] = gh[r] * fmaf(a, 1.f - t * t, 1.f);
          } else {
            gz[r] = gh[r];
          }
          if (gp) gp[off[i] + fo * fi + r] += gz[r];  // d/db
        }
      }
      float nh[MAXF];
#pragma unroll
      for (int c = 0; c < MAXF; ++c) {
        nh[c] = 0.f;
        if (c < fi) {
          float acc = 0.f;
#pragma unroll
          for (int r = 0; r < MAXF; ++r) {
            if (r < fo) {
              acc = fmaf(q[r * fi + c], gz[r], acc);
              if (gp) gp[off[i] + r * fi + c] += gz[r] * hs[i][c];  // d/dW
            }
          }
          nh[c] = acc;
        }
      }
#pragma unroll
      for (int c = 0; c < MAXF; ++c) gh[c] = nh[c];
    }
    return gh[0];
  }
};

__device__ __forceinline__ void eb_stage_params(float *s_par, const float *__restrict__ tparams, int C, const EbSpec &sp) {
  for (int i = threadIdx.x; i < C * sp.P; i += blockDim.x) {
    const int c = i / sp.P, k = i - c * sp.P;
    s_par[c * sp.stride + k] = tparams[i];
  }
  __syncthreads();
}

__device__ __forceinline__ int eb_channel(int64_t i, int layout, int64_t C, int64_t HW) {
  return (layout == CAI_LAYOUT_NHWC) ? static_cast<int>(i % C) : static_cast<int>((i / HW) % C);
}

// mode 0: x~ = x + noise; mode 1: x~ = rint(x - med) + med; mode 2: x~ = x
template <int MAXF>
__global__ void __launch_bounds__(256)
eb_forward_kernel(const float *__restrict__ x, const float *__restrict__ tparams, const float *__restrict__ medians,
                  const float *__restrict__ noise, EbSpec sp, int mode, float bound_lik, int layout, int64_t C,
                  int64_t HW, int64_t n, float *__restrict__ out, float *__restrict__ lik) {
  extern __shared__ float s_par[];
  eb_stage_params(s_par, tparams, static_cast<int>(C), sp);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int c = eb_channel(i, layout, C, HW);
    float v = __ldcs(x + i);
    if (mode == 0) {
      v = __fadd_rn(v, __ldcs(noise + i));
    } else if (mode == 1) {
      const float m = __ldg(medians + c);
      v = __fadd_rn(rintf(__fsub_rn(v, m)), m);
    }
    if (out) out[i] = v;
    if (lik) {
      EbNet<MAXF> net(s_par + c * sp.stride, sp);
      const float lo = net.template fwd<false>(v - 0.5f, nullptr, nullptr);
      const float up = net.template fwd<false>(v + 0.5f, nullptr, nullptr);
      const float sg = -signf_(lo + up);
      float L = fabsf(sigmoidf_(sg * up) - sigmoidf_(sg * lo));
      if (bound_lik > 0.f) L = (L < bound_lik) ? bound_lik : L;
      lik[i] = L;
    }
  }
}

// grads: g_x (same layout as x) and g_tparams [C, P] (atomically accumulated; caller zero-fills)
template <int MAXF>
__global__ void __launch_bounds__(128)
eb_backward_kernel(const float *__restrict__ xt /* x~ */, const float *__restrict__ tparams,
                   const float *__restrict__ g_lik, EbSpec sp, float bound_lik, int layout, int64_t C, int64_t HW,
                   int64_t N, float *__restrict__ g_x, float *__restrict__ g_tparams) {
  // grid: x = chunks, y = channel.  Every thread of a block works on the same channel.
  const int c = blockIdx.y;
  __shared__ float s_p[kEbMaxP];
  __shared__ float s_acc[kEbMaxP];
  for (int k = threadIdx.x; k < sp.P; k += blockDim.x) {
    s_p[k] = tparams[static_cast<int64_t>(c) * sp.P + k];
    s_acc[k] = 0.f;
  }
  __syncthreads();
  float gp[kEbMaxP];
#pragma unroll
  for (int k = 0; k < kEbMaxP; ++k) gp[k] = 0.f;
  EbNet<MAXF> net(s_p, sp);
  const int64_t per_c = N * HW;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < per_c; e += stride) {
    const int64_t nimg = e / HW, hw = e - nimg * HW;
    const int64_t i = (layout == CAI_LAYOUT_NHWC) ? ((nimg * HW + hw) * C + c) : ((nimg * C + c) * HW + hw);
    const float v = __ldcs(xt + i);
    float hs0[kEbMaxLayers][MAXF], ts0[kEbMaxLayers][MAXF], hs1[kEbMaxLayers][MAXF], ts1[kEbMaxLayers][MAXF];
    const float lo = net.template fwd<true>(v - 0.5f, hs0, ts0);
    const float up = net.template fwd<true>(v + 0.5f, hs1, ts1);
    const float sg = -signf_(lo + up);
    const float su = sigmoidf_(sg * up), sl = sigmoidf_(sg * lo);
    const float D = su - sl;
    const float L = fabsf(D);
    float g = __ldcs(g_lik + i);
    if (bound_lik > 0.f && !(L >= bound_lik || g < 0.f)) g = 0.f;
    const float gD = g * signf_(D);
    const float g_up = gD * sg * su * (1.f - su);
    const float g_lo = -gD * sg * sl * (1.f - sl);
    float gx = net.bwd(g_lo, hs0, ts0, g_tparams ? gp : nullptr);
    gx += net.bwd(g_up, hs1, ts1, g_tparams ? gp : nullptr);
    if (g_x) g_x[i] = gx;
  }
  if (g_tparams) {
#pragma unroll
    for (int k = 0; k < kEbMaxP; ++k) {
      if (k < sp.P) {
        const float v = warp_sum(gp[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[k], v);
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < sp.P; k += blockDim.x) atomicAdd(&g_tparams[static_cast<int64_t>(c) * sp.P + k], s_acc[k]);
  }
}

// logits at arbitrary per-channel sample points: x [C, L] -> out [C, L]; optional backward wrt x
template <int MAXF>
__global__ void __launch_bounds__(256)
eb_logits_kernel(const float *__restrict__ x, const float *__restrict__ tparams, const float *__restrict__ g_out,
                 EbSpec sp, int64_t C, int64_t L, float *__restrict__ out, float *__restrict__ g_x) {
  const int64_t n = C * L;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t c = i / L;
    EbNet<MAXF> net(tparams + c * sp.P, sp);
    if (g_x) {
      float hs[kEbMaxLayers][MAXF], ts[kEbMaxLayers][MAXF];
      const float v = net.template fwd<true>(x[i], hs, ts);
      if (out) out[i] = v;
      g_x[i] = net.bwd(g_out[i], hs, ts, nullptr);
    } else {
      out[i] = net.template fwd<false>(x[i], nullptr, nullptr);
    }
  }
}

static int make_spec(const int32_t *filters_host, int32_t n_filters, EbSpec *sp) {
  if (n_filters < 0 || n_filters + 1 > kEbMaxLayers) return -1;
  sp->nlayers = n_filters + 1;
  sp->width[0] = 1;
  for (int i = 0; i < n_filters; ++i) sp->width[i + 1] = filters_host[i];
  sp->width[n_filters + 1] = 1;
  int P = 0, maxf = 1;
  for (int i = 0; i < sp->nlayers; ++i) {
    const int fi = sp->width[i], fo = sp->width[i + 1];
    if (fo < 1 || fo > 8) return -1;
    maxf = fo > maxf ? fo : maxf;
    P += fo * fi + fo + ((i == sp->nlayers - 1) ? 0 : fo);
  }
  sp->P = P;
  sp->stride = P | 1;
  return maxf;
}

static int ew_grid(const DeviceProps &dp, int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(dp.sm_count) * 16;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace cai

using namespace cai;

extern "C" {

int cai_gc_forward(const float *y, const float *scales, const float *means, const float *noise, int32_t mode,
                   float bound_scale, float bound_lik, int64_t n, float *y_hat, float *lik, cai_stream_t stream_) {
  CAI_CHECK_ARG(n >= 0 && mode >= 0 && mode <= 2, "cai_gc_forward: bad n / mode");
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(y && (y_hat || lik), "cai_gc_forward: NULL input / no output");
  CAI_CHECK_ARG(mode != 0 || noise, "cai_gc_forward: mode 0 needs noise");
  CAI_CHECK_ARG(!lik || scales, "cai_gc_forward: likelihood needs scales");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  gc_forward_kernel<<<ew_grid(dp, n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      y, scales, means, noise, mode, bound_scale, bound_lik, n, y_hat, lik);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_gc_backward(const float *y_hat, const float *scales, const float *means, const float *g_lik,
                    float bound_scale, float bound_lik, int64_t n, float *g_y, float *g_scales, float *g_means,
                    cai_stream_t stream_) {
  CAI_CHECK_ARG(n >= 0, "cai_gc_backward: n < 0");
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(y_hat && scales && g_lik, "cai_gc_backward: NULL input");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  gc_backward_kernel<<<ew_grid(dp, n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      y_hat, scales, means, g_lik, bound_scale, bound_lik, n, g_y, g_scales, g_means);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_eb_forward(const float *x, const float *tparams, const int32_t *filters_host, int32_t n_filters,
                   const float *medians, const float *noise, int32_t mode, float bound_lik, int32_t layout, int64_t N,
                   int64_t C, int64_t HW, float *out, float *lik, cai_stream_t stream_) {
  CAI_CHECK_ARG(N >= 0 && C >= 0 && HW >= 0 && mode >= 0 && mode <= 2, "cai_eb_forward: bad size / mode");
  const int64_t n = N * C * HW;
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(x && tparams && (out || lik), "cai_eb_forward: NULL pointer");
  CAI_CHECK_ARG(mode != 0 || noise, "cai_eb_forward: mode 0 needs noise");
  CAI_CHECK_ARG(mode != 1 || medians, "cai_eb_forward: mode 1 needs medians");
  EbSpec sp;
  const int maxf = make_spec(filters_host, n_filters, &sp);
  CAI_CHECK_ARG(maxf > 0, "cai_eb_forward: unsupported filters (need <= %d layers, widths in [1, 8])", kEbMaxLayers - 1);
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  const size_t smem = sizeof(float) * static_cast<size_t>(C) * sp.stride;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(dp, n, 256);
  const void *fn = (maxf <= 3) ? reinterpret_cast<const void *>(eb_forward_kernel<3>)
                               : reinterpret_cast<const void *>(eb_forward_kernel<8>);
  int max_dyn = 0;
  rc = optin_max_smem(fn, dp, &max_dyn);  // once per device; never per launch (host threads share the attribute)
  if (rc != CAI_OK) return rc;
  CAI_CHECK_ARG(smem <= static_cast<size_t>(max_dyn), "cai_eb_forward: too many channels (%lld)",
                static_cast<long long>(C));
  if (maxf <= 3) {
    eb_forward_kernel<3><<<grid, 256, smem, stream>>>(x, tparams, medians, noise, sp, mode, bound_lik, layout, C, HW, n, out, lik);
  } else {
    eb_forward_kernel<8><<<grid, 256, smem, stream>>>(x, tparams, medians, noise, sp, mode, bound_lik, layout, C, HW, n, out, lik);
  }
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_eb_backward(const float *x_tilde, const float *tparams, const int32_t *filters_host, int32_t n_filters,
                    const float *g_lik, float bound_lik, int32_t layout, int64_t N, int64_t C, int64_t HW, float *g_x,
                    float *g_tparams, cai_stream_t stream_) {
  CAI_CHECK_ARG(N >= 0 && C >= 0 && HW >= 0, "cai_eb_backward: bad size");
  if (N * C * HW == 0) return CAI_OK;
  CAI_CHECK_ARG(x_tilde && tparams && g_lik && (g_x || g_tparams), "cai_eb_backward: NULL pointer");
  EbSpec sp;
  const int maxf = make_spec(filters_host, n_filters, &sp);
  CAI_CHECK_ARG(maxf > 0 && sp.P <= kEbMaxP, "cai_eb_backward: unsupported filters (packed size %d > %d)", sp.P, kEbMaxP);
  CAI_CHECK_ARG(C <= 65535, "cai_eb_backward: too many channels");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (g_tparams) CAI_CUDA(cudaMemsetAsync(g_tparams, 0, sizeof(float) * static_cast<size_t>(C) * sp.P, stream));
  const int64_t per_c = N * HW;
  int gx = static_cast<int>((per_c + 127) / 128);
  const int cap = (dp.sm_count * 8 + static_cast<int>(C) - 1) / static_cast<int>(C);
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  dim3 grid(gx, static_cast<unsigned>(C));
  if (maxf <= 3)
    eb_backward_kernel<3><<<grid, 128, 0, stream>>>(x_tilde, tparams, g_lik, sp, bound_lik, layout, C, HW, N, g_x, g_tparams);
  else
    eb_backward_kernel<8><<<grid, 128, 0, stream>>>(x_tilde, tparams, g_lik, sp, bound_lik, layout, C, HW, N, g_x, g_tparams);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_eb_logits(const float *x, const float *tparams, const int32_t *filters_host, int32_t n_filters,
                  const float *g_out, int64_t C, int64_t L, float *out, float *g_x, cai_stream_t stream_) {
  CAI_CHECK_ARG(C >= 0 && L >= 0, "cai_eb_logits: bad size");
  if (C * L == 0) return CAI_OK;
  CAI_CHECK_ARG(x && tparams && (out || g_x) && (!g_x || g_out), "cai_eb_logits: NULL pointer");
  EbSpec sp;
  const int maxf = make_spec(filters_host, n_filters, &sp);
  CAI_CHECK_ARG(maxf > 0, "cai_eb_logits: unsupported filters");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(dp, C * L, 256);
  if (maxf <= 3)
    eb_logits_kernel<3><<<grid, 256, 0, stream>>>(x, tparams, g_out, sp, C, L, out, g_x);
  else
    eb_logits_kernel<8><<<grid, 256, 0, stream>>>(x, tparams, g_out, sp, C, L, out, g_x);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

}  // extern "C"
