// gdn.cu -- backward of GDN / IGDN (training mode), the CUDA-core parts.
//
// Reference forward: compressai/layers/gdn.py:77-92   n = beta + gamma . x^2 ;  GDN y = x n^-1/2, IGDN y = x n^+1/2.
// Backward, with upstream g:
//   GDN : t = g x n^-3/2 ; dx = g n^-1/2 - x (gamma^T t) ; dbeta = -1/2 sum_pix t ; dgamma_ij = -1/2 sum_pix t_i x_j^2
//   IGDN: t = g x n^-1/2 ; dx = g n^+1/2 + x (gamma^T t) ; dbeta = +1/2 sum_pix t ; dgamma_ij = +1/2 sum_pix t_i x_j^2
// The two channel-mixing products (n = gamma . x^2 and gamma^T t) are 1x1 GEMMs on the tcgen05 kernel of conv.cu;
// this file holds the elementwise stages around them and the pixel-reduction outer product for dgamma / dbeta.
// All tensors are NHWC (channel innermost), P = N*H*W pixels, C channels.
#include <cuda_bf16.h>

#include "common.cuh"

namespace cai {

__device__ __forceinline__ void split2(float v, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// stage 1: t (fp32 + split planes for the gamma^T GEMM) and p = g * n^(-/+ 1/2)
__global__ void __launch_bounds__(256)
gdn_bwd_prepare_kernel(const float *__restrict__ x, const float *__restrict__ norm, const float *__restrict__ g,
                       int inverse, int64_t n, float *__restrict__ t, __nv_bfloat16 *__restrict__ t_hi,
                       __nv_bfloat16 *__restrict__ t_lo, float *__restrict__ p) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float nv = __ldcs(norm + i), xv = __ldcs(x + i), gv = __ldcs(g + i);
    const float r = rsqrtf(nv);  // n^-1/2
    float tv, pv;
    if (inverse) {
      tv = gv * xv * r;
      pv = gv * sqrtf(nv);
    } else {
      tv = gv * xv * r * r * r;
      pv = gv * r;
    }
    t[i] = tv;
    p[i] = pv;
    __nv_bfloat16 h, l;
    split2(tv, h, l);
    t_hi[i] = h;
    t_lo[i] = l;
  }
}

// stage 3: dx = p -/+ x * u     (u = gamma^T t)
__global__ void __launch_bounds__(256)
gdn_bwd_finish_kernel(const float *__restrict__ p, const float *__restrict__ x, const float *__restrict__ u, int inverse,
                      int64_t n, float *__restrict__ gx) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float xu = __ldcs(x + i) * __ldcs(u + i);
    gx[i] = inverse ? (__ldcs(p + i) + xu) : (__ldcs(p + i) - xu);
  }
}

// stage 4: dgamma[i][j] += s * sum_pix t[pix][i] * x[pix][j]^2 ; dbeta[i] += s * sum_pix t[pix][i]   (s = -/+ 1/2)
// 16 x 16 threads; thread (ty, tx) owns entries (ty + 16 a, tx + 16 b); pixels are streamed through shared memory in
// tiles of kPix.  Partial sums are merged with atomics (one add per entry per CTA).
constexpr int kPix = 16;
constexpr int kMaxCB = 12;  // C <= 192

__global__ void __launch_bounds__(256)
gdn_bwd_params_kernel(const float *__restrict__ t, const float *__restrict__ x, int64_t P, int C, float scale,
                      float *__restrict__ g_beta, float *__restrict__ g_gamma) {
  extern __shared__ float sm[];
  float *st = sm;                 // [kPix][C]
  float *sx = sm + kPix * C;      // [kPix][C]  (x^2)
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nb = (C + 15) / 16;
  float acc[kMaxCB][kMaxCB];
#pragma unroll
  for (int a = 0; a < kMaxCB; ++a)
#pragma unroll
    for (int b = 0; b < kMaxCB; ++b) acc[a][b] = 0.f;
  float bsum[kMaxCB];
#pragma unroll
  for (int a = 0; a < kMaxCB; ++a) bsum[a] = 0.f;

  const int64_t tiles = (P + kPix - 1) / kPix;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t p0 = tile * kPix;
    __syncthreads();
    for (int e = threadIdx.x; e < kPix * C; e += 256) {
      const int64_t pix = p0 + e / C;
      float tv = 0.f, xv = 0.f;
      if (pix < P) {
        tv = __ldcs(t + p0 * C + e);
        xv = __ldcs(x + p0 * C + e);
      }
      st[e] = tv;
      sx[e] = xv * xv;
    }
    __syncthreads();
#pragma unroll 1
    for (int q = 0; q < kPix; ++q) {
      float tv[kMaxCB], xv[kMaxCB];
#pragma unroll
      for (int a = 0; a < kMaxCB; ++a) {
        tv[a] = (a < nb && ty + 16 * a < C) ? st[q * C + ty + 16 * a] : 0.f;
        xv[a] = (a < nb && tx + 16 * a < C) ? sx[q * C + tx + 16 * a] : 0.f;
      }
#pragma unroll
      for (int a = 0; a < kMaxCB; ++a) {
        if (a < nb) {
          bsum[a] += tv[a];
#pragma unroll
          for (int b = 0; b < kMaxCB; ++b)
            if (b < nb) acc[a][b] = fmaf(tv[a], xv[b], acc[a][b]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < kMaxCB; ++a) {
    const int i = ty + 16 * a;
    if (a < nb && i < C) {
      if (tx == 0) atomicAdd(g_beta + i, scale * bsum[a]);
#pragma unroll
      for (int b = 0; b < kMaxCB; ++b) {
        const int j = tx + 16 * b;
        if (b < nb && j < C) atomicAdd(g_gamma + static_cast<int64_t>(i) * C + j, scale * acc[a][b]);
      }
    }
  }
}

static int grid_for(const DeviceProps &dp, int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(dp.sm_count) * 16;
  if (g > cap) g = cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace cai

using namespace cai;

extern "C" {

int cai_gdn_bwd_prepare(const float *x, const float *norm, const float *g, int32_t inverse, int64_t n, float *t,
                        void *t_hi, void *t_lo, float *p, cai_stream_t stream_) {
  CAI_CHECK_ARG(n >= 0, "cai_gdn_bwd_prepare: n < 0");
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(x && norm && g && t && t_hi && t_lo && p, "cai_gdn_bwd_prepare: NULL pointer");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  gdn_bwd_prepare_kernel<<<grid_for(dp, n), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      x, norm, g, inverse, n, t, static_cast<__nv_bfloat16 *>(t_hi), static_cast<__nv_bfloat16 *>(t_lo), p);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_gdn_bwd_finish(const float *p, const float *x, const float *u, int32_t inverse, int64_t n, float *gx,
                       cai_stream_t stream_) {
  CAI_CHECK_ARG(n >= 0, "cai_gdn_bwd_finish: n < 0");
  if (n == 0) return CAI_OK;
  CAI_CHECK_ARG(p && x && u && gx, "cai_gdn_bwd_finish: NULL pointer");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  gdn_bwd_finish_kernel<<<grid_for(dp, n), 256, 0, static_cast<cudaStream_t>(stream_)>>>(p, x, u, inverse, n, gx);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

int cai_gdn_bwd_params(const float *t, const float *x, int64_t P, int32_t C, int32_t inverse, float *g_beta,
                       float *g_gamma, cai_stream_t stream_) {
  CAI_CHECK_ARG(P >= 0 && C >= 1 && C <= 16 * kMaxCB, "cai_gdn_bwd_params: C=%d not in [1, %d]", C, 16 * kMaxCB);
  CAI_CHECK_ARG(t && x && g_beta && g_gamma, "cai_gdn_bwd_params: NULL pointer");
  DeviceProps dp;
  int rc = get_device_props(&dp);
  if (rc != CAI_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CAI_CUDA(cudaMemsetAsync(g_beta, 0, sizeof(float) * C, stream));
  CAI_CUDA(cudaMemsetAsync(g_gamma, 0, sizeof(float) * static_cast<size_t>(C) * C, stream));
  if (P == 0) return CAI_OK;
  const int64_t tiles = (P + kPix - 1) / kPix;
  int grid = static_cast<int>(tiles < dp.sm_count * 2 ? tiles : dp.sm_count * 2);
  const size_t smem = sizeof(float) * 2 * kPix * static_cast<size_t>(C);
  gdn_bwd_params_kernel<<<grid, 256, smem, stream>>>(t, x, P, C, inverse ? 0.5f : -0.5f, g_beta, g_gamma);
  CAI_LAUNCH_CHECK();
  return CAI_OK;
}

}  // extern "C"
