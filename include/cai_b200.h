/*
 * cai_b200.h -- C ABI of libcai_b200.so, the B200 (sm_100a) implementation of CompressAI's codec hot
 * path.  This is the drop-in boundary: every entry point below replaces one function of the
 * reference's two native extensions (compressai.ans, compressai._CXX) or one torch op sequence of
 * compressai.entropy_models / compressai.layers that sits on the compress / decompress / update /
 * training-forward path.  Paths cited are relative to the reference tree.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No torch / pybind types.
 *   - Every pointer is DEVICE memory unless its name ends in _host.
 *   - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *     results are ordered on that stream.  Calls that must read back a size say so.
 *   - Return value: 0 = ok, < 0 = error (CAI_E_*); the message is available from cai_last_error()
 *     (thread local).  No C++ exception crosses this boundary.
 *   - Per-string / per-row failures detected on the device are written to the `status` arrays
 *     (0 = ok, CAI_S_* otherwise); they never abort the launch.
 *   - "coder order" = the order the reference feeds symbols to the coder: for a (N, C, H, W) latent,
 *     string b is image b flattened as C, H, W (compressai/entropy_models/entropy_models.py:259-267).
 */
#ifndef CAI_B200_H_
#define CAI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define CAI_ABI_VERSION 2

/* host-side error codes */
#define CAI_OK 0
#define CAI_E_INVALID (-1)  /* bad argument */
#define CAI_E_CUDA (-2)     /* CUDA runtime error, see cai_last_error() */
#define CAI_E_NO_DEVICE (-3)
#define CAI_E_TOO_LARGE (-4) /* table does not fit the kernel's limits */

/* device-side per-string / per-row status */
#define CAI_S_OK 0
#define CAI_S_OVERFLOW 1   /* encoder: slot capacity exceeded */
#define CAI_S_BAD_INDEX 2  /* cdf index outside [0, K) */
#define CAI_S_BAD_PMF 3    /* pmf has a negative / non-finite entry (std::domain_error in ops.cpp:46-52) */
#define CAI_S_ZERO_PMF 4   /* pmf sums to zero (ops.cpp:60-64) */
#define CAI_S_NO_DONOR 5   /* no symbol left to steal frequency from (assert in ops.cpp:88) */
#define CAI_S_TRUNCATED 6  /* decoder ran past the end of its string (zeros were fed) */

/* memory layout of a 4-D latent handed to the fused kernels */
#define CAI_LAYOUT_NCHW 0 /* contiguous N, C, H, W */
#define CAI_LAYOUT_NHWC 1 /* channels-last storage of a logical N, C, H, W tensor */

typedef void *cai_stream_t;            /* cudaStream_t */
typedef struct cai_table *cai_table_t; /* opaque: packed CDF tables resident in HBM */

int cai_abi_version(void);
const char *cai_last_error(void);
/* Number of SMs / bytes of shared memory per block of the current device (for host-side planning). */
int cai_device_info(int *sm_count, int *max_smem_per_block);

/* ------------------------------------------------------------------------------------------------
 * CDF tables.  Replaces the per-call list -> std::vector<std::vector<int>> conversion of
 * rans_interface.cpp:108-113 / :215-221 (4.2 ms per call for the 64 x 3133 Gaussian table): tables
 * are packed ONCE per update() into a ragged uint16 blob (+ row metadata + a decode lookup table)
 * that the coder kernels stage into shared memory with one TMA bulk copy.
 *   cdfs     int32 [K, Lmax]  (EntropyModel._quantized_cdf)
 *   cdf_len  int32 [K]        (EntropyModel._cdf_length)
 *   offsets  int32 [K]        (EntropyModel._offset)
 * Synchronises `stream` once (it reads back the packed size).
 * ---------------------------------------------------------------------------------------------- */
int cai_table_create(const int32_t *cdfs, const int32_t *cdf_len, const int32_t *offsets, int32_t K,
                     int32_t Lmax, cai_stream_t stream, cai_table_t *out);
void cai_table_destroy(cai_table_t t);
/* K, bytes of the blob, decode-LUT buckets per row, 1 if the blob fits shared memory */
int cai_table_info(cai_table_t t, int32_t *K, int64_t *blob_bytes, int32_t *lut_buckets, int32_t *in_smem);

/* ------------------------------------------------------------------------------------------------
 * rANS coder.  Bit-exact with ryg_rans rans64.h:59-142 as driven by rans_interface.cpp (16-bit
 * precision, 4-bit bypass escapes).  One warp-cooperative coder per string.
 *
 * Strings: B strings; string b covers elements [str_begin[b], str_begin[b+1]) of `symbols` /
 * `indexes` when str_begin != NULL (int64 [B+1], device), else [b*n_per_string, (b+1)*n_per_string).
 * ---------------------------------------------------------------------------------------------- */

/* Words a slot must hold so that no input can overflow it (multiple of 32). */
int64_t cai_rans_slot_words(int64_t n_symbols);

/*
 * RansEncoder::encode_with_indexes (rans_interface.cpp:202-213) for B strings.
 *   slots    uint32 [B, slot_words]: string b's words end at slots + (b+1)*slot_words; the byte
 *            string is the LAST n_words[b] words of the slot, little endian.
 *   n_words  int32 [B]
 *   status   int32 [B] or NULL
 */
int cai_rans_encode_batch(cai_table_t t, const int32_t *symbols, const int32_t *indexes,
                          const int64_t *str_begin, int64_t n_per_string, int32_t B, uint32_t *slots,
                          int64_t slot_words, int32_t *n_words, int32_t *status, cai_stream_t stream);

/*
 * Gather the used tail of every slot into one contiguous buffer:
 *   out_begin int64 [B+1] (device, word offsets, written by this call; out_begin[B] = total words)
 *   out_words uint32 [>= sum(n_words)]   (pass NULL to only compute out_begin)
 */
int cai_rans_compact(const uint32_t *slots, int64_t slot_words, const int32_t *n_words, int32_t B,
                     int64_t *out_begin, uint32_t *out_words, int64_t out_capacity_words,
                     cai_stream_t stream);

/*
 * RansDecoder::decode_with_indexes (rans_interface.cpp:215-284) for B strings.
 *   words      uint32: all strings; string b starts at word_begin[b] (int64, device) and has
 *              word_count[b] words (int32 [B]); word_count == NULL: word_begin has B+1 entries and
 *              string b ends where string b+1 begins (packed layout of cai_rans_compact).
 *              With word_count the strings may be read in place from the encoder's slots.
 *   out        int32 symbols in coder order (same indexing as `indexes`)
 *   state      NULL, or uint64 [B, 2] = (rANS state x, next word position) carried between calls:
 *              RansDecoder::set_stream / decode_stream (rans_interface.cpp:286-359).
 *              resume = 0: initialise from the first two words (and store the final state if
 *              state != NULL); resume = 1: continue from `state`.
 */
int cai_rans_decode_batch(cai_table_t t, const uint32_t *words, const int64_t *word_begin,
                          const int32_t *word_count, const int32_t *indexes, const int64_t *str_begin, int64_t n_per_string,
                          int32_t B, int32_t *out, uint64_t *state, int32_t resume, int32_t *status,
                          cai_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused quantise + index kernels (HBM-bound, one pass).
 * Inputs are logical (N, C, HW) latents stored as `layout`; outputs are int32 in CODER ORDER.
 * ---------------------------------------------------------------------------------------------- */

/*
 * GaussianConditional: EntropyModel.quantize(y, "symbols", means) (entropy_models.py:155-180) fused
 * with GaussianConditional.build_indexes(scales) (entropy_models.py:684-689):
 *   sym = int32(rint(y - mean)),  idx = #{ j < T-1 : !(max(scale, bound) <= table[j]) } counted from
 *   the left, i.e. (T-1) - sum_j [max(scale, bound) <= table[j]].
 * means may be NULL.  y / scales / means may be NULL individually to skip that half
 * (y == NULL -> only indexes; scales == NULL -> only symbols).
 */
int cai_gc_quantize_index(const float *y, const float *scales, const float *means,
                          const float *scale_table, int32_t T, float scale_bound, int32_t layout,
                          int64_t N, int64_t C, int64_t HW, int32_t *sym, int32_t *idx,
                          cai_stream_t stream);

/*
 * EntropyBottleneck.compress front end (entropy_models.py:518-541): sym = int32(rint(x - median[c])),
 * idx = c.   sym or idx may be NULL.
 */
int cai_eb_quantize_index(const float *x, const float *medians, int32_t layout, int64_t N, int64_t C,
                          int64_t HW, int32_t *sym, int32_t *idx, cai_stream_t stream);

/*
 * EntropyModel.dequantize (entropy_models.py:188-197) from coder-order int32 symbols to a float
 * latent stored as `layout`:  out = float(sym) + mean.   Exactly one of means (full tensor, same
 * layout as out) / medians ([C]) may be non-NULL; both NULL -> plain conversion.
 */
int cai_dequantize(const int32_t *sym, const float *means, const float *medians, int32_t layout,
                   int64_t N, int64_t C, int64_t HW, float *out, cai_stream_t stream);

/* 8-bit pixels <-> unit-range fp32 on the device: out = in / 255 (the ToTensor() convention the reference's callers
 * apply on the host before model.compress, examples/codec.py:112-128) and out = round_half_even(clamp(in, 0, 1) * 255)
 * (NaN -> 0).  They let the serving path move uint8 images over PCIe (a quarter of the fp32 bytes).  n elements,
 * any layout; both buffers 16-byte aligned. */
int cai_pixels_u8_to_f32(const uint8_t *in, int64_t n, float *out, cai_stream_t stream);
int cai_pixels_f32_to_u8(const float *in, int64_t n, uint8_t *out, cai_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Table build.  pmf_to_quantized_cdf (compressai/cpp_exts/ops/ops.cpp:40-109) for K rows at once with
 * the caller's row convention of EntropyModel._pmf_to_cdf (entropy_models.py:204-212):
 *   row k = cat(pmf[k, :pmf_len[k]], tail[k])  ->  cdf[k, :pmf_len[k]+2], zero padded to Lp+2.
 * pmf float32 [K, Lp]; tail float32 [K]; cdf int32 [K, Lp + 2]; status int32 [K].
 * tail == NULL: rows are plain pmfs of pmf_len[k] entries -> pmf_len[k]+1 cdf entries
 * (the bare ops.cpp signature).
 * ---------------------------------------------------------------------------------------------- */
int cai_pmf_to_quantized_cdf(const float *pmf, const int32_t *pmf_len, const float *tail, int32_t K,
                             int32_t Lp, int32_t precision, int32_t *cdf, int32_t *status,
                             cai_stream_t stream);


/* ------------------------------------------------------------------------------------------------
 * Likelihood kernels (forward + backward), elementwise over HBM.  All tensors of one call share one
 * memory layout (any), n = number of elements, except the entropy-bottleneck calls which need the
 * channel of every element and therefore take (layout, N, C, HW).
 * ---------------------------------------------------------------------------------------------- */

/*
 * GaussianConditional.forward (entropy_models.py:669-682) = quantize + _likelihood (:650-667) +
 * LowerBound (compressai/ops/bound_ops.py:36-42), one pass:
 *   mode 0 (training): y_hat = y + noise           (means are ignored by quantize("noise"), :161-165)
 *   mode 1 (eval):     y_hat = rint(y - mean) + mean
 *   mode 2:            y_hat = y
 *   lik = max( Phi((.5 - |y_hat - mean|)/s) - Phi((-.5 - |y_hat - mean|)/s), bound_lik ),  s = max(scale, bound_scale)
 * means / noise / y_hat / lik may be NULL where unused; bound_lik <= 0 disables the likelihood bound.
 */
int cai_gc_forward(const float *y, const float *scales, const float *means, const float *noise, int32_t mode,
                   float bound_scale, float bound_lik, int64_t n, float *y_hat, float *lik, cai_stream_t stream);

/* Backward of the above w.r.t. y_hat, scales and means given g_lik, including
 * both LowerBound gates.  Output pointers may be NULL. */
int cai_gc_backward(const float *y_hat, const float *scales, const float *means, const float *g_lik,
                    float bound_scale, float bound_lik, int64_t n, float *g_y, float *g_scales, float *g_means,
                    cai_stream_t stream);

/*
 * EntropyBottleneck.forward (entropy_models.py:471-516) without the two permutes: quantize + _likelihood
 * (:457-469, built on _logits_cumulative :436-455) + LowerBound.
 *   tparams float32 [C, P]: per channel, for every layer i: softplus(_matrix_i) row major, _bias_i,
 *   tanh(_factor_i) (no factor for the last layer).  filters_host: HOST array of the hidden widths
 *   (default (3, 3, 3, 3) -> P = 58).  medians float32 [C].  mode as in cai_gc_forward.
 */
int cai_eb_forward(const float *x, const float *tparams, const int32_t *filters_host, int32_t n_filters,
                   const float *medians, const float *noise, int32_t mode, float bound_lik, int32_t layout, int64_t N,
                   int64_t C, int64_t HW, float *out, float *lik, cai_stream_t stream);

/* Backward (Appendix D.3): g_x (layout of x) and g_tparams [C, P] (overwritten) from g_lik and the saved
 * x_tilde = `out` of the forward.  The sign factor carries no gradient (:464-465). */
int cai_eb_backward(const float *x_tilde, const float *tparams, const int32_t *filters_host, int32_t n_filters,
                    const float *g_lik, float bound_lik, int32_t layout, int64_t N, int64_t C, int64_t HW, float *g_x,
                    float *g_tparams, cai_stream_t stream);

/* _logits_cumulative at per-channel sample points x [C, L] -> out [C, L] (update() :418-421, loss() :431-434).
 * If g_x != NULL also returns g_x = g_out * d out / d x (parameters are treated as constants). */
int cai_eb_logits(const float *x, const float *tparams, const int32_t *filters_host, int32_t n_filters,
                  const float *g_out, int64_t C, int64_t L, float *out, float *g_x, cai_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Transforms: tcgen05 implicit-GEMM convolution / transposed convolution / GDN.
 * Replaces the cuDNN / ATen kernels behind compressai/models/utils.py:128-146 (conv, deconv),
 * compressai/layers/gdn.py:77-92 (GDN / IGDN) as used by the g_a / g_s / h_a / h_s stacks of
 * compressai/models/google.py:134-152, :219-254, :339-353.
 *
 * Activations are "split planes": an fp32 NHWC tensor stored as two bf16 NHWC tensors hi = bf16(x),
 * lo = bf16(x - hi) (same bytes as fp32; x ~= hi + lo to 2^-17 relative).  One call computes, for one
 * output phase grid of Hp x Wp pixels per image,
 *     D[pixel, co] = sum_{t < ntaps} sum_{ci} A[n, i*is + dy[t], j*is + dx[t], ci] * Wt[co, t, ci]
 * (out-of-range input pixels read as zero) and writes pixel (i*os + o0y, j*os + o0x) of the output.
 * w_packed holds the weights pre-split and pre-tiled by the host layer (see transforms.py:pack_weights):
 * [n_tile][k-step][hi | lo][BN x 32 bf16 in UMMA canonical K-major order] (k-step = 32 channels of one tap).
 * k-step order: taps come in groups of glen[g] consecutive taps that share dy and whose dx differ by multiples of
 * `is` (their input pixel sets coincide up to a shift, so the later taps of a group hit L1); the kernel walks
 *   for group g: for channel chunk kc: for tap t in group g
 * and w_packed is laid out in that order.  glen[] all zero = one tap per group (k-step = t * kchunks + kc).
 * epilogue: 0 linear, 1 ReLU, 2 LeakyReLU(0.01), 3 GDN  out = aux * rsqrt(D + bias),
 *           4 IGDN out = aux * sqrt(D + bias)   (aux given as split planes shaped like the output).
 * Outputs (any subset): out_f32 fp32 NHWC, out_hi/lo split planes, sq_hi/lo planes of out^2 (the GDN
 * input), abs_hi/lo planes of |out| (the hyper-analysis input of ScaleHyperprior).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cai_conv_desc {
  const void *a_hi, *a_lo;   /* bf16 [N, H, W, Cin] */
  const void *w_packed;
  const float *bias;         /* [Cout] or NULL */
  const void *aux_hi, *aux_lo;
  float *out_f32;
  void *out_hi, *out_lo, *sq_hi, *sq_lo, *abs_hi, *abs_lo;
  int32_t N, H, W, Cin, Ho, Wo, Cout;
  int32_t Hp, Wp, os, o0y, o0x, is;
  int32_t ntaps;
  int32_t BN;                /* output channels per tile: multiple of 16, <= 256 */
  int32_t epilogue;
  float clamp_lo, clamp_hi;  /* clamp applied when clamp_lo < clamp_hi */
  int8_t dy[32], dx[32];
  int8_t glen[32];           /* tap group lengths, sum = ntaps (see above) */
  /* Fused GDN / IGDN (compressai/layers/gdn.py:77-92) as a second in-kernel GEMM: with v = D + bias,
   * out = v * rsqrt(gamma . v^2 + beta) (gdn_mode 1) or v * sqrt(...) (gdn_mode 2).  gdn_w = gamma packed like
   * w_packed with one tap ([kchunk][hi | lo][Cout x 32 bf16]); needs BN == Cout <= 256 and epilogue == 0.
   * NULL = no fusion. */
  const void *gdn_w;
  const float *gdn_beta;
  int32_t gdn_mode;
  /* 0: per-tile kernel, w_packed k-steps ordered tap group -> channel chunk -> tap (conv.cu).
   * 1: persistent TMA-fed kernel (conv_tma.cu), w_packed k-steps ordered channel chunk -> tap group -> tap; only
   *    valid when cai_conv_tma_eligible() accepts the same descriptor. */
  int32_t mode;
} cai_conv_desc;

int cai_conv_gemm(const cai_conv_desc *d, cai_stream_t stream);

/* 1 if the persistent TMA-fed kernel takes this layer (row-segment tiles of a grid at least 64 pixels wide, all output
 * channels in one tile of at most 128, input stride 1 or 2, grouped taps, 16-byte aligned planes), else 0.  The caller
 * then packs the weights in the order of mode 1 and sets d->mode = 1. */
int cai_conv_tma_eligible(const cai_conv_desc *d);

/* ------------------------------------------------------------------------------------------------
 * Training mode: weight gradient of a conv / transposed conv (csrc/wgrad.cu), replacing the cuDNN call behind
 * autograd of nn.Conv2d / nn.ConvTranspose2d (compressai/models/utils.py:128-146, examples/train.py:132-165):
 *   grad[cs][cb][ky][kx] = sum_{n,i,j} small[n,cs,i,j] * big[n,cb, stride*i + ky - pad, stride*j + kx - pad]
 * small fp32 NCHW [N, Cs, Hs, Ws] = dY of a conv / X of a transposed conv; big fp32 NCHW [N, Cb, Hb, Wb] = X of a conv /
 * dY of a transposed conv; grad fp32 [Cs, Cb, ksize, ksize] = the layer's weight gradient in torch's layout for both
 * kinds.  stride 1 or 2, ksize^2 <= 32.  workspace: device scratch of cai_conv_wgrad_workspace() bytes (256-byte
 * aligned; split bf16 operand planes + split-K partial sums).  The data gradient of either layer kind is the other
 * kind's forward (cai_conv_gemm with the same weights).
 * ---------------------------------------------------------------------------------------------- */
int64_t cai_conv_wgrad_workspace(int64_t N, int32_t Cs, int32_t Hs, int32_t Ws, int32_t Cb, int32_t Hb, int32_t Wb,
                                 int32_t ksize, int32_t stride);
int cai_conv_wgrad(const float *small, const float *big, int64_t N, int32_t Cs, int32_t Hs, int32_t Ws, int32_t Cb,
                   int32_t Hb, int32_t Wb, int32_t ksize, int32_t stride, int32_t pad, float *grad, void *workspace,
                   int64_t workspace_bytes, cai_stream_t stream);

/* fp32 latent / image (layout NCHW or NHWC) -> split planes [N, HW, Cpad] (channels zero padded to Cpad) */
int cai_split_planes(const float *x, int32_t layout, int64_t N, int64_t C, int64_t HW, int64_t Cpad, void *hi, void *lo,
                     cai_stream_t stream);

/* im2col for tiny Cin (the 3-channel first layer): split planes [N*Ho*Wo, Kpad], k = (ky*ksize + kx)*C + c */
int cai_im2col_split(const float *x, int32_t layout, int32_t N, int32_t C, int32_t H, int32_t W, int32_t Ho, int32_t Wo,
                     int32_t ksize, int32_t stride, int32_t pad, int32_t Kpad, void *hi, void *lo, cai_stream_t stream);

/* col2im gather for tiny Cout (the 3-channel last layer of g_s): cols fp32 [N*H*W, Npad] with
 * column (ky*ksize + kx)*Cout + co  ->  out fp32 [N, Cout, Ho, Wo] (out_layout NCHW) or NHWC, + bias, optional clamp */
int cai_col2im(const float *cols, const float *bias, int32_t N, int32_t Cout, int32_t H, int32_t W, int32_t Ho,
               int32_t Wo, int32_t ksize, int32_t stride, int32_t pad, int32_t Npad, int32_t out_layout, float clamp_lo,
               float clamp_hi, float *out, cai_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GDN / IGDN backward (training mode; forward: compressai/layers/gdn.py:77-92, formulas: SURVEY.md Appendix D.1).
 * NHWC fp32 tensors of n = P * C elements.  The two channel-mixing products (norm = gamma . x^2 + beta and
 * u = gamma^T t) are cai_conv_gemm calls; these entry points are the stages around them:
 *   prepare: t = g x norm^(-3/2) (GDN) / g x norm^(-1/2) (IGDN), as fp32 and split planes; p = g norm^(-/+ 1/2)
 *   finish : dx = p - x u (GDN) / p + x u (IGDN)
 *   params : dbeta[i] = s sum_pix t_i, dgamma[i][j] = s sum_pix t_i x_j^2, s = -1/2 (GDN) / +1/2 (IGDN)  (overwrites)
 * ---------------------------------------------------------------------------------------------- */
int cai_gdn_bwd_prepare(const float *x, const float *norm, const float *g, int32_t inverse, int64_t n, float *t,
                        void *t_hi, void *t_lo, float *p, cai_stream_t stream);
int cai_gdn_bwd_finish(const float *p, const float *x, const float *u, int32_t inverse, int64_t n, float *gx,
                       cai_stream_t stream);
int cai_gdn_bwd_params(const float *t, const float *x, int64_t P, int32_t C, int32_t inverse, float *g_beta,
                       float *g_gamma, cai_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Autoregressive context model scan (csrc/ar.cu), replacing the per-pixel Python loops of
 * JointAutoregressiveHierarchicalPriors._compress_ar / _decompress_ar (compressai/models/google.py:535-577,
 * :620-661; inherited by the Cheng2020 models, compressai/models/waseda.py:44-153).  Per latent pixel, in raster order:
 *   ctx = context_prediction (MaskedConv2d type "A", layers.py:52-78) on the ksize x ksize crop of y_hat
 *   (scales | means) = entropy_parameters(cat(params, ctx))   -- three 1x1 convolutions, LeakyReLU(slope) between
 *   idx = GaussianConditional.build_indexes(scales)
 *   encoder: sym = int32(rint(y - means)), y_hat = sym + means            (symbols / indexes coded afterwards by
 *                                                                          cai_rans_encode_batch, one string per image)
 *   decoder: sym = M symbols from the image's rANS stream (RansDecoder::set_stream / decode_stream,
 *            rans_interface.cpp:286-359), y_hat = sym + means
 * All tensors fp32 NHWC on the device.  Weight matrices are row-major [rows][K padded to a multiple of 4 with zeros]:
 *   w_ctx [n_ctx][ntaps * M]   the ntaps = (ksize/2) * ksize + ksize/2 unmasked taps in (ky, kx) order, k = tap * M + c
 *   w1 [n1][pad4(P + n_ctx)], w2 [n2][pad4(n1)], w3 [n3 = 2 M][pad4(n2)]
 * y_hat: [B, H + 2 (ksize/2), W + 2 (ksize/2), M] with a zero border (the decoder needs it zero-initialised).
 * sym / idx: int32 [B, H * W * M], pixel-major and channel-minor -- the order in which the reference pushes them.
 * One thread-block cluster of `cluster` CTAs (0 = 8) scans `group` images (0 = automatic); a row of a layer is always
 * summed in the same order, so encoder and decoder obtain bit-identical parameters whatever the launch shape.
 * ---------------------------------------------------------------------------------------------- */
typedef struct cai_ar_desc {
  const float *w_ctx, *b_ctx, *w1, *b1, *w2, *b2, *w3, *b3;
  const float *params;       /* [B, H, W, P]: hyper-synthesis output */
  const float *scale_table;  /* [T] */
  float scale_bound, slope;
  int32_t T, B, H, W, M, P, n_ctx, n1, n2, n3, ksize;
  int32_t cluster, group;
  int32_t flags;             /* bit 0: decoder also stages the decode LUT (slower on B200: it costs L1) */
} cai_ar_desc;

int cai_ar_encode(const cai_ar_desc *d, const float *y, float *y_hat, int32_t *sym, int32_t *idx, cai_stream_t stream);
/* words / word_begin [B + 1]: the packed strings as in cai_rans_decode_batch; sym may be NULL; status int32 [B] */
int cai_ar_decode(const cai_ar_desc *d, cai_table_t t, const uint32_t *words, const int64_t *word_begin, float *y_hat,
                  int32_t *sym, int32_t *status, cai_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CAI_B200_H_ */
