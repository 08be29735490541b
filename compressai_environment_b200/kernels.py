"""Thin tensor wrappers over the elementwise / fused kernels of the C ABI (``cai_gc_quantize_index``,
``cai_eb_quantize_index``, ``cai_dequantize`` ...).  Torch is only used for device memory and streams.

Layout handling: a latent is a logical (N, C, *spatial) tensor stored either contiguous (NCHW) or
channels-last (NHWC, what the conv kernels of this package produce).  Integer outputs are always in
coder order, shape [N, C * prod(spatial)].
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import CAI_LAYOUT_NCHW, CAI_LAYOUT_NHWC, CaiError, check, current_stream, lib, ptr, require_cuda


def _ncs(t: torch.Tensor) -> Tuple[int, int, int]:
    if t.dim() < 2:
        raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
    N, C = int(t.size(0)), int(t.size(1))
    hw = 1
    for s in t.shape[2:]:
        hw *= int(s)
    return N, C, hw


def layout_of(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Return (tensor usable as-is, layout code); copies only if the storage is neither NCHW nor NHWC."""
    if t.is_contiguous():
        return t, CAI_LAYOUT_NCHW
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return t, CAI_LAYOUT_NHWC
    return t.contiguous(), CAI_LAYOUT_NCHW


def _same_layout(t: Optional[torch.Tensor], ref: torch.Tensor, layout: int):
    if t is None:
        return None
    require_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    if t.shape != ref.shape:
        t = t.expand(ref.shape)
    if layout == CAI_LAYOUT_NHWC:
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


def gc_quantize_index(y: Optional[torch.Tensor], scales: Optional[torch.Tensor], means: Optional[torch.Tensor],
                      scale_table: torch.Tensor, scale_bound: float):
    """Fused EntropyModel.quantize("symbols") + GaussianConditional.build_indexes.
    Returns (sym, idx) int32 [N, C*HW] in coder order (either may be None if its input is None)."""
    ref = y if y is not None else scales
    require_cuda(ref)
    ref = ref.detach()
    if ref.dtype != torch.float32:
        ref = ref.float()
    ref, layout = layout_of(ref)
    N, C, HW = _ncs(ref)
    ty = _same_layout(y.detach(), ref, layout) if y is not None else None
    ts = _same_layout(scales.detach(), ref, layout) if scales is not None else None
    tm = _same_layout(means.detach(), ref, layout) if means is not None else None
    dev = ref.device
    sym = torch.empty((N, C * HW), dtype=torch.int32, device=dev) if ty is not None else None
    idx = torch.empty((N, C * HW), dtype=torch.int32, device=dev) if ts is not None else None
    tab = scale_table.detach().to(device=dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        check(lib().cai_gc_quantize_index(ptr(ty), ptr(ts), ptr(tm), ptr(tab), int(tab.numel()), float(scale_bound),
                                          layout, N, C, HW, ptr(sym), ptr(idx), current_stream()),
              "cai_gc_quantize_index")
    return sym, idx


def eb_quantize_index(x: torch.Tensor, medians: torch.Tensor, want_sym: bool = True, want_idx: bool = True):
    """EntropyBottleneck front end: sym = rint(x - median[c]), idx = c; int32 [N, C*HW] coder order."""
    require_cuda(x)
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    x, layout = layout_of(x)
    N, C, HW = _ncs(x)
    dev = x.device
    med = medians.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    if med.numel() != C:
        raise ValueError("medians must have one entry per channel")
    sym = torch.empty((N, C * HW), dtype=torch.int32, device=dev) if want_sym else None
    idx = torch.empty((N, C * HW), dtype=torch.int32, device=dev) if want_idx else None
    with torch.cuda.device(dev):
        check(lib().cai_eb_quantize_index(ptr(x), ptr(med), layout, N, C, HW, ptr(sym), ptr(idx), current_stream()),
              "cai_eb_quantize_index")
    return sym, idx


def channel_indexes(N: int, C: int, HW: int, device) -> torch.Tensor:
    """EntropyBottleneck._build_indexes (entropy_models.py:518-529) in coder order."""
    idx = torch.empty((N, C * HW), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().cai_eb_quantize_index(None, None, CAI_LAYOUT_NCHW, N, C, HW, None, ptr(idx), current_stream()),
              "cai_eb_quantize_index")
    return idx


def dequantize(sym: torch.Tensor, means: Optional[torch.Tensor], medians: Optional[torch.Tensor], shape,
               memory_format=torch.contiguous_format) -> torch.Tensor:
    """EntropyModel.dequantize from coder-order int32 symbols to a float32 latent of logical ``shape``."""
    require_cuda(sym)
    shape = tuple(int(s) for s in shape)
    dev = sym.device
    nhwc = memory_format == torch.channels_last and len(shape) == 4
    out = torch.empty(shape, dtype=torch.float32, device=dev,
                      memory_format=torch.channels_last if nhwc else torch.contiguous_format)
    layout = CAI_LAYOUT_NHWC if nhwc else CAI_LAYOUT_NCHW
    N, C, HW = _ncs(out)
    tm = _same_layout(means.detach(), out, layout) if means is not None else None
    med = None
    if medians is not None:
        med = medians.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    sym = sym.contiguous()
    with torch.cuda.device(dev):
        check(lib().cai_dequantize(ptr(sym), ptr(tm), ptr(med), layout, N, C, HW, ptr(out), current_stream()),
              "cai_dequantize")
    return out


def pixels_to_float(x_u8: torch.Tensor) -> torch.Tensor:
    """uint8 image tensor (any shape) -> float32 in [0, 1]: x / 255, the reference callers' ToTensor() convention
    (examples/codec.py:112-128), computed on the device."""
    require_cuda(x_u8)
    if x_u8.dtype != torch.uint8:
        raise CaiError("pixels_to_float expects a uint8 tensor")
    x_u8 = x_u8.contiguous()
    out = torch.empty(x_u8.shape, dtype=torch.float32, device=x_u8.device)
    with torch.cuda.device(x_u8.device):
        check(lib().cai_pixels_u8_to_f32(ptr(x_u8), x_u8.numel(), ptr(out), current_stream()), "cai_pixels_u8_to_f32")
    return out


def pixels_to_u8(x: torch.Tensor) -> torch.Tensor:
    """float32 reconstruction -> uint8: round(clamp(x, 0, 1) * 255) (half to even), computed on the device."""
    require_cuda(x)
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().cai_pixels_f32_to_u8(ptr(x), x.numel(), ptr(out), current_stream()), "cai_pixels_f32_to_u8")
    return out


class ArWeights:
    """Context model + entropy-parameter network of the autoregressive models, laid out for ``cai_ar_encode`` /
    ``cai_ar_decode`` (row-major fp32 matrices, reduction dimension zero-padded to a multiple of 4; the masked
    convolution keeps its (k/2)*k + k/2 causal taps in (ky, kx) order with k = tap * M + channel)."""

    def __init__(self, ctx_weight, ctx_bias, convs, slope: float = 0.01):
        require_cuda(ctx_weight, "context_prediction.weight")
        n_ctx, M, k, k2 = ctx_weight.shape
        assert k == k2 and k % 2 == 1 and len(convs) == 3
        ntaps = (k // 2) * k + k // 2
        w = ctx_weight.detach().float().permute(0, 2, 3, 1).reshape(n_ctx, k * k * M)[:, :ntaps * M]
        self.w_ctx, self.b_ctx = w.contiguous(), ctx_bias.detach().float().contiguous()
        self.mats = []
        for cw, cb in convs:
            m = cw.detach().float().reshape(cw.shape[0], -1)
            kp = (m.shape[1] + 3) // 4 * 4
            if kp != m.shape[1]:
                m = torch.nn.functional.pad(m, (0, kp - m.shape[1]))
            self.mats.append((m.contiguous(), cb.detach().float().contiguous()))
        self.M, self.n_ctx, self.ksize, self.slope = int(M), int(n_ctx), int(k), float(slope)
        self.widths = [int(cw.shape[0]) for cw, _ in convs]
        self.in_width = int(convs[0][0].shape[1])

    def desc(self, params: torch.Tensor, scale_table: torch.Tensor, scale_bound: float, cluster: int = 0, group: int = 0,
             flags: int = 0):
        """``params``: fp32 [B, H, W, P] contiguous (NHWC).  Returns (ArDesc, keep-alive tuple)."""
        from ._lib import ArDesc

        B, H, W, P = (int(v) for v in params.shape)
        if P + self.n_ctx != self.in_width:
            raise ValueError(f"entropy_parameters expects {self.in_width} input channels, got {P} + {self.n_ctx}")
        tab = scale_table.detach().to(device=params.device, dtype=torch.float32).contiguous()
        d = ArDesc()
        d.w_ctx, d.b_ctx = ptr(self.w_ctx), ptr(self.b_ctx)
        (d.w1, d.b1), (d.w2, d.b2), (d.w3, d.b3) = ((ptr(m), ptr(b)) for m, b in self.mats)
        d.params, d.scale_table = ptr(params), ptr(tab)
        d.scale_bound, d.slope = float(scale_bound), self.slope
        d.T, d.B, d.H, d.W, d.M, d.P, d.n_ctx = int(tab.numel()), B, H, W, self.M, P, self.n_ctx
        d.n1, d.n2, d.n3 = self.widths
        d.ksize, d.cluster, d.group, d.flags = self.ksize, int(cluster), int(group), int(flags)
        return d, (tab, params)


def ar_encode(weights: ArWeights, y: torch.Tensor, params: torch.Tensor, scale_table, scale_bound: float,
              cluster: int = 0, group: int = 0):
    """Encoder scan of the autoregressive models (reference ``_compress_ar``, models/google.py:535-577) for a batch.
    ``y`` fp32 [B, H, W, M] and ``params`` fp32 [B, H, W, P], both NHWC-contiguous.  Returns (sym, idx) int32
    [B, H*W*M] in the reference's coding order (pixel-major) and the padded y_hat [B, H+2p, W+2p, M]."""
    import ctypes

    require_cuda(y)
    assert y.dtype == torch.float32 and params.dtype == torch.float32 and y.is_contiguous() and params.is_contiguous()
    B, H, W, M = (int(v) for v in y.shape)
    p = weights.ksize // 2
    dev = y.device
    y_hat = torch.zeros((B, H + 2 * p, W + 2 * p, M), dtype=torch.float32, device=dev)
    sym = torch.empty((B, H * W * M), dtype=torch.int32, device=dev)
    idx = torch.empty((B, H * W * M), dtype=torch.int32, device=dev)
    d, keep = weights.desc(params, scale_table, scale_bound, cluster, group)
    with torch.cuda.device(dev):
        check(lib().cai_ar_encode(ctypes.byref(d), ptr(y), ptr(y_hat), ptr(sym), ptr(idx), current_stream()), "cai_ar_encode")
    return sym, idx, y_hat


def ar_decode(weights: ArWeights, table, words: torch.Tensor, word_begin: torch.Tensor, params: torch.Tensor, scale_table,
              scale_bound: float, cluster: int = 0, group: int = 0, want_symbols: bool = False, use_lut: bool = False):
    """Decoder scan (reference ``_decompress_ar``, models/google.py:620-661): one rANS stream per image, M symbols per
    latent pixel in raster order.  ``table``: coder.CdfTable; ``words`` / ``word_begin``: the packed strings on the
    device (coder.strings_to_device).  Returns (padded y_hat [B, H+2p, W+2p, M], status int32 [B], symbols or None)."""
    import ctypes

    require_cuda(params)
    assert params.dtype == torch.float32 and params.is_contiguous()
    B, H, W, _ = (int(v) for v in params.shape)
    p, M = weights.ksize // 2, weights.M
    dev = params.device
    y_hat = torch.zeros((B, H + 2 * p, W + 2 * p, M), dtype=torch.float32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    sym = torch.empty((B, H * W * M), dtype=torch.int32, device=dev) if want_symbols else None
    d, keep = weights.desc(params, scale_table, scale_bound, cluster, group, 1 if use_lut else 0)
    with torch.cuda.device(dev):
        check(lib().cai_ar_decode(ctypes.byref(d), table.handle, ptr(words), ptr(word_begin), ptr(y_hat), ptr(sym),
                                  ptr(status), current_stream()), "cai_ar_decode")
    return y_hat, status, sym
