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
