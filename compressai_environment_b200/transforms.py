"""Analysis / synthesis transforms on the tcgen05 implicit-GEMM kernel (``cai_conv_gemm``).

Reference: the ``conv`` / ``deconv`` factories (compressai/models/utils.py:128-146: k=5, s=2, pad=k//2,
output_padding=s-1) build ``nn.Conv2d`` / ``nn.ConvTranspose2d``; GDN is compressai/layers/gdn.py:77-92;
the stacks are compressai/models/google.py:134-152, :219-254, :339-353.  The modules here keep torch's
parameter names and shapes (``weight`` [Cout, Cin, k, k] for conv, [Cin, Cout, k, k] for transposed conv,
``bias`` [Cout]) so reference checkpoints load unchanged.

Inference data flow (``run_stack``): activations travel between layers as *split planes* (two bf16 NHWC
tensors hi + lo, see include/cai_b200.h) so that the A operand of the implicit GEMM is a pure cp.async copy;
weights are split and tiled once per parameter version (``pack_weights``) into the UMMA canonical layout.
  conv   -> one launch (one phase, k*k taps)
  deconv -> stride^2 launches (one per output phase; 9/6/6/4 taps for k5 s2)
  GDN    -> one launch: 1x1 GEMM of the x^2 planes with gamma, epilogue out = x * rsqrt(. + beta)
  Cin=3 first layer  -> im2col to K=80 + 1x1 GEMM;  Cout=3 last layer -> 1x1 GEMM to N=80 + col2im gather
ReLU / LeakyReLU / abs / clamp are folded into the producing launch's epilogue.

Training (autograd): ``_ConvFunction`` runs the forward on the same implicit-GEMM kernel, the data gradient as the
OTHER layer kind's forward with the same weights (dgrad of a conv is a transposed conv and vice versa), and the weight
gradient on the tcgen05 pixel-reduction GEMM of csrc/wgrad.cu (``cai_conv_wgrad``); GDN and the likelihoods have their
own forward / backward kernels.  No cuDNN call is left on the training step's transforms.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _cache, _lib
from ._lib import CAI_LAYOUT_NCHW, CAI_LAYOUT_NHWC, ConvDesc, check, current_stream, lib, ptr, require_cuda

_BK = 32

# Optional profiling hook used by bench.py (same convention as coder.TIMING): when set to a dict, every
# cai_conv_gemm launch appends a (start, end) CUDA-event pair recorded on the launching stream.
TIMING = None
# Same hook with one (label, start, end) record per launch, the label naming the layer shape (tools/layer_times.py).
DETAIL = None
# Which eligible layers go to the persistent TMA-fed kernel (conv_tma.cu): "auto" = where it measures faster than
# the per-tile kernel on B200 (single-tap GEMMs: the K = 80 first layer and the last-layer GEMM, whose tiles are
# epilogue-bound), "all" = every layer cai_conv_tma_eligible() accepts (tests, experiments), "none".
TMA_POLICY = os.environ.get("CAI_TMA_POLICY", "auto")


class Conv2d(_cache.CacheOwner, nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=5, stride=2, padding=None):
        super().__init__()
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride = int(kernel_size), int(stride)
        self.padding = self.kernel_size // 2 if padding is None else int(padding)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):  # same init as nn.Conv2d
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = self.in_channels * self.kernel_size * self.kernel_size
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad or self.bias.requires_grad):
            return _ConvFunction.apply(x, self.weight, self.bias, self)  # forward, dgrad and wgrad on our kernels
        return run_stack([self], x)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"


class ConvTranspose2d(_cache.CacheOwner, nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=5, stride=2, output_padding=1, padding=None):
        super().__init__()
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride = int(kernel_size), int(stride)
        self.padding = self.kernel_size // 2 if padding is None else int(padding)
        self.output_padding = int(output_padding)
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):  # same init as nn.ConvTranspose2d (fan_in computed on dim 1)
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = self.out_channels * self.kernel_size * self.kernel_size
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad or self.bias.requires_grad):
            return _ConvFunction.apply(x, self.weight, self.bias, self)  # forward, dgrad and wgrad on our kernels
        return run_stack([self], x)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"


class TransformStack(nn.Sequential):
    """``nn.Sequential`` whose inference forward runs the whole stack on the fused kernels (same submodule
    names / state_dict keys as the reference's plain Sequential)."""

    def forward(self, x, **kw):
        if torch.is_grad_enabled() and (getattr(x, "requires_grad", False) or any(p.requires_grad for p in self.parameters())):
            for m in self:
                x = m(x)
            return x
        return run_stack(list(self), x, **kw)


# ---- split planes ---------------------------------------------------------------------------------------------
class Planes:
    """fp32 NHWC activation stored as two bf16 NHWC tensors (hi + lo)."""

    __slots__ = ("hi", "lo", "N", "H", "W", "C")

    def __init__(self, hi, lo, N, H, W, C):
        self.hi, self.lo, self.N, self.H, self.W, self.C = hi, lo, N, H, W, C

    @staticmethod
    def empty(N, H, W, C, device):
        return Planes(torch.empty((N, H, W, C), dtype=torch.bfloat16, device=device),
                      torch.empty((N, H, W, C), dtype=torch.bfloat16, device=device), N, H, W, C)


def to_planes(x: Tensor, c_pad: Optional[int] = None) -> Planes:
    require_cuda(x, "inputs")
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    N, C, H, W = x.shape
    if x.is_contiguous():
        layout = CAI_LAYOUT_NCHW
    elif x.is_contiguous(memory_format=torch.channels_last):
        layout = CAI_LAYOUT_NHWC
    else:
        x, layout = x.contiguous(), CAI_LAYOUT_NCHW
    Cp = C if c_pad is None else c_pad
    out = Planes.empty(N, H, W, Cp, x.device)
    with torch.cuda.device(x.device):
        check(lib().cai_split_planes(ptr(x), layout, N, C, H * W, Cp, ptr(out.hi), ptr(out.lo), current_stream()),
              "cai_split_planes")
    return out


# ---- weight packing -------------------------------------------------------------------------------------------
def _c16(c: int) -> int:
    """Channel counts are padded to a multiple of 16 (UMMA N granularity at M = 128; K chunks of 8)."""
    return (int(c) + 15) // 16 * 16


def _pad_taps(wt: Tensor, cout_p: int, cin_p: int) -> Tensor:
    T, cout, cin = wt.shape
    if cout == cout_p and cin == cin_p:
        return wt
    out = torch.zeros((T, cout_p, cin_p), dtype=wt.dtype, device=wt.device)
    out[:, :cout, :cin] = wt
    return out


def _pad_vec(v: Tensor, n: int, fill: float = 0.0) -> Tensor:
    v = v.detach().float().reshape(-1)
    if v.numel() == n:
        return v.contiguous()
    out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
    out[:v.numel()] = v
    return out


def _choose_bn(cout: int) -> int:
    if cout <= 256:
        return cout
    for parts in range(2, 64):
        if cout % parts == 0 and (cout // parts) % 16 == 0 and cout // parts <= 256:
            return cout // parts
    return 128


class Taps(list):
    """List of (dy, dx) taps in launch order plus ``glen``: lengths of the consecutive tap groups whose input pixel
    sets coincide up to a one-pixel shift (same dy, dx congruent modulo the input stride).  The kernel walks
    k-steps as  for group: for channel chunk: for tap in group,  so the taps of a group re-read from L1 the
    activation lines the first one brought in from L2."""
    glen: List[int] = None


def _grouped(taps, rows, step: int):
    """Reorder ``taps`` (and the parallel weight slices ``rows``) into groups; returns (Taps, reordered rows)."""
    keys = {}
    for i, (dy, dx) in enumerate(taps):
        keys.setdefault((dy, dx % step), []).append(i)
    order, glen = [], []
    for _, idxs in sorted(keys.items()):
        idxs = sorted(idxs, key=lambda i: taps[i][1])
        order += idxs
        glen.append(len(idxs))
    out = Taps(taps[i] for i in order)
    out.glen = glen
    return out, [rows[i] for i in order]


_SCHED_CACHE = {}


def _group_schedule(glen, kc: int, device) -> Tensor:
    """Index tensor of the "group" k-step order (cached per shape and device: building it from a Python list costs a
    synchronous host-to-device copy, which a training step would pay for every layer, every step)."""
    key = (glen, kc, str(device))
    idx = _SCHED_CACHE.get(key)
    if idx is None:
        sched, t0 = [], 0
        for g in glen:
            sched += [t * kc + c for c in range(kc) for t in range(t0, t0 + g)]
            t0 += g
        idx = _SCHED_CACHE[key] = torch.tensor(sched, device=device)
    return idx


def pack_weights(w_taps: Tensor, bn: int, glen=None, order: str = "group") -> Tensor:
    """w_taps fp32 [T, Cout, Cin] -> uint8 blob [n_tiles][k-step][hi | lo][BN x 32 bf16] in the UMMA canonical
    K-major order: element (r, k) of a tile at ((k // 8) * BN * 16 + (r // 8) * 128 + (r % 8) * 16 + (k % 8) * 2).
    k-step order ``"group"`` (per-tile kernel, conv.cu): for tap group (``glen``, default one tap per group): for
    channel chunk: for tap in group.  ``"chunk"`` (persistent TMA kernel, conv_tma.cu): for channel chunk: for tap
    (taps already sorted by group)."""
    T, cout, cin = w_taps.shape
    nt = (cout + bn - 1) // bn
    kc = (cin + _BK - 1) // _BK
    wp = torch.zeros((T, nt * bn, kc * _BK), dtype=torch.float32, device=w_taps.device)
    wp[:, :cout, :cin] = w_taps
    wp = wp.view(T, nt, bn // 8, 8, kc, _BK // 8, 8).permute(1, 0, 4, 5, 2, 3, 6)  # nt, T, kc, k8, r8, r%8, k%8
    hi = wp.to(torch.bfloat16)
    lo = (wp - hi.float()).to(torch.bfloat16)
    packed = torch.stack([hi, lo], dim=3)  # nt, T, kc, 2, k8, r8, r%8, k%8
    if order == "chunk":
        if T > 1 and kc > 1:
            packed = packed.transpose(1, 2)  # nt, kc, T, ...
    elif glen is not None and any(g > 1 for g in glen):
        assert sum(glen) == T
        flat = packed.reshape(nt, T * kc, *packed.shape[3:])
        packed = flat[:, _group_schedule(tuple(glen), kc, flat.device)]
    return packed.contiguous().view(torch.uint8).reshape(-1)


class _Layer:
    """One GEMM launch family prepared from a module (cached on the module, keyed by parameter versions)."""

    def __init__(self, kind, packed, bias, taps_per_phase, bn, cin, cout, geom, packed_c=None):
        self.kind, self.packed, self.bias, self.phases, self.bn = kind, packed, bias, taps_per_phase, bn
        self.cin, self.cout, self.geom = cin, cout, geom
        self.true_cout = cout  # un-padded output channels (set by the builders)
        # the same weights with k-steps in the persistent TMA kernel's order; None = not built (the layer then stays on
        # the per-tile kernel).  Single-tap layers share one order; multi-tap layers build it only under
        # TMA_POLICY == "all" (a training step repacks every layer's weights every step)
        self.packed_c = packed_c


def _prep_conv(m: "Conv2d") -> _Layer:
    """Packed layer of a conv module; cached per parameter version and published across streams (_cache.py)."""
    return _cache.cached(m, "packed", _cache.tensor_key(m.weight, m.bias), lambda: _build_conv(m), m.weight.device)


def _build_conv(m: "Conv2d") -> _Layer:
    k, s, p = m.kernel_size, m.stride, m.padding
    w = m.weight.detach().float()
    cout, cin = w.shape[0], w.shape[1]
    if cin <= 4:  # im2col path: one 1x1 GEMM over K = k*k*cin (padded to a multiple of 8)
        kflat = k * k * cin
        kpad = (kflat + 15) // 16 * 16
        wt = w.permute(0, 2, 3, 1).reshape(1, cout, kflat)  # k index = (ky*k + kx)*cin + ci
        cp = _c16(cout)
        bn = _choose_bn(cp)
        blob = pack_weights(_pad_taps(wt, cp, kflat), bn)
        lay = _Layer("conv_im2col", blob, _pad_vec(m.bias, cp), [[(0, 0)]], bn, kpad, cp, (k, s, p, kpad), packed_c=blob)
    else:
        wt = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
        taps = [(ky - p, kx - p) for ky in range(k) for kx in range(k)]
        taps, rows = _grouped(taps, list(wt.unbind(0)), s)
        wt = torch.stack(rows, 0)
        cp, kp = _c16(cout), _c16(cin)
        bn = _choose_bn(cp)
        wpad = _pad_taps(wt, cp, kp)
        lay = _Layer("conv", pack_weights(wpad, bn, taps.glen), _pad_vec(m.bias, cp), [taps], bn, kp, cp,
                     (k, s, p), packed_c=pack_weights(wpad, bn, taps.glen, order="chunk") if TMA_POLICY == "all" else None)
    lay.true_cout = cout
    return lay


def _prep_deconv(m: "ConvTranspose2d") -> _Layer:
    return _cache.cached(m, "packed", _cache.tensor_key(m.weight, m.bias), lambda: _build_deconv(m), m.weight.device)


def _build_deconv(m: "ConvTranspose2d") -> _Layer:
    k, s, p, op = m.kernel_size, m.stride, m.padding, m.output_padding
    w = m.weight.detach().float()
    cin, cout = w.shape[0], w.shape[1]
    if cout <= 4:  # col2im path: 1x1 GEMM to N = k*k*cout columns, then a gather
        nflat = k * k * cout
        npad = (nflat + 15) // 16 * 16
        wt = w.permute(2, 3, 1, 0).reshape(1, nflat, cin)  # row index = (ky*k + kx)*cout + co
        blob = pack_weights(_pad_taps(wt, npad, _c16(cin)), npad)
        lay = _Layer("deconv_col2im", blob, m.bias.detach().float().contiguous(), [[(0, 0)]], npad, _c16(cin), cout,
                     (k, s, p, op, npad), packed_c=blob)
    else:
        cp, kp = _c16(cout), _c16(cin)
        bn = _choose_bn(cp)
        phases, blobs, blobs_c = [], [], []
        for py in range(s):
            for px in range(s):
                taps, ws = [], []
                for ky in range(k):
                    if (py + p - ky) % s:
                        continue
                    for kx in range(k):
                        if (px + p - kx) % s:
                            continue
                        taps.append(((py + p - ky) // s, (px + p - kx) // s))
                        ws.append(w[:, :, ky, kx].t())  # [cout, cin]
                if ws:
                    taps, ws = _grouped(taps, ws, 1)
                phases.append(taps)
                wpad = _pad_taps(torch.stack(ws, 0), cp, kp) if ws else None
                blobs.append(pack_weights(wpad, bn, taps.glen) if ws else None)
                blobs_c.append(pack_weights(wpad, bn, taps.glen, order="chunk") if (ws and TMA_POLICY == "all") else None)
        lay = _Layer("deconv", blobs, _pad_vec(m.bias, cp), phases, bn, kp, cp, (k, s, p, op), packed_c=blobs_c)
    lay.true_cout = cout
    return lay


def _prep_gdn(m) -> _Layer:
    return _cache.cached(m, "packed", _cache.tensor_key(m.beta, m.gamma), lambda: _build_gdn(m), m.beta.device)


def _build_gdn(m) -> _Layer:
    with torch.no_grad():
        beta, gamma = m.effective_params()
    C = gamma.shape[0]
    cp = _c16(C)
    bn = _choose_bn(cp)
    # padded channels: gamma = 0, beta = 1 -> out = 0 * rsqrt(1) = 0
    lay = _Layer("igdn" if m.inverse else "gdn", pack_weights(_pad_taps(gamma.detach().float().reshape(1, C, C), cp, cp), bn),
                 _pad_vec(beta, cp, 1.0), [[(0, 0)]], bn, cp, cp, None)
    return lay


# ---- launches -------------------------------------------------------------------------------------------------
def _launch(a: Planes, packed, bias, taps, bn, cout, Ho, Wo, Hp, Wp, os_, o0y, o0x, is_, epilogue=0, aux: Planes = None,
            out_f32=None, out: Planes = None, sq: Planes = None, ab: Planes = None, clamp=None, gdn=None, packed_c=None):
    d = ConvDesc()
    if gdn is not None:  # fused GDN / IGDN: (packed gamma, beta, mode)
        d.gdn_w, d.gdn_beta, d.gdn_mode = gdn[0].data_ptr(), gdn[1].data_ptr(), gdn[2]
    d.a_hi, d.a_lo, d.w_packed = a.hi.data_ptr(), a.lo.data_ptr(), packed.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    if aux is not None:
        d.aux_hi, d.aux_lo = aux.hi.data_ptr(), aux.lo.data_ptr()
    if out_f32 is not None:
        d.out_f32 = out_f32.data_ptr()
    if out is not None:
        d.out_hi, d.out_lo = out.hi.data_ptr(), out.lo.data_ptr()
    if sq is not None:
        d.sq_hi, d.sq_lo = sq.hi.data_ptr(), sq.lo.data_ptr()
    if ab is not None:
        d.abs_hi, d.abs_lo = ab.hi.data_ptr(), ab.lo.data_ptr()
    d.N, d.H, d.W, d.Cin, d.Ho, d.Wo, d.Cout = a.N, a.H, a.W, a.C, Ho, Wo, cout
    d.Hp, d.Wp, d.os, d.o0y, d.o0x, d.is_ = Hp, Wp, os_, o0y, o0x, is_
    d.ntaps, d.BN, d.epilogue = len(taps), bn, epilogue
    d.clamp_lo, d.clamp_hi = (clamp if clamp is not None else (0.0, 0.0))
    for t, (dy, dx) in enumerate(taps):
        d.dy[t], d.dx[t] = dy, dx
    for g, n in enumerate(getattr(taps, "glen", None) or [1] * len(taps)):
        d.glen[g] = n
    with torch.cuda.device(a.hi.device):
        # the wide layers go to the persistent TMA-fed kernel (its own k-step order of the same weights)
        want_tma = TMA_POLICY == "all" or (TMA_POLICY == "auto" and len(taps) == 1)
        if want_tma and packed_c is not None and aux is None and sq is None and lib().cai_conv_tma_eligible(d):
            d.w_packed, d.mode = packed_c.data_ptr(), 1
        if TIMING is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(lib().cai_conv_gemm(d, current_stream()), "cai_conv_gemm")
        if TIMING is not None:
            e1.record()
            TIMING.setdefault("conv_gemm_kernel", []).append((e0, e1))
            if DETAIL is not None:
                DETAIL.append((f"gemm {a.C}->{cout} taps={len(taps)} grid={a.N}x{Hp}x{Wp} in={a.H}x{a.W} is={is_} "
                               f"{'gdn' if gdn is not None else 'epi%d' % epilogue}{' tma' if d.mode else ''}", e0, e1))


_ACT = {None: 0, "relu": 1, "leaky": 2}


def _outputs(N, Ho, Wo, C, device, want):
    out_f32 = torch.empty((N, Ho, Wo, C), dtype=torch.float32, device=device) if "f32" in want else None
    out = Planes.empty(N, Ho, Wo, C, device) if "planes" in want else None
    sq = Planes.empty(N, Ho, Wo, C, device) if "sq" in want else None
    ab = Planes.empty(N, Ho, Wo, C, device) if "abs" in want else None
    return out_f32, out, sq, ab


def _run_conv(m, x, act, want, clamp=None, gdn=None):
    """x: Planes (or fp32 tensor for the im2col first layer). Returns (f32 NHWC tensor | None, planes, sq, abs)."""
    lay = m if isinstance(m, _Layer) else _prep_conv(m)
    dev = lay.bias.device
    if lay.kind == "conv_im2col":
        k, s, p, kpad = lay.geom
        if isinstance(x, Planes):
            raise _lib.CaiError("a conv with at most 4 input channels must be the first layer of a stack")
        xt = x.detach().float()
        N, C, H, W = xt.shape
        layout = CAI_LAYOUT_NCHW
        if not xt.is_contiguous():
            if xt.is_contiguous(memory_format=torch.channels_last):
                layout = CAI_LAYOUT_NHWC
            else:
                xt = xt.contiguous()
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        a = Planes.empty(N, Ho, Wo, kpad, dev)
        with torch.cuda.device(dev):
            check(lib().cai_im2col_split(ptr(xt), layout, N, C, H, W, Ho, Wo, k, s, p, kpad, ptr(a.hi), ptr(a.lo),
                                         current_stream()), "cai_im2col_split")
        o = _outputs(N, Ho, Wo, lay.cout, dev, want)
        _launch(a, lay.packed, lay.bias, [(0, 0)], lay.bn, lay.cout, Ho, Wo, Ho, Wo, 1, 0, 0, 1, _ACT[act], None, *o,
                clamp=clamp, gdn=gdn, packed_c=lay.packed_c)
        return o
    if not isinstance(x, Planes):
        x = to_planes(x, lay.cin)
    k, s, p = lay.geom
    Ho, Wo = (x.H + 2 * p - k) // s + 1, (x.W + 2 * p - k) // s + 1
    o = _outputs(x.N, Ho, Wo, lay.cout, dev, want)
    _launch(x, lay.packed, lay.bias, lay.phases[0], lay.bn, lay.cout, Ho, Wo, Ho, Wo, 1, 0, 0, s, _ACT[act], None, *o,
            clamp=clamp, gdn=gdn, packed_c=lay.packed_c)
    return o


def _run_deconv(m, x, act, want, clamp=None, final_layout_nchw=False, gdn=None, dst=None):
    lay = m if isinstance(m, _Layer) else _prep_deconv(m)
    dev = lay.bias.device
    if not isinstance(x, Planes):
        x = to_planes(x, lay.cin)
    if lay.kind == "deconv_col2im":
        k, s, p, op, npad = lay.geom
        Ho, Wo = (x.H - 1) * s - 2 * p + k + op, (x.W - 1) * s - 2 * p + k + op
        cols = torch.empty((x.N, x.H, x.W, npad), dtype=torch.float32, device=dev)
        _launch(x, lay.packed, None, [(0, 0)], lay.bn, npad, x.H, x.W, x.H, x.W, 1, 0, 0, 1, 0, None, cols, None, None,
                None, packed_c=lay.packed_c)
        if final_layout_nchw:
            if dst is not None:  # caller-provided destination (a contiguous slice of a batch tensor)
                if tuple(dst.shape) != (x.N, lay.cout, Ho, Wo) or not dst.is_contiguous() or dst.dtype != torch.float32:
                    raise _lib.CaiError(f"out must be a contiguous float32 tensor of shape {(x.N, lay.cout, Ho, Wo)}")
                out = dst
            else:
                out = torch.empty((x.N, lay.cout, Ho, Wo), dtype=torch.float32, device=dev)
            layout = CAI_LAYOUT_NCHW
        else:
            out = torch.empty((x.N, Ho, Wo, lay.cout), dtype=torch.float32, device=dev)
            layout = CAI_LAYOUT_NHWC
        lo, hi = clamp if clamp is not None else (0.0, 0.0)
        with torch.cuda.device(dev):
            check(lib().cai_col2im(ptr(cols), ptr(lay.bias), x.N, lay.cout, x.H, x.W, Ho, Wo, k, s, p, npad, layout, lo,
                                   hi, ptr(out), current_stream()), "cai_col2im")
        if act == "relu":
            out.clamp_(min=0)
        return out, None, None, None
    k, s, p, op = lay.geom
    Ho, Wo = (x.H - 1) * s - 2 * p + k + op, (x.W - 1) * s - 2 * p + k + op
    o = _outputs(x.N, Ho, Wo, lay.cout, dev, want)
    i = 0
    for py in range(s):
        for px in range(s):
            taps, blob, blob_c = lay.phases[i], lay.packed[i], lay.packed_c[i]
            i += 1
            Hp, Wp = (Ho - py + s - 1) // s, (Wo - px + s - 1) // s
            if Hp <= 0 or Wp <= 0:
                continue
            if not taps:
                raise _lib.CaiError("transposed convolution phase without taps is not supported (kernel < stride)")
            _launch(x, blob, lay.bias, taps, lay.bn, lay.cout, Ho, Wo, Hp, Wp, s, py, px, 1, _ACT[act], None, *o,
                    clamp=clamp, gdn=gdn, packed_c=blob_c)
    return o


def _run_gdn(m, x: Planes, sq: Planes, want):
    lay = _prep_gdn(m)
    o = _outputs(x.N, x.H, x.W, lay.cout, x.hi.device, want)
    _launch(sq, lay.packed, lay.bias, [(0, 0)], lay.bn, lay.cout, x.H, x.W, x.H, x.W, 1, 0, 0, 1,
            4 if lay.kind == "igdn" else 3, x, *o)
    return o


def run_stack(mods: List[nn.Module], x, want_abs: bool = False, clamp=None, nchw_out: bool = False, out=None):
    """Run a conv / GDN / activation stack on the fused kernels (inference).  ``x`` is an fp32 tensor (logical
    NCHW, any memory format) or ``Planes``.  Returns the fp32 output as a logical-NCHW tensor stored channels-last
    (or true NCHW when the last layer is the 3-channel col2im with ``nchw_out``); with ``want_abs`` also returns
    the split planes of |output| for the following hyper-analysis stack.  ``out`` (only with ``nchw_out`` and a
    3-channel last layer): preallocated contiguous NCHW destination the last layer writes into."""
    from .layers.gdn import GDN

    i, n = 0, len(mods)
    cur = x
    result = None
    is_nchw = False
    while i < n:
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < n else None
        if isinstance(m, (Conv2d, ConvTranspose2d)):
            act = None
            consumed = 1
            if isinstance(nxt, nn.ReLU):
                act, consumed = "relu", 2
            elif isinstance(nxt, nn.LeakyReLU):
                act, consumed = "leaky", 2
            after = mods[i + consumed] if i + consumed < n else None
            last = after is None
            fused = None
            if isinstance(after, GDN) and act is None:
                glay = _prep_gdn(after)
                width = _prep_conv(m).cout if isinstance(m, Conv2d) else _prep_deconv(m).cout
                kind_ok = not (isinstance(m, ConvTranspose2d) and m.out_channels <= 4)
                if kind_ok and glay.cout == width and width <= 256 and glay.bn == width:
                    fused = (glay.packed, glay.bias, 2 if glay.kind == "igdn" else 1)
            if fused is not None:
                # conv + GDN in ONE launch (second in-kernel GEMM): the pre-GDN activation never leaves the SM
                glast = i + consumed + 1 >= n
                want = (("f32", "abs") if want_abs else ("f32",)) if glast else ("planes",)
                if isinstance(m, Conv2d):
                    o = _run_conv(m, cur, None, want, clamp if glast else None, gdn=fused)
                else:
                    o = _run_deconv(m, cur, None, want, clamp if glast else None, gdn=fused)
                i += consumed + 1
                if glast:
                    result = (o[0], o[3])
                else:
                    cur = o[1]
                continue
            if isinstance(after, GDN):
                want = ("planes", "sq")
            elif last:
                want = ("f32", "abs") if want_abs else ("f32",)
            else:
                want = ("planes",)
            if isinstance(m, Conv2d):
                o = _run_conv(m, cur, act, want, clamp if last else None)
            else:
                o = _run_deconv(m, cur, act, want, clamp if last else None, final_layout_nchw=nchw_out and last,
                                dst=out if (nchw_out and last) else None)
                is_nchw = bool(nchw_out and last and m.out_channels <= 4)
            i += consumed
            if isinstance(after, GDN):
                glast = i + 1 >= n
                go = _run_gdn(after, o[1], o[2], ("f32",) if glast else ("planes",))
                i += 1
                if glast:
                    result = (go[0], None)
                else:
                    cur = go[1]
            elif last:
                result = (o[0], o[3])
            else:
                cur = o[1]
        elif isinstance(m, GDN):
            # GDN not preceded by a conv of this stack: build its operands from the fp32 input
            xt = cur if not isinstance(cur, Planes) else None
            if xt is None:
                raise _lib.CaiError("standalone GDN needs an fp32 tensor input")
            cpad = _c16(xt.size(1))
            xp = to_planes(xt, cpad)
            sq = to_planes(xt.detach().float() ** 2, cpad)
            glast = i + 1 >= n
            go = _run_gdn(m, xp, sq, ("f32",) if glast else ("planes",))
            i += 1
            if glast:
                result = (go[0], None)
            else:
                cur = go[1]
        else:
            raise _lib.CaiError(f"run_stack: unsupported module {type(m).__name__}")
    out, ab = result
    if not is_nchw:
        out = out.permute(0, 3, 1, 2)  # NHWC storage -> logical NCHW (channels_last strides)
        true_c = _true_out_channels(mods)
        if true_c is not None and true_c != out.size(1):
            out = out[:, :true_c]  # drop the zero channels added to reach the tensor-core granularity
    return (out, ab) if want_abs else out


def _true_out_channels(mods):
    from .layers.gdn import GDN

    for m in reversed(mods):
        if isinstance(m, (Conv2d, ConvTranspose2d)):
            return m.out_channels
        if isinstance(m, GDN):
            return int(m.beta.numel())
    return None


def _nhwc_padded(x: Tensor, cp: int) -> Tensor:
    """logical NCHW fp32 -> contiguous [N, H, W, cp] (zero padded channels)."""
    xn = x.detach().float().permute(0, 2, 3, 1)
    if cp != xn.size(3):
        xn = F.pad(xn, (0, cp - xn.size(3)))
    return xn.contiguous()


def _planes_of(xn: Tensor) -> Planes:
    N, H, W, C = xn.shape
    out = Planes.empty(N, H, W, C, xn.device)
    with torch.cuda.device(xn.device):
        check(lib().cai_split_planes(ptr(xn), CAI_LAYOUT_NHWC, N, C, H * W, C, ptr(out.hi), ptr(out.lo), current_stream()),
              "cai_split_planes")
    return out


# ---- training mode: conv / transposed conv with all three GEMMs on this package's kernels ---------------------
def _run_conv_layer(lay: "_Layer", x: Tensor) -> Tensor:
    """One prepared conv layer on an fp32 tensor -> logical-NCHW fp32 result (true channel count)."""
    o = _run_conv(lay, x, None, ("f32",))
    return o[0].permute(0, 3, 1, 2)[:, :lay.true_cout]


def _run_deconv_layer(lay: "_Layer", x: Tensor) -> Tensor:
    o = _run_deconv(lay, x, None, ("f32",), final_layout_nchw=True)
    out = o[0]
    if lay.kind != "deconv_col2im":
        out = out.permute(0, 3, 1, 2)
    return out[:, :lay.true_cout]


class _LayerShim:
    """The attributes ``_build_conv`` / ``_build_deconv`` read from a module: lets the data gradient reuse the forward
    kernels with the SAME weight tensor (a conv's weight [Cout, Cin, k, k] is a transposed conv's [in, out, k, k] with
    in = Cout, and vice versa) and a zero bias."""

    def __init__(self, weight, in_channels, out_channels, k, s, p, op=0):
        self.weight, self.in_channels, self.out_channels = weight, in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.output_padding = k, s, p, op
        self.bias = torch.zeros(out_channels, dtype=torch.float32, device=weight.device)


def conv_wgrad(small: Tensor, big: Tensor, k: int, s: int, p: int) -> Tensor:
    """grad[cs, cb, ky, kx] = sum_{n,i,j} small[n,cs,i,j] * big[n,cb, s*i+ky-p, s*j+kx-p]  (``cai_conv_wgrad``)."""
    require_cuda(small, "gradients")
    small = small.detach().float().contiguous()
    big = big.detach().float().contiguous()
    N, Cs, Hs, Ws = small.shape
    Nb, Cb, Hb, Wb = big.shape
    if N != Nb:
        raise _lib.CaiError("conv_wgrad: batch sizes differ")
    dev = small.device
    with torch.cuda.device(dev):
        need = lib().cai_conv_wgrad_workspace(N, Cs, Hs, Ws, Cb, Hb, Wb, k, s)
        if need < 0:
            check(-1, "cai_conv_wgrad_workspace")
        ws = torch.empty(int(need) + 256, dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        grad = torch.empty((Cs, Cb, k, k), dtype=torch.float32, device=dev)
        check(lib().cai_conv_wgrad(ptr(small), ptr(big), N, Cs, Hs, Ws, Cb, Hb, Wb, k, s, p, ptr(grad),
                                   _lib.c_void_p(ws.data_ptr() + off), int(need), current_stream()), "cai_conv_wgrad")
    return grad


class _ConvFunction(torch.autograd.Function):
    """``Conv2d`` / ``ConvTranspose2d`` in training mode (autograd of compressai/models/utils.py:128-146 layers as
    driven by examples/train.py:132-165).  forward: the inference kernel.  backward: dX = the other layer kind's
    forward on dY with the same weights; dW = ``cai_conv_wgrad``; db = per-channel sum of dY."""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        require_cuda(x, "inputs")
        with torch.no_grad():
            out = run_stack([mod], x)
        ctx.save_for_backward(x, weight)
        ctx.mod = mod
        # run_stack returns a permuted view of its NHWC buffer; an in-place op on a view created inside a custom
        # Function (nn.ReLU(inplace=True) follows the hyper-transform layers) is refused by autograd, so hand out a
        # plain tensor over the same storage instead
        res = torch.empty(0, dtype=out.dtype, device=out.device)
        res.set_(out.untyped_storage(), out.storage_offset(), out.size(), out.stride())
        return res

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        m = ctx.mod
        k, s, p = m.kernel_size, m.stride, m.padding
        transposed = isinstance(m, ConvTranspose2d)
        g = g.detach()
        gx = gw = gb = None
        with torch.no_grad():
            if ctx.needs_input_grad[0]:
                if transposed:   # dX = conv2d(dY, W): W [Cin, Cout, k, k] read as a conv weight [out = Cin, in = Cout]
                    shim = _LayerShim(weight.detach(), m.out_channels, m.in_channels, k, s, p)
                    gx = _run_conv_layer(_build_conv(shim), g)
                else:            # dX = conv_transpose2d(dY, W): W [Cout, Cin, k, k] read as [in = Cout, out = Cin]
                    op = x.size(2) - ((g.size(2) - 1) * s - 2 * p + k)
                    opw = x.size(3) - ((g.size(3) - 1) * s - 2 * p + k)
                    if op != opw or not 0 <= op < max(s, 1):
                        raise _lib.CaiError("conv backward: input size is not reachable by a transposed convolution")
                    shim = _LayerShim(weight.detach(), m.out_channels, m.in_channels, k, s, p, op)
                    gx = _run_deconv_layer(_build_deconv(shim), g)
            if ctx.needs_input_grad[1]:
                gw = conv_wgrad(x, g, k, s, p) if transposed else conv_wgrad(g, x, k, s, p)
            if ctx.needs_input_grad[2]:
                gb = g.float().sum(dim=(0, 2, 3))
        return gx, gw, gb, None


class _GDNFunction(torch.autograd.Function):
    """GDN / IGDN with both directions on this package's kernels (training mode): forward = 1x1 tcgen05 GEMM of x^2
    with gamma + fused finalize; backward = norm GEMM, elementwise stage, gamma^T GEMM, elementwise stage, and the
    pixel-reduction outer product for dgamma / dbeta (cai_gdn_bwd_*; formulas SURVEY.md Appendix D.1)."""

    @staticmethod
    def forward(ctx, x, beta, gamma, inverse):
        require_cuda(x, "inputs")
        C = x.size(1)
        cp = _c16(C)
        bn = _choose_bn(cp)
        if bn != cp:
            raise _lib.CaiError("GDN with more than 256 channels is not supported by the fused kernel")
        xn = _nhwc_padded(x, cp)
        N, H, W, _ = xn.shape
        g_packed = pack_weights(_pad_taps(gamma.detach().float().reshape(1, C, C), cp, cp), bn)
        b_pad = _pad_vec(beta, cp, 1.0)
        xp, sq = _planes_of(xn), _planes_of(xn * xn)
        out = torch.empty((N, H, W, cp), dtype=torch.float32, device=xn.device)
        _launch(sq, g_packed, b_pad, [(0, 0)], bn, cp, H, W, H, W, 1, 0, 0, 1, 4 if inverse else 3, xp, out, None, None,
                None)
        ctx.save_for_backward(xn, gamma.detach(), beta.detach())
        ctx.cfg = (bool(inverse), C, cp, bn)
        return out[..., :C].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        xn, gamma, beta = ctx.saved_tensors
        inverse, C, cp, bn = ctx.cfg
        N, H, W, _ = xn.shape
        dev = xn.device
        gn = _nhwc_padded(g, cp)
        gam = _pad_taps(gamma.float().reshape(1, C, C), cp, cp)
        b_pad = _pad_vec(beta, cp, 1.0)
        sq = _planes_of(xn * xn)
        norm = torch.empty_like(xn)
        _launch(sq, pack_weights(gam, bn), b_pad, [(0, 0)], bn, cp, H, W, H, W, 1, 0, 0, 1, 0, None, norm, None, None, None)
        t = torch.empty_like(xn)
        p = torch.empty_like(xn)
        tp = Planes.empty(N, H, W, cp, dev)
        n_el = xn.numel()
        with torch.cuda.device(dev):
            check(lib().cai_gdn_bwd_prepare(ptr(xn), ptr(norm), ptr(gn), int(inverse), n_el, ptr(t), ptr(tp.hi), ptr(tp.lo),
                                            ptr(p), current_stream()), "cai_gdn_bwd_prepare")
        u = torch.empty_like(xn)
        _launch(tp, pack_weights(gam.transpose(1, 2).contiguous(), bn), None, [(0, 0)], bn, cp, H, W, H, W, 1, 0, 0, 1, 0,
                None, u, None, None, None)
        gx = torch.empty_like(xn)
        with torch.cuda.device(dev):
            check(lib().cai_gdn_bwd_finish(ptr(p), ptr(xn), ptr(u), int(inverse), n_el, ptr(gx), current_stream()),
                  "cai_gdn_bwd_finish")
        # dgamma[i, j] = -+1/2 sum_pixels t_i x_j^2 is the weight gradient of a 1x1 convolution: the same tcgen05
        # pixel-reduction GEMM as the conv layers (cai_conv_wgrad); dbeta_i = -+1/2 sum_pixels t_i
        half = 0.5 if inverse else -0.5
        g_gamma = conv_wgrad(t.permute(0, 3, 1, 2), (xn * xn).permute(0, 3, 1, 2), 1, 1, 0).reshape(cp, cp) * half
        g_beta = t.sum(dim=(0, 1, 2)) * half
        return gx[..., :C].permute(0, 3, 1, 2), g_beta[:C], g_gamma[:C, :C], None


def gdn(x, beta, gamma, inverse):
    """Functional GDN used by ``layers.GDN.forward`` in training mode: forward and backward on our kernels."""
    return _GDNFunction.apply(x, beta, gamma, bool(inverse))
