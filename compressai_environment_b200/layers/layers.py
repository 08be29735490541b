"""Layers of compressai/layers/layers.py that the hot path's "next" rows need.

``MaskedConv2d`` (reference :52-78), ``conv3x3`` / ``subpel_conv3x3`` / ``conv1x1`` (:81-98) and the Cheng2020 blocks
``ResidualBlockWithStride`` (:101-135), ``ResidualBlockUpsample`` (:138-168), ``ResidualBlock`` (:171-203),
``AttentionBlock`` (:206-244): same constructor arguments, submodule names and ``state_dict`` keys as the reference.
Every convolution is this package's ``transforms.Conv2d`` (tcgen05 implicit GEMM in inference, with the following
LeakyReLU / ReLU folded into the launch's epilogue; _ConvFunction in training); the residual additions, the sigmoid
gate and the pixel shuffle are elementwise glue between launches.

``QReLU`` (reference :247-296): forward clamps to the integer range of ``bit_depth``; backward passes the gradient
through inside the range and, outside it, attenuates it by exp(-alpha^beta * |2 x / max - 1|^beta) with the
pre-computed alpha of the reference (the generalised-Gaussian surrogate of "Integer networks for data compression with
latent-variable models", Balle et al., ICLR 2019).  In inference the clamp is folded into the producing conv launch's
epilogue (models/video/google.py); this autograd Function is the training-mode path.
"""
import torch
import torch.nn as nn
from torch import Tensor
from torch.autograd import Function

from ..transforms import Conv2d, run_stack
from .gdn import GDN

__all__ = ["AttentionBlock", "MaskedConv2d", "ResidualBlock", "ResidualBlockUpsample", "ResidualBlockWithStride",
           "conv3x3", "subpel_conv3x3", "conv1x1", "QReLU"]


def _inference(x: Tensor, *mods) -> bool:
    return not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for m in mods for p in m.parameters())))


def _conv_act(conv: Conv2d, act, x: Tensor) -> Tensor:
    """conv followed by an activation module: one launch in inference (activation in the epilogue)."""
    if _inference(x, conv):
        return run_stack([conv, act], x)
    return act(conv(x))


class MaskedConv2d(Conv2d):
    """Masked convolution for autoregressive context models: type "A" hides the current pixel and everything after it
    in raster order, type "B" keeps the current pixel.  The mask is a buffer (``state_dict`` key ``mask``) multiplied
    into ``weight.data`` on every forward, as in the reference."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, padding=None, mask_type: str = "A"):
        if mask_type != "A" and mask_type != "B":
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        super().__init__(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        k = self.kernel_size
        causal = torch.zeros(k * k)
        causal[: (k // 2) * k + k // 2 + (1 if mask_type == "B" else 0)] = 1  # raster-order prefix of the window
        self.register_buffer("mask", causal.view(1, 1, k, k).expand_as(self.weight).clone())

    def forward(self, x: Tensor) -> Tensor:
        with torch.no_grad():
            self.weight.mul_(self.mask)  # bumps the version counter: the packed-weight cache sees the change
        return super().forward(x)


def conv3x3(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch: int, out_ch: int, r: int = 1) -> nn.Sequential:
    """3x3 convolution to out_ch * r^2 channels + pixel shuffle (sub-pixel up-sampling)."""
    return nn.Sequential(Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, stride=1, padding=1), nn.PixelShuffle(r))


def conv1x1(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return Conv2d(in_ch, out_ch, kernel_size=1, stride=stride, padding=0)


def _conv_gdn(conv: Conv2d, gdn: GDN, x: Tensor) -> Tensor:
    """conv followed by GDN / IGDN: one launch in inference (second in-kernel GEMM)."""
    if _inference(x, conv, gdn):
        return run_stack([conv, gdn], x)
    return gdn(conv(x))


class ResidualBlockWithStride(nn.Module):
    """conv3x3(stride) - LeakyReLU - conv3x3 - GDN, plus a 1x1 strided skip (reference layers.py:101-135)."""

    def __init__(self, in_ch: int, out_ch: int, stride: int = 2):
        super().__init__()
        needs_skip = stride != 1 or in_ch != out_ch
        self.conv1, self.leaky_relu = conv3x3(in_ch, out_ch, stride=stride), nn.LeakyReLU(inplace=True)
        self.conv2, self.gdn = conv3x3(out_ch, out_ch), GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if needs_skip else None

    def forward(self, x: Tensor) -> Tensor:
        main = _conv_gdn(self.conv2, self.gdn, _conv_act(self.conv1, self.leaky_relu, x))
        return main + (x if self.skip is None else self.skip(x))


class ResidualBlockUpsample(nn.Module):
    """subpel conv - LeakyReLU - conv3x3 - IGDN, plus a sub-pixel skip (reference layers.py:138-168)."""

    def __init__(self, in_ch: int, out_ch: int, upsample: int = 2):
        super().__init__()
        self.subpel_conv, self.leaky_relu = subpel_conv3x3(in_ch, out_ch, upsample), nn.LeakyReLU(inplace=True)
        self.conv, self.igdn = conv3x3(out_ch, out_ch), GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def forward(self, x: Tensor) -> Tensor:
        conv, shuffle = self.subpel_conv
        # LeakyReLU commutes with the pixel shuffle (a permutation): fold it into the conv launch
        main = _conv_gdn(self.conv, self.igdn, shuffle(_conv_act(conv, self.leaky_relu, x)))
        return main + self.upsample(x)


class ResidualBlock(nn.Module):
    """Two conv3x3 + LeakyReLU with an identity (or 1x1) skip (reference layers.py:171-203)."""

    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.conv1, self.leaky_relu = conv3x3(in_ch, out_ch), nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = None if in_ch == out_ch else conv1x1(in_ch, out_ch)

    def forward(self, x: Tensor) -> Tensor:
        main = _conv_act(self.conv2, self.leaky_relu, _conv_act(self.conv1, self.leaky_relu, x))
        return main + (x if self.skip is None else self.skip(x))


class _ResidualUnit(nn.Module):
    """1x1 - ReLU - 3x3 - ReLU - 1x1 bottleneck with an identity skip (the unit of ``AttentionBlock``)."""

    def __init__(self, N: int):
        super().__init__()
        half = N // 2
        self.conv = nn.Sequential(conv1x1(N, half), nn.ReLU(inplace=True), conv3x3(half, half), nn.ReLU(inplace=True),
                                  conv1x1(half, N))
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x: Tensor) -> Tensor:
        # inference: three launches, activations stay in split planes between them
        y = run_stack(list(self.conv), x) if _inference(x, self.conv) else self.conv(x)
        return self.relu(y + x)


class AttentionBlock(nn.Module):
    """Simplified self-attention block of Cheng et al. 2020: out = a(x) * sigmoid(b(x)) + x (reference layers.py:206-244)."""

    def __init__(self, N: int):
        super().__init__()
        self.conv_a = nn.Sequential(*(_ResidualUnit(N) for _ in range(3)))
        self.conv_b = nn.Sequential(*(_ResidualUnit(N) for _ in range(3)), conv1x1(N, N))

    def forward(self, x: Tensor) -> Tensor:
        gate = torch.sigmoid(self.conv_b(x))
        return self.conv_a(x) * gate + x


_QRELU_ALPHA = 0.9943258522851727


class QReLU(Function):
    @staticmethod
    def forward(ctx, input, bit_depth, beta):
        ctx.beta = beta
        ctx.max_value = 2 ** bit_depth - 1
        ctx.save_for_backward(input)
        return input.clamp(min=0, max=ctx.max_value)

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        outside = (x < 0) | (x > ctx.max_value)
        damp = torch.exp(-(_QRELU_ALPHA ** ctx.beta) * torch.abs(2.0 * x / ctx.max_value - 1) ** ctx.beta)
        return torch.where(outside, damp * grad_output, grad_output), None, None
