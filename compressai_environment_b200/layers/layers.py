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
    in raster order, type "B" keeps the current pixel.  The mask is a buffer applied to ``weight.data`` on every
    forward, as in the reference."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, padding=None, mask_type: str = "A"):
        super().__init__(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        if mask_type not in ("A", "B"):
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x: Tensor) -> Tensor:
        self.weight.data *= self.mask
        return super().forward(x)


def conv3x3(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch: int, out_ch: int, r: int = 1) -> nn.Sequential:
    """3x3 convolution to out_ch * r^2 channels + pixel shuffle (sub-pixel up-sampling)."""
    return nn.Sequential(Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, stride=1, padding=1), nn.PixelShuffle(r))


def conv1x1(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return Conv2d(in_ch, out_ch, kernel_size=1, stride=stride, padding=0)


class ResidualBlockWithStride(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, stride: int = 2):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch, stride=stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.gdn = GDN(out_ch)
        self.skip = conv1x1(in_ch, out_ch, stride=stride) if (stride != 1 or in_ch != out_ch) else None

    def forward(self, x: Tensor) -> Tensor:
        out = _conv_act(self.conv1, self.leaky_relu, x)
        if _inference(out, self.conv2, self.gdn):
            out = run_stack([self.conv2, self.gdn], out)  # conv + GDN in one launch
        else:
            out = self.gdn(self.conv2(out))
        identity = self.skip(x) if self.skip is not None else x
        return out + identity


class ResidualBlockUpsample(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, upsample: int = 2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(in_ch, out_ch, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(out_ch, out_ch)
        self.igdn = GDN(out_ch, inverse=True)
        self.upsample = subpel_conv3x3(in_ch, out_ch, upsample)

    def forward(self, x: Tensor) -> Tensor:
        # LeakyReLU commutes with the pixel shuffle (a permutation): fold it into the conv launch
        out = self.subpel_conv[1](_conv_act(self.subpel_conv[0], self.leaky_relu, x))
        if _inference(out, self.conv, self.igdn):
            out = run_stack([self.conv, self.igdn], out)
        else:
            out = self.igdn(self.conv(out))
        return out + self.upsample(x)


class ResidualBlock(nn.Module):
    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.conv1 = conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(out_ch, out_ch)
        self.skip = conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward(self, x: Tensor) -> Tensor:
        out = _conv_act(self.conv1, self.leaky_relu, x)
        out = _conv_act(self.conv2, self.leaky_relu, out)
        identity = self.skip(x) if self.skip is not None else x
        return out + identity


class AttentionBlock(nn.Module):
    """Simplified self-attention block of Cheng et al. 2020: out = a(x) * sigmoid(b(x)) + x."""

    def __init__(self, N: int):
        super().__init__()

        class ResidualUnit(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                          nn.ReLU(inplace=True), conv1x1(N // 2, N))
                self.relu = nn.ReLU(inplace=True)

            def forward(self, x: Tensor) -> Tensor:
                if _inference(x, self.conv):
                    out = run_stack(list(self.conv), x)  # three launches, activations stay in split planes
                else:
                    out = self.conv(x)
                return self.relu(out + x)

        self.conv_a = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit())
        self.conv_b = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit(), conv1x1(N, N))

    def forward(self, x: Tensor) -> Tensor:
        a = self.conv_a(x)
        b = self.conv_b(x)
        return a * torch.sigmoid(b) + x

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
