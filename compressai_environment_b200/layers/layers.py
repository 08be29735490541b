"""Layers of compressai/layers/layers.py that the hot path's "next" rows need.

``QReLU`` (reference :247-296): forward clamps to the integer range of ``bit_depth``; backward passes the gradient
through inside the range and, outside it, attenuates it by exp(-alpha^beta * |2 x / max - 1|^beta) with the
pre-computed alpha of the reference (the generalised-Gaussian surrogate of "Integer networks for data compression with
latent-variable models", Balle et al., ICLR 2019).  In inference the clamp is folded into the producing conv launch's
epilogue (models/video/google.py); this autograd Function is the training-mode path.
"""
import torch
from torch.autograd import Function

__all__ = ["QReLU"]

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
