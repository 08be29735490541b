"""Analysis / synthesis transform building blocks: ``Conv2d``, ``ConvTranspose2d`` and ``gdn``.

Reference: the ``conv`` / ``deconv`` factories (compressai/models/utils.py:128-146: k=5, s=2, pad=k//2,
output_padding=s-1) build ``nn.Conv2d`` / ``nn.ConvTranspose2d``; GDN is compressai/layers/gdn.py:77-92.
The modules here keep torch's parameter names and shapes (``weight`` [Cout, Cin, k, k] for conv,
[Cin, Cout, k, k] for transposed conv, ``bias`` [Cout]) so reference checkpoints load unchanged.
Activations are kept channels-last (NHWC) between layers: that is the layout the implicit-GEMM kernels
read (K = Cin contiguous) and the fused quantize/index kernels transpose from.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from ._lib import require_cuda


class Conv2d(nn.Module):
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
        return conv2d(x, self.weight, self.bias, self.stride, self.padding)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"


class ConvTranspose2d(nn.Module):
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
        return conv_transpose2d(x, self.weight, self.bias, self.stride, self.padding, self.output_padding)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"


# ---- compute entry points (single dispatch point for the transform kernels) ---------------------------------
def conv2d(x, weight, bias, stride, padding):
    require_cuda(x, "inputs")
    return F.conv2d(x, weight, bias, stride=stride, padding=padding)


def conv_transpose2d(x, weight, bias, stride, padding, output_padding):
    require_cuda(x, "inputs")
    return F.conv_transpose2d(x, weight, bias, stride=stride, padding=padding, output_padding=output_padding)


def gdn(x, beta, gamma, inverse):
    require_cuda(x, "inputs")
    C = x.size(1)
    norm = F.conv2d(x * x, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm
