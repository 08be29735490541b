"""Cheng2020 models with the reference's interface (compressai/models/waseda.py:44-153): ``Cheng2020Anchor`` and
``Cheng2020Attention`` -- residual blocks of 3x3 / 1x1 convolutions, sub-pixel up-sampling and (attention variant)
simplified attention blocks around the joint autoregressive + hierarchical entropy model.  Same submodule names and
``state_dict`` keys.  ``compress`` / ``decompress`` / ``forward`` are inherited from
``JointAutoregressiveHierarchicalPriors``: the context-model scan is the cluster kernel of csrc/ar.cu, every
convolution the tcgen05 implicit-GEMM kernel (layers/layers.py)."""
import torch.nn as nn

from ..layers import (AttentionBlock, ResidualBlock, ResidualBlockUpsample, ResidualBlockWithStride, conv3x3,
                      subpel_conv3x3)
from ..transforms import TransformStack
from .google import JointAutoregressiveHierarchicalPriors

__all__ = ["Cheng2020Anchor", "Cheng2020Attention"]


class _Blocks(nn.Sequential):
    """Sequential of blocks; tolerates the keyword arguments the conv stacks of the other models accept."""

    def forward(self, x, clamp=None, nchw_out=False, **kw):
        for m in self:
            x = m(x)
        if clamp is not None:
            x = x.clamp(*clamp)
        return x.contiguous() if nchw_out else x


class Cheng2020Anchor(JointAutoregressiveHierarchicalPriors):
    """Anchor variant of Cheng et al., CVPR 2020.  Args: N (channels)."""

    def __init__(self, N=192, **kwargs):
        super().__init__(N=N, M=N, **kwargs)
        self.g_a = _Blocks(ResidualBlockWithStride(3, N, stride=2), ResidualBlock(N, N),
                           ResidualBlockWithStride(N, N, stride=2), ResidualBlock(N, N),
                           ResidualBlockWithStride(N, N, stride=2), ResidualBlock(N, N), conv3x3(N, N, stride=2))
        self.h_a = TransformStack(conv3x3(N, N), nn.LeakyReLU(inplace=True), conv3x3(N, N), nn.LeakyReLU(inplace=True),
                                 conv3x3(N, N, stride=2), nn.LeakyReLU(inplace=True), conv3x3(N, N),
                                 nn.LeakyReLU(inplace=True), conv3x3(N, N, stride=2))
        self.h_s = _Blocks(conv3x3(N, N), nn.LeakyReLU(inplace=True), subpel_conv3x3(N, N, 2), nn.LeakyReLU(inplace=True),
                           conv3x3(N, N * 3 // 2), nn.LeakyReLU(inplace=True), subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2),
                           nn.LeakyReLU(inplace=True), conv3x3(N * 3 // 2, N * 2))
        self.g_s = _Blocks(ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
                           ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
                           ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.conv1.weight"].size(0)
        net = cls(N)
        net.load_state_dict(state_dict)
        return net


class Cheng2020Attention(Cheng2020Anchor):
    """Self-attention variant of Cheng et al., CVPR 2020.  Args: N (channels)."""

    def __init__(self, N=192, **kwargs):
        super().__init__(N=N, **kwargs)
        self.g_a = _Blocks(ResidualBlockWithStride(3, N, stride=2), ResidualBlock(N, N),
                           ResidualBlockWithStride(N, N, stride=2), AttentionBlock(N), ResidualBlock(N, N),
                           ResidualBlockWithStride(N, N, stride=2), ResidualBlock(N, N), conv3x3(N, N, stride=2),
                           AttentionBlock(N))
        self.g_s = _Blocks(AttentionBlock(N), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
                           ResidualBlockUpsample(N, N, 2), AttentionBlock(N), ResidualBlock(N, N),
                           ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))
