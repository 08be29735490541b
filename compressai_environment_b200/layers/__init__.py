from .gdn import GDN
from .layers import (AttentionBlock, MaskedConv2d, QReLU, ResidualBlock, ResidualBlockUpsample, ResidualBlockWithStride,
                     conv1x1, conv3x3, subpel_conv3x3)

__all__ = ["GDN", "AttentionBlock", "MaskedConv2d", "QReLU", "ResidualBlock", "ResidualBlockUpsample",
           "ResidualBlockWithStride", "conv1x1", "conv3x3", "subpel_conv3x3"]
