"""The hyperprior of the scale-space-flow video model (ssf2020) on this package's kernels.

Reference: the ``Hyperprior`` class nested in ``ScaleSpaceFlow.__init__``
(compressai/models/video/google.py:150-196) with its ``HyperEncoder`` (:103-114), ``HyperDecoder`` (:116-127) and
``HyperDecoderWithQReLU`` (:129-148), and ``QReLU`` (compressai/layers/layers.py:247-296).  An inter frame codes
THREE latents through this class (image, motion and residual hyperpriors, :198-206), i.e. three times the independent
strings of an image model per step -- the same GaussianConditional-with-means / EntropyBottleneck kernels as
``MeanScaleHyperprior``.  Same constructor arguments, submodule names and ``state_dict`` keys as the reference class;
``compress`` returns ``(y_hat, {"strings": [y_strings, z_strings], "shape": ...})`` and ``decompress`` returns
``y_hat`` exactly like the reference methods.

Inference data flow: every layer is one launch of the implicit-GEMM kernel; ReLU is folded into the producing launch's
epilogue, and so is QReLU, whose forward is ``clamp(x, 0, 2**bit_depth - 1)`` (the epilogue's clamp).  The encoder
never decodes its own z strings: ``z_hat = dequantize(quantize(z))`` by construction (reference :175-176 calls
``decompress`` on the strings it just wrote).
"""
import torch
import torch.nn as nn

from ... import _lib, kernels
from ...entropy_models import GaussianConditional
from ...layers import QReLU
from ...ops import ste_round
from ...transforms import TransformStack, run_stack
from ..google import CompressionModel
from ..utils import conv, deconv, update_registered_buffers

_CL = torch.channels_last


def _nhwc(x):
    return x.contiguous(memory_format=_CL) if x.dim() == 4 else x


class HyperEncoder(TransformStack):
    def __init__(self, in_planes: int = 192, mid_planes: int = 192, out_planes: int = 192):
        super().__init__(conv(in_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, mid_planes, kernel_size=5, stride=2))


class HyperDecoder(TransformStack):
    def __init__(self, in_planes: int = 192, mid_planes: int = 192, out_planes: int = 192):
        super().__init__(deconv(in_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, out_planes, kernel_size=5, stride=2))


class HyperDecoderWithQReLU(nn.Module):
    """deconv -> QReLU three times (reference :129-148; bit_depth 8, beta 100)."""

    bit_depth, beta = 8, 100

    def __init__(self, in_planes: int = 192, mid_planes: int = 192, out_planes: int = 192):
        super().__init__()
        self.deconv1 = deconv(in_planes, mid_planes, kernel_size=5, stride=2)
        self.deconv2 = deconv(mid_planes, mid_planes, kernel_size=5, stride=2)
        self.deconv3 = deconv(mid_planes, out_planes, kernel_size=5, stride=2)

    def qrelu(self, x):
        return QReLU.apply(x, self.bit_depth, self.beta)

    def forward(self, x):
        layers = (self.deconv1, self.deconv2, self.deconv3)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            for m in layers:
                x = self.qrelu(m(x))
            return x
        hi = float(2 ** self.bit_depth - 1)
        for m in layers:  # the clamp of QReLU's forward runs in the deconv launch's epilogue
            x = run_stack([m], x, clamp=(0.0, hi))
        return x


class Hyperprior(CompressionModel):
    def __init__(self, planes: int = 192, mid_planes: int = 192):
        super().__init__(entropy_bottleneck_channels=mid_planes)
        self.hyper_encoder = HyperEncoder(planes, mid_planes, planes)
        self.hyper_decoder_mean = HyperDecoder(planes, mid_planes, planes)
        self.hyper_decoder_scale = HyperDecoderWithQReLU(planes, mid_planes, planes)
        self.gaussian_conditional = GaussianConditional(None)

    def forward(self, y):
        y = _nhwc(y)
        z = self.hyper_encoder(y)
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales = self.hyper_decoder_scale(z_hat)
        means = self.hyper_decoder_mean(z_hat)
        _, y_likelihoods = self.gaussian_conditional(y, scales, means)
        y_hat = ste_round(y - means) + means
        return y_hat, {"y": y_likelihoods, "z": z_likelihoods}

    def load_state_dict(self, state_dict):
        update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        super().load_state_dict(state_dict)

    def update(self, scale_table=None, force=False):
        from ..google import get_scale_table

        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    def _params(self, z_hat):
        return self.hyper_decoder_scale(z_hat), self.hyper_decoder_mean(z_hat)

    @torch.no_grad()
    def compress(self, y):
        _lib.require_cuda(y, "inputs")
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = _nhwc(y)
        z = self.hyper_encoder(y)
        z_strings = eb.compress(z)
        z_sym, _ = kernels.eb_quantize_index(z, eb._get_medians())
        z_hat = kernels.dequantize(z_sym, None, eb._get_medians(), tuple(z.shape), _CL)
        scales, means = self._params(z_hat)
        indexes = gc.build_indexes(scales)
        y_strings = gc.compress(y, indexes, means)
        y_hat = gc.quantize(y, "dequantize", means)
        return y_hat, {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape, memory_format=_CL)
        scales, means = self._params(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales)
        return self.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype, means)
