"""The model classes on the hot path with the reference's interface (compressai/models/google.py):
``CompressionModel`` (:56-116), ``FactorizedPrior`` (:119-191), ``ScaleHyperprior`` (:204-321),
``MeanScaleHyperprior`` (:324-392) and ``get_scale_table`` (:195-201).

Same constructor arguments, submodule names (``g_a``, ``g_s``, ``h_a``, ``h_s``, ``entropy_bottleneck``,
``gaussian_conditional``) and ``state_dict`` keys, same return dictionaries.  Differences underneath:
  * activations are channels-last between layers;
  * ``compress`` never decodes its own z strings: the decoder's ``z_hat`` is by construction
    ``dequantize(quantize(z))``, which the fused quantize kernel already produced (reference :307,
    :374 call ``decompress`` on the strings it just wrote);
  * quantize + build_indexes + the whole batch's strings are one fused kernel + one coder launch.
"""
import math
import warnings

import torch
import torch.nn as nn

from ..entropy_models import EntropyBottleneck, GaussianConditional
from ..layers import GDN
from ..transforms import TransformStack
from .utils import conv, deconv, update_registered_buffers

__all__ = [
    "CompressionModel",
    "FactorizedPrior",
    "ScaleHyperprior",
    "MeanScaleHyperprior",
    "get_scale_table",
    "SCALES_MIN",
    "SCALES_MAX",
    "SCALES_LEVELS",
]

_CL = torch.channels_last


def _nhwc(x):
    return x.contiguous(memory_format=_CL) if x.dim() == 4 else x


class CompressionModel(nn.Module):
    """Base class for an auto-encoder with at least one entropy bottleneck module."""

    def __init__(self, entropy_bottleneck_channels, init_weights=None):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)
        if init_weights is not None:
            warnings.warn("init_weights was removed as it was never functional", DeprecationWarning)

    def aux_loss(self):
        """Aggregated loss over the auxiliary entropy bottleneck module(s)."""
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def forward(self, *args):
        raise NotImplementedError()

    def update(self, force=False):
        """Updates the entropy bottleneck(s) CDF values; True if one of them was updated."""
        updated = False
        for m in self.children():
            if not isinstance(m, EntropyBottleneck):
                continue
            rv = m.update(force=force)
            updated |= rv
        return updated

    def load_state_dict(self, state_dict):
        update_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck",
                                  ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        super().load_state_dict(state_dict)


class FactorizedPrior(CompressionModel):
    r"""Factorized Prior model (Ballé et al., ICLR 2018).  Args: N, M as in the reference."""

    def __init__(self, N, M, **kwargs):
        super().__init__(entropy_bottleneck_channels=M, **kwargs)
        self.g_a = TransformStack(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = TransformStack(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.N = N
        self.M = M

    @property
    def downsampling_factor(self) -> int:
        return 2**4

    def forward(self, x):
        y = self.g_a(_nhwc(x))
        y_hat, y_likelihoods = self.entropy_bottleneck(y)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods}}

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M)
        net.load_state_dict(state_dict)
        return net

    def compress(self, x):
        y = self.g_a(_nhwc(x))
        y_strings = self.entropy_bottleneck.compress(y)
        return {"strings": [y_strings], "shape": y.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape, memory_format=_CL)
        x_hat = self.g_s(y_hat).clamp_(0, 1)
        return {"x_hat": x_hat}


# From Balle's tensorflow compression examples
SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class ScaleHyperprior(CompressionModel):
    r"""Scale Hyperprior model (Ballé et al., ICLR 2018)."""

    def __init__(self, N, M, **kwargs):
        super().__init__(entropy_bottleneck_channels=N, **kwargs)
        self.g_a = TransformStack(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = TransformStack(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = TransformStack(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def _hyper_in(self, y):
        return torch.abs(y)

    def _params(self, z_hat):
        """(scales_hat, means_hat or None) from the hyper-synthesis transform."""
        return self.h_s(z_hat), None

    def forward(self, x):
        y = self.g_a(_nhwc(x))
        z = self.h_a(self._hyper_in(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat, means_hat = self._params(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat, means=means_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}

    def load_state_dict(self, state_dict):
        update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        super().load_state_dict(state_dict)

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M)
        net.load_state_dict(state_dict)
        return net

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    def compress(self, x):
        y = self.g_a(_nhwc(x))
        z = self.h_a(self._hyper_in(y))
        z_enc, z_hat = self.entropy_bottleneck.compress_symbols(z)
        scales_hat, means_hat = self._params(z_hat)
        y_enc, _ = self.gaussian_conditional.compress_from_scales(y, scales_hat, means_hat)
        return {"strings": [y_enc.to_bytes(), z_enc.to_bytes()], "shape": z.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape, memory_format=_CL)
        scales_hat, means_hat = self._params(z_hat)
        y_hat = self.gaussian_conditional.decompress_from_scales(strings[0], scales_hat, means_hat)
        x_hat = self.g_s(y_hat).clamp_(0, 1)
        return {"x_hat": x_hat}

    # ---- device-resident variants (same kernels, strings never leave HBM): used for kernel-only timing ----
    def compress_to_device(self, x):
        y = self.g_a(_nhwc(x))
        z = self.h_a(self._hyper_in(y))
        z_enc, z_hat = self.entropy_bottleneck.compress_symbols(z)
        scales_hat, means_hat = self._params(z_hat)
        y_enc, _ = self.gaussian_conditional.compress_from_scales(y, scales_hat, means_hat)
        return {"strings": [y_enc, z_enc], "shape": z.size()[-2:]}

    def decompress_from_device(self, enc, shape):
        y_enc, z_enc = enc
        z_hat = self.entropy_bottleneck.decompress(None, shape, memory_format=_CL, device_words=z_enc.device_words())
        scales_hat, means_hat = self._params(z_hat)
        y_hat = self.gaussian_conditional.decompress_from_scales(None, scales_hat, means_hat,
                                                                 device_words=y_enc.device_words())
        x_hat = self.g_s(y_hat).clamp_(0, 1)
        return {"x_hat": x_hat}


class MeanScaleHyperprior(ScaleHyperprior):
    r"""Scale Hyperprior with non zero-mean Gaussian conditionals (Minnen et al., NeurIPS 2018)."""

    def __init__(self, N, M, **kwargs):
        super().__init__(N, M, **kwargs)
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True), conv(N, N),
                                 nn.LeakyReLU(inplace=True), conv(N, N))
        self.h_s = TransformStack(deconv(N, M), nn.LeakyReLU(inplace=True), deconv(M, M * 3 // 2),
                                 nn.LeakyReLU(inplace=True), conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))

    def _hyper_in(self, y):
        return y

    def _params(self, z_hat):
        gaussian_params = self.h_s(z_hat)
        scales_hat, means_hat = gaussian_params.chunk(2, 1)
        return scales_hat, means_hat
