"""Generalized Divisive Normalization (compressai/layers/gdn.py:41-92):

    GDN:   y[i] = x[i] / sqrt(beta[i] + sum_j gamma[i, j] * x[j]^2)        IGDN: y[i] = x[i] * sqrt(...)

Same constructor, parameter names (``beta``, ``gamma``) and reparametrisation as the reference so that
checkpoints load unchanged.  The compute goes through :mod:`compressai_environment_b200.transforms`.
"""
import torch
import torch.nn as nn
from torch import Tensor

from .._cache import CacheOwner
from ..ops.parametrizers import NonNegativeParametrizer

__all__ = ["GDN"]


class GDN(CacheOwner, nn.Module):
    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)

        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        beta = torch.ones(in_channels)
        self.beta = nn.Parameter(self.beta_reparam.init(beta))

        self.gamma_reparam = NonNegativeParametrizer()
        gamma = gamma_init * torch.eye(in_channels)
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma))

    def effective_params(self):
        """(beta [C], gamma [C, C]) after the non-negative reparametrisation (tiny tensors)."""
        return self.beta_reparam(self.beta), self.gamma_reparam(self.gamma)

    def forward(self, x: Tensor) -> Tensor:
        from .. import transforms

        if torch.is_grad_enabled() and (x.requires_grad or self.beta.requires_grad or self.gamma.requires_grad):
            beta, gamma = self.effective_params()
            return transforms.gdn(x, beta, gamma, self.inverse)
        return transforms.run_stack([self], x)
