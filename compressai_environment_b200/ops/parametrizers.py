"""``NonNegativeParametrizer`` (compressai/ops/parametrizers.py:38-64): p = max(raw, b)^2 - pedestal with
b = sqrt(minimum + pedestal), pedestal = reparam_offset^2.  Applied to the C and C x C GDN parameters."""
import torch
import torch.nn as nn
from torch import Tensor

from .bound_ops import LowerBound


class NonNegativeParametrizer(nn.Module):
    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2**-18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset**2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset**2) ** 0.5)

    def init(self, x: Tensor) -> Tensor:
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x: Tensor) -> Tensor:
        out = self.lower_bound(x)
        return out**2 - self.pedestal
