"""``NonNegativeParametrizer`` -- the square reparametrisation GDN uses for ``beta`` and ``gamma``.

Behaviour contract (compressai/ops/parametrizers.py:38-64 of the reference): with ``pedestal = offset^2`` and
``floor = sqrt(minimum + pedestal)``, a stored raw tensor ``r`` represents ``max(r, floor)^2 - pedestal`` (never
below ``minimum``), and ``init(v) = sqrt(max(v + pedestal, pedestal))`` maps a desired value to its raw form.
Buffers are named ``pedestal`` and ``lower_bound.bound`` as in reference checkpoints.
"""
import math

import torch
import torch.nn as nn
from torch import Tensor

from .bound_ops import LowerBound


class NonNegativeParametrizer(nn.Module):
    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2**-18):
        super().__init__()
        self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
        ped = self.reparam_offset * self.reparam_offset
        self.register_buffer("pedestal", torch.tensor([ped], dtype=torch.float32))
        self.lower_bound = LowerBound(math.sqrt(self.minimum + ped))

    def init(self, value: Tensor) -> Tensor:
        """Raw representation of ``value`` (used once, when the owning layer creates its parameter)."""
        return (value + self.pedestal).clamp_min(self.pedestal).sqrt()

    def forward(self, raw: Tensor) -> Tensor:
        return self.lower_bound(raw).square() - self.pedestal
