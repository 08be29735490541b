"""``LowerBound`` -- a clamp-from-below whose gradient is gated instead of zeroed.

Behaviour contract (what compressai/ops/bound_ops.py:36-80 of the reference defines and its checkpoints rely on):
``y = max(x, bound)``; the incoming gradient reaches ``x`` wherever ``x`` already satisfies the bound OR the
gradient would move ``x`` upwards (negative gradient); elsewhere it is dropped.  ``bound`` is a one-element buffer
named ``bound`` so that reference ``state_dict`` keys (``...lower_bound.bound``) line up.

On the hot path this gate is folded into the fused likelihood / GDN kernels (csrc/likelihood.cu, csrc/gdn.cu);
the module form below only ever sees parameter-sized tensors (C or C x C) and scalars.
"""
import torch
import torch.nn as nn
from torch import Tensor


class LowerBoundFunction(torch.autograd.Function):
    """Autograd node of :class:`LowerBound`: forward ``maximum``, backward the gated pass-through."""

    @staticmethod
    def forward(ctx, x: Tensor, bound: Tensor) -> Tensor:
        ctx.save_for_backward(x, bound)
        return torch.maximum(x, bound)

    @staticmethod
    def backward(ctx, g: Tensor):
        x, bound = ctx.saved_tensors
        inside = x.ge(bound)          # the clamp is inactive
        pushes_up = g.lt(0)           # a descent step would raise x back towards the feasible side
        return torch.where(inside.logical_or(pushes_up), g, torch.zeros_like(g)), None


class LowerBound(nn.Module):
    """``max(x, bound)`` with the gated gradient above.  ``bound`` is fixed at construction."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.tensor([float(bound)], dtype=torch.float32))

    def forward(self, x: Tensor) -> Tensor:
        return LowerBoundFunction.apply(x, self.bound)

    def extra_repr(self) -> str:
        return f"bound={float(self.bound):g}" if self.bound.numel() == 1 and not self.bound.is_meta else ""
