"""``LowerBound``: ``max(x, bound)`` whose gradient passes when ``x >= bound`` or when the gradient pushes
``x`` up (compressai/ops/bound_ops.py:36-80).  On the hot path the bound and its gate are folded into the
fused likelihood / GDN kernels; this module form serves parameters (tiny tensors) and API parity."""
import torch
import torch.nn as nn
from torch import Tensor


class LowerBoundFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        keep = (x >= bound) | (grad_output < 0)
        return keep * grad_output, None


class LowerBound(nn.Module):
    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    @torch.jit.unused
    def lower_bound(self, x):
        return LowerBoundFunction.apply(x, self.bound)

    def forward(self, x):
        if torch.jit.is_scripting():
            return torch.max(x, self.bound)
        return self.lower_bound(x)
