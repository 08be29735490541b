import torch
from torch import Tensor


def ste_round(x: Tensor) -> Tensor:
    """Rounding with identity gradient (compressai/ops/ops.py:35-49). Not on the hot path; kept for API parity."""
    return torch.round(x) - x.detach() + x
