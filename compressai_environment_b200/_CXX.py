"""Drop-in for the reference's ``compressai._CXX`` pybind11 module (compressai/cpp_exts/ops/ops.cpp:111-118).

``pmf_to_quantized_cdf(pmf: list[float], precision: int) -> list[int]`` with the reference's error
behaviour (``ValueError`` for negative / non-finite / all-zero pmfs, ops.cpp:46-64 via pybind's
std::domain_error translation).  The work is done by the ``cai_pmf_to_quantized_cdf`` kernel.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from . import _lib
from ._lib import CaiError, check, current_stream, lib, ptr

__name__ = "compressai._CXX"  # noqa: A001  (ops.cpp:112)


def pmf_rows_to_quantized_cdf(pmf: torch.Tensor, pmf_length: torch.Tensor, tail_mass, precision: int = 16):
    """Batched device entry: pmf float32 [K, Lp] (+ optional tail [K]) -> int32 cdf [K, Lp + 2 or Lp + 1].
    Row convention of EntropyModel._pmf_to_cdf (entropy_models.py:204-212).  Raises ValueError like the
    reference when a row is invalid."""
    _lib.require_cuda(pmf, "pmf")
    pmf = pmf.detach().to(torch.float32).contiguous()
    K, Lp = int(pmf.size(0)), int(pmf.size(1))
    ln = pmf_length.detach().reshape(-1).to(device=pmf.device, dtype=torch.int32).contiguous()
    tail = None
    if tail_mass is not None:
        tail = tail_mass.detach().reshape(-1).to(device=pmf.device, dtype=torch.float32).contiguous()
    W = Lp + (2 if tail is not None else 1)
    cdf = torch.empty((K, W), dtype=torch.int32, device=pmf.device)
    status = torch.empty(K, dtype=torch.int32, device=pmf.device)
    with torch.cuda.device(pmf.device):
        check(lib().cai_pmf_to_quantized_cdf(ptr(pmf), ptr(ln), ptr(tail), K, Lp, int(precision), ptr(cdf),
                                             ptr(status), current_stream()), "cai_pmf_to_quantized_cdf")
    st = status.cpu()
    if bool((st != 0).any()):
        k = int(torch.nonzero(st != 0)[0])
        raise ValueError(_lib.STATUS_TEXT.get(int(st[k]), f"pmf_to_quantized_cdf failed ({int(st[k])})"))
    return cdf


def pmf_to_quantized_cdf(pmf: List[float], precision: int) -> List[int]:
    if not torch.cuda.is_available():
        raise CaiError("compressai._CXX needs a CUDA device: there is no CPU fallback")
    p = np.asarray(pmf, dtype=np.float32).reshape(1, -1)
    if p.shape[1] == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(p).to(dev)
    ln = torch.tensor([p.shape[1]], dtype=torch.int32, device=dev)
    cdf = pmf_rows_to_quantized_cdf(t, ln, None, precision)
    return [int(v) & 0xFFFFFFFF for v in cdf.reshape(-1).cpu().tolist()]
