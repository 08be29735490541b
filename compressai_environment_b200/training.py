"""Training-step pieces of the hot path (config C5), mirroring examples/train.py of the reference:
``RateDistortionLoss`` (:49-69), ``configure_optimizers`` (:98-129: main Adam on everything but ``*.quantiles``,
aux Adam on the quantiles) and one optimisation step (:132-165).  The reference's single-process
``nn.DataParallel`` (:88-95, :323-324) is replaced by one process per GPU with ``torch.distributed`` DDP: the only
collective is the bucketed NCCL all-reduce of gradients (20.3 MB fp32 for N=128, M=192); the codec path itself has no
collective.  Forward likelihoods run on the fused kernels (cai_gc_forward/backward, cai_eb_forward/backward).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.optim as optim


class RateDistortionLoss(nn.Module):
    """loss = lmbda * 255^2 * MSE(x_hat, x) + sum_k sum(log(lik_k)) / (-ln 2 * N * H * W)."""

    def __init__(self, lmbda=1e-2):
        super().__init__()
        self.mse = nn.MSELoss()
        self.lmbda = lmbda

    def forward(self, output, target):
        N, _, H, W = target.size()
        num_pixels = N * H * W
        out = {}
        out["bpp_loss"] = sum(torch.log(lik).sum() / (-math.log(2) * num_pixels) for lik in output["likelihoods"].values())
        out["mse_loss"] = self.mse(output["x_hat"], target)
        out["loss"] = self.lmbda * 255**2 * out["mse_loss"] + out["bpp_loss"]
        return out


def configure_optimizers(net, learning_rate=1e-4, aux_learning_rate=1e-3):
    """Separate the entropy-bottleneck quantiles (aux optimiser) from every other parameter (main optimiser)."""
    named = dict(net.named_parameters())
    parameters = {n for n, p in named.items() if not n.endswith(".quantiles") and p.requires_grad}
    aux_parameters = {n for n, p in named.items() if n.endswith(".quantiles") and p.requires_grad}
    assert len(parameters & aux_parameters) == 0
    assert len(parameters | aux_parameters) == len([p for p in named.values() if p.requires_grad])
    optimizer = optim.Adam((named[n] for n in sorted(parameters)), lr=learning_rate)
    aux_optimizer = optim.Adam((named[n] for n in sorted(aux_parameters)), lr=aux_learning_rate)
    return optimizer, aux_optimizer


def wrap_ddp(net, device=None):
    """DistributedDataParallel over NCCL (GPU) / gloo (CPU tensors); no-op when torch.distributed is not initialised."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return net
    ids = [device.index] if device is not None and device.type == "cuda" else None
    return nn.parallel.DistributedDataParallel(net, device_ids=ids, broadcast_buffers=False)


def train_step(model, criterion, batch, optimizer, aux_optimizer, clip_max_norm=1.0):
    """One step of examples/train.py:132-165.  ``model`` may be DDP-wrapped."""
    optimizer.zero_grad()
    aux_optimizer.zero_grad()
    out_net = model(batch)
    out = criterion(out_net, batch)
    out["loss"].backward()
    if clip_max_norm > 0:
        torch.nn.utils.clip_grad_norm_(model.parameters(), clip_max_norm)
    optimizer.step()
    core = model.module if hasattr(model, "module") else model
    aux_loss = core.aux_loss()
    aux_loss.backward()
    aux_optimizer.step()
    out["aux_loss"] = aux_loss
    return out
