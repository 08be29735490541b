"""CPU suite: host logic of the training step (loss formula of examples/train.py:57-69, optimiser partition of
:98-129) and the DDP gradient all-reduce over gloo with world_size 2 on a stand-in module (the model's kernels need a
GPU; the collective logic does not)."""
import math
import os
import subprocess
import sys
import textwrap

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rate_distortion_loss_formula():
    from compressai_environment_b200.training import RateDistortionLoss

    x = torch.rand(2, 3, 8, 8)
    out = {"x_hat": x + 0.1, "likelihoods": {"y": torch.full((2, 4, 2, 2), 0.5), "z": torch.full((2, 2, 1, 1), 0.25)}}
    r = RateDistortionLoss(lmbda=0.01)(out, x)
    bits = 32 * 1.0 + 4 * 2.0  # -log2(0.5) per y element, -log2(0.25) per z element
    assert abs(float(r["bpp_loss"]) - bits / (2 * 8 * 8)) < 1e-6
    assert abs(float(r["mse_loss"]) - 0.01) < 1e-6
    assert abs(float(r["loss"]) - (0.01 * 255 ** 2 * 0.01 + bits / 128)) < 1e-4


def test_optimizer_partition():
    from compressai_environment_b200.models import ScaleHyperprior
    from compressai_environment_b200.training import configure_optimizers

    net = ScaleHyperprior(16, 16)
    opt, aux = configure_optimizers(net)
    n_aux = sum(p.numel() for g in aux.param_groups for p in g["params"])
    n_main = sum(p.numel() for g in opt.param_groups for p in g["params"])
    assert n_aux == net.entropy_bottleneck.quantiles.numel()
    assert n_main + n_aux == sum(p.numel() for p in net.parameters())


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist, torch.nn as nn
    from compressai_environment_b200.training import wrap_ddp
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    torch.manual_seed(0)
    net = wrap_ddp(nn.Linear(4, 3))
    x = torch.full((2, 4), float(rank + 1))
    net(x).sum().backward()
    g = net.module.weight.grad
    # DDP averages: rank0 grad rows = 2*1, rank1 = 2*2 -> mean 3
    assert torch.allclose(g, torch.full_like(g, 3.0)), g
    if rank == 0:
        print("DDP_OK")
    dist.destroy_process_group()
""")


def test_ddp_gradient_allreduce_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "DDP_OK" in out.stdout
