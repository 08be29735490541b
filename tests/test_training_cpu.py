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


def test_tap_groups_and_weight_schedule_cpu():
    """Host-side k-step schedule of the transform kernel (pure torch, no GPU): taps are grouped by (dy, dx mod
    stride) and the packed weight blob follows  group -> channel chunk -> tap in group  (include/cai_b200.h)."""
    import torch
    from compressai_environment_b200.transforms import _grouped, pack_weights, _BK

    k, s, p = 5, 2, 2
    taps = [(ky - p, kx - p) for ky in range(k) for kx in range(k)]
    rows = list(range(len(taps)))
    g, order = _grouped(taps, rows, s)
    assert sorted(order) == rows and sum(g.glen) == 25 and g.glen == [3, 2] * 5
    t0 = 0
    for n in g.glen:
        grp = g[t0:t0 + n]
        assert len({dy for dy, _ in grp}) == 1 and len({dx % s for _, dx in grp}) == 1
        assert [dx for _, dx in grp] == sorted(dx for _, dx in grp)
        t0 += n
    g1, _ = _grouped([(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], list(range(9)), 1)
    assert g1.glen == [3, 3, 3]

    # blob order: tile (g, kc, t) must hold tap t's channels [kc*32, kc*32+32)
    T, cout, cin, bn = 5, 16, 64, 16
    w = torch.arange(T * cout * cin, dtype=torch.float32).reshape(T, cout, cin) / 7.0
    glen = [3, 2]
    blob = pack_weights(w, bn, glen).view(torch.bfloat16).reshape(-1, 2, _BK // 8, bn // 8, 8, 8)  # kstep, hi/lo, k8, r8, r, k
    sched = [(t, c) for t0, n in ((0, 3), (3, 2)) for c in range(cin // _BK) for t in range(t0, t0 + n)]
    assert blob.shape[0] == len(sched)
    for ks, (t, c) in enumerate(sched):
        tile = blob[ks, 0].permute(1, 2, 0, 3).reshape(bn, _BK).float()        # [row, k]
        ref = w[t, :, c * _BK:(c + 1) * _BK].to(torch.bfloat16).float()
        assert torch.equal(tile, ref), (ks, t, c)
    # default (no groups) keeps tap-major order
    blob0 = pack_weights(w, bn).view(torch.bfloat16).reshape(-1, 2, _BK // 8, bn // 8, 8, 8)
    assert torch.equal(blob0[1, 0].permute(1, 2, 0, 3).reshape(bn, _BK).float(), w[0, :, _BK:2 * _BK].to(torch.bfloat16).float())
