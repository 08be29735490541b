"""torch.profiler view of the C5 training step (hyperprior q4, 16 x 3 x 256 x 256): kernels by CUDA time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from compressai_environment_b200.training import RateDistortionLoss, configure_optimizers, train_step
from compressai_environment_b200.zoo import bmshj2018_hyperprior

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = bmshj2018_hyperprior(4).to(dev).train()
opt, aux = configure_optimizers(net)
crit = RateDistortionLoss(0.018)
x = torch.rand(16, 3, 256, 256, device=dev)
for _ in range(3):
    train_step(net, crit, x, opt, aux)
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        train_step(net, crit, x, opt, aux)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages() if e.device_type.name == "CUDA" or e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"self CUDA time per step: {tot / n / 1e3:.2f} ms")
for e in rows[:28]:
    print(f"{e.self_device_time_total / n / 1e3:8.3f} ms/step  {e.count / n:7.1f} calls/step  {e.key[:90]}")
