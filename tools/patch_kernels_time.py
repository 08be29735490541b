"""Device time of the tiled im2col (first layer) and col2im (last layer) kernels at bench shapes (one micro-batch of
32 768x512 images), from torch.profiler (kernel durations, L2 not flushed: inputs are far larger than L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d, run_stack

dev = "cuda"
torch.manual_seed(0)
B = 32
c1 = Conv2d(3, 128, 5, 2).to(dev)
d4 = ConvTranspose2d(128, 3, 5, 2).to(dev)
x = torch.rand(B, 3, 512, 768, device=dev)
z = torch.randn(B, 128, 256, 384, device=dev)
with torch.no_grad():
    for _ in range(3):
        run_stack([c1], x)
        run_stack([d4], z, clamp=(0.0, 1.0), nchw_out=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            run_stack([c1], x)
            run_stack([d4], z, clamp=(0.0, 1.0), nchw_out=True)
        torch.cuda.synchronize()
for e in prof.key_averages():
    if "im2col" in e.key or "col2im" in e.key:
        print(f"{e.key[:60]:60s} {e.self_device_time_total / e.count / 1e3:.3f} ms x {e.count}")
