"""Probe: last synthesis layer (128 -> 3 deconv as GEMM to 80 columns + col2im) and first layer (+GDN), one micro-batch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressai_environment_b200 import transforms as T
from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d, run_stack, to_planes
dev = "cuda"; torch.manual_seed(0); B = 32
d4 = ConvTranspose2d(128, 3, 5, 2).to(dev)
z = to_planes(torch.randn(B, 128, 256, 384, device=dev))
def timed(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    T.TIMING = {}
    ms = timed(lambda: run_stack([d4], z, clamp=(0.0, 1.0), nchw_out=True))
    ev = T.TIMING["conv_gemm_kernel"][-5:]; T.TIMING = None
    print(f"last layer (GEMM + col2im): {ms:.3f} ms; GEMM alone {sum(a.elapsed_time(b) for a, b in ev) / len(ev):.3f} ms")
