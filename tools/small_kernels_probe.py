"""Probe for the HBM-bound helper kernels at bench shapes (one micro-batch of 32 768x512 images): fused
quantize+index (NHWC), tiled im2col of the first layer, tiled col2im of the last layer."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressai_environment_b200 import kernels
from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d, run_stack
dev = "cuda"; torch.manual_seed(0); B = 32
tab = torch.exp(torch.linspace(torch.log(torch.tensor(0.11)), torch.log(torch.tensor(256.0)), 64)).to(dev)
y = (torch.randn(B, 32, 48, 192, device=dev) * 2).permute(0, 3, 1, 2); sc = (torch.rand(B, 32, 48, 192, device=dev) * 4 + 0.05).permute(0, 3, 1, 2)
c1 = Conv2d(3, 128, 5, 2).to(dev); d4 = ConvTranspose2d(128, 3, 5, 2).to(dev)
x = torch.rand(B, 3, 512, 768, device=dev); z = torch.randn(B, 128, 256, 384, device=dev)
with torch.no_grad():
    for _ in range(2):
        kernels.gc_quantize_index(y, sc, None, tab, 0.11)
        run_stack([c1], x)
        run_stack([d4], z, clamp=(0.0, 1.0), nchw_out=True)
    torch.cuda.synchronize()
print("ok")
