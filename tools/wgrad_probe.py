"""One weight-gradient launch at C5 shape (16 x 128 x 64 x 64 dY against 16 x 128 x 128 x 128 X, 5x5 stride 2: the second
analysis layer of a 256 x 256 training patch), timed; used under ncu for the wgrad kernel capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from compressai_environment_b200 import transforms as T

torch.backends.cudnn.allow_tf32 = False  # the comparison below is against a true fp32 cuDNN result
torch.manual_seed(0)
small = torch.randn(16, 128, 64, 64, device="cuda")
big = torch.randn(16, 128, 128, 128, device="cuda")
for _ in range(3):
    g = T.conv_wgrad(small, big, 5, 2, 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g = T.conv_wgrad(small, big, 5, 2, 2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flop = 2 * 16 * 64 * 64 * 128 * 128 * 25
ref = torch.nn.grad.conv2d_weight(big, (128, 128, 5, 5), small, stride=2, padding=2)
print(f"wgrad (split + GEMM + reduce): {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s algorithmic  max rel err {float((g - ref).abs().max() / ref.abs().max()):.2e}")
