import torch, sys, os
sys.path.insert(0, os.getcwd())
from compressai_environment_b200 import transforms as T
torch.manual_seed(0)
small = torch.randn(2, 128, 16, 64, device="cuda"); big = torch.randn(2, 128, 32, 128, device="cuda")
g = T.conv_wgrad(small, big, 5, 2, 2)
torch.cuda.synchronize()
print("ok", g.abs().max().item())
