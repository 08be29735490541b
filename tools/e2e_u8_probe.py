"""Serial timing of one 256-image request through the public API with fp32 vs uint8 host buffers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from compressai_environment_b200.zoo import bmshj2018_hyperprior

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True); net.micro_batch = 32
B = 256
x = bench.make_images(B).pin_memory()
xq = (x * 255).round().to(torch.uint8).pin_memory()
of = torch.empty((B, 3, bench.H, bench.W), dtype=torch.float32).pin_memory()
oq = torch.empty((B, 3, bench.H, bench.W), dtype=torch.uint8).pin_memory()


def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r


for name, xin, out in (("fp32", x, of), ("u8", xq, oq), ("fp32", x, of), ("u8", xq, oq)):
    c_ms, enc = t(lambda: net.compress(xin))
    d_ms, _ = t(lambda: net.decompress(enc["strings"], enc["shape"], out=out))
    print(f"{name}: compress {c_ms:.1f} ms  decompress {d_ms:.1f} ms")
