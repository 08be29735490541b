import os, sys, time, cProfile, pstats, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from compressai_environment_b200.zoo import bmshj2018_hyperprior
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda"); torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True); net.micro_batch = 32
xh = bench.make_images(B).pin_memory(); oh = torch.empty_like(xh).pin_memory()
def step():
    t = [time.perf_counter()]
    xb = xh.to(dev, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    enc = net.compress(xb); t.append(time.perf_counter())
    dec = net.decompress(enc["strings"], enc["shape"]); torch.cuda.synchronize(); t.append(time.perf_counter())
    oh.copy_(dec["x_hat"], non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    return [round((b - a) * 1e3, 1) for a, b in zip(t, t[1:])]
with torch.no_grad():
    step(); print("h2d, compress, decompress, d2h (ms):", step())
    pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(32)
