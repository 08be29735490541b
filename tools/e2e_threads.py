import os, sys, time, threading, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from concurrent.futures import ThreadPoolExecutor
from compressai_environment_b200.zoo import bmshj2018_hyperprior
B = 256; dev = torch.device("cuda"); torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True); net.micro_batch = 32
xh = bench.make_images(B).pin_memory(); ohs = [torch.empty_like(xh).pin_memory() for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]
T0 = time.perf_counter()
def step(slot):
    t = [time.perf_counter() - T0]
    with torch.cuda.stream(streams[slot]), torch.no_grad():
        xb = xh.to(dev, non_blocking=True)
        enc = net.compress(xb); t.append(time.perf_counter() - T0)
        dec = net.decompress(enc["strings"], enc["shape"]); t.append(time.perf_counter() - T0)
        ohs[slot].copy_(dec["x_hat"], non_blocking=True); torch.cuda.current_stream().synchronize(); t.append(time.perf_counter() - T0)
    return slot, [round(v * 1e3) for v in t]
for nw in (1, 2, 3):
    with ThreadPoolExecutor(nw) as ex:
        list(ex.map(step, [i % nw for i in range(nw)]))
        T0 = time.perf_counter()
        res = list(ex.map(step, [i % nw for i in range(12)]))
        tot = time.perf_counter() - T0
    print(nw, "workers: %.0f ms per request" % (tot * 1e3 / 12), res[-2:])
