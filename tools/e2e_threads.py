"""e2e serving loop timeline: N host threads, one request each in flight; prints per-request phase times (ms since
start: begin, compress done, decompress done, result on host) and allocator activity inside the timed region."""
import os, sys, time, queue, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from concurrent.futures import ThreadPoolExecutor
from compressai_environment_b200.zoo import bmshj2018_hyperprior
B = 256; dev = torch.device("cuda"); torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True); net.micro_batch = 32
xh = bench.make_images(B).pin_memory(); ohs = [torch.empty_like(xh).pin_memory() for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]
T0 = time.perf_counter()
free = queue.SimpleQueue()
def step(_):
    slot = free.get()
    t = [time.perf_counter() - T0]
    with torch.cuda.stream(streams[slot]), torch.no_grad():
        xb = xh.to(dev, non_blocking=True)
        enc = net.compress(xb); t.append(time.perf_counter() - T0)
        dec = net.decompress(enc["strings"], enc["shape"]); t.append(time.perf_counter() - T0)
        ohs[slot].copy_(dec["x_hat"], non_blocking=True); torch.cuda.current_stream().synchronize(); t.append(time.perf_counter() - T0)
    free.put(slot)
    return slot, [round(v * 1e3) for v in t]
for nw in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    while not free.empty(): free.get()
    for i in range(nw): free.put(i)
    with ThreadPoolExecutor(nw) as ex:
        list(ex.map(step, range(nw * 2)))
        torch.cuda.synchronize()
        m0 = torch.cuda.memory_stats()
        T0 = time.perf_counter()
        res = list(ex.map(step, range(12)))
        tot = time.perf_counter() - T0
        m1 = torch.cuda.memory_stats()
    print(nw, "workers: %.0f ms per request;" % (tot * 1e3 / 12), "cudaMalloc calls in region:", m1["num_device_alloc"] - m0["num_device_alloc"],
          "frees:", m1["num_device_free"] - m0["num_device_free"], "retries:", m1["num_alloc_retries"] - m0["num_alloc_retries"],
          "reserved GB: %.1f" % (m1["reserved_bytes.all.current"] / 2**30))
    for r in res: print("   ", r)
