"""e2e serving loop, GPU-side view: with N requests in flight, what fraction of wall time has a conv_gemm kernel
running, a coder kernel running, or nothing of ours running?  (CUDA-event intervals from the TIMING hooks.)"""
import os, sys, time, queue, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from concurrent.futures import ThreadPoolExecutor
from compressai_environment_b200 import coder, transforms
from compressai_environment_b200.zoo import bmshj2018_hyperprior
NW = int(sys.argv[1]) if len(sys.argv) > 1 else 3
NREQ = int(sys.argv[2]) if len(sys.argv) > 2 else 9
B = 256; dev = torch.device("cuda"); torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True); net.micro_batch = 32
xh = bench.make_images(B).pin_memory(); ohs = [torch.empty_like(xh).pin_memory() for _ in range(NW)]
streams = [torch.cuda.Stream() for _ in range(NW)]
free = queue.SimpleQueue()
for i in range(NW): free.put(i)
copies = []
import threading
REQ = threading.local()
class TaggedList(list):
    def append(self, pair):
        super().append((pair[0], pair[1], getattr(REQ, "rid", -1), getattr(REQ, "phase", "?")))
class TaggedDict(dict):
    def setdefault(self, k, d=None):
        if k not in self: self[k] = TaggedList()
        return self[k]
def step(rid):
    slot = free.get()
    REQ.rid = rid
    with torch.cuda.stream(streams[slot]), torch.no_grad():
        a = torch.cuda.Event(enable_timing=True); a.record()
        xb = xh.to(dev, non_blocking=True)
        b = torch.cuda.Event(enable_timing=True); b.record(); copies.append(("h2d", a, b, rid))
        REQ.phase = "c"
        enc = net.compress(xb)
        REQ.phase = "d"
        dec = net.decompress(enc["strings"], enc["shape"])
        a = torch.cuda.Event(enable_timing=True); a.record()
        ohs[slot].copy_(dec["x_hat"], non_blocking=True)
        b = torch.cuda.Event(enable_timing=True); b.record(); copies.append(("d2h", a, b, rid))
        torch.cuda.current_stream().synchronize()
    free.put(slot)
with ThreadPoolExecutor(NW) as ex:
    list(ex.map(step, range(NW * 2)))
    torch.cuda.synchronize(); copies.clear()
    coder.TIMING = TaggedDict(); transforms.TIMING = TaggedDict()
    base = torch.cuda.Event(enable_timing=True); base.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    list(ex.map(step, range(NREQ)))
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
def ivals(pairs): return sorted((base.elapsed_time(p[0]), base.elapsed_time(p[1])) for p in pairs)
def union(iv):
    tot, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None: tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else: cur_b = max(cur_b, b)
    if cur_b is not None: tot += cur_b - cur_a
    return tot
conv = ivals(transforms.TIMING.get("conv_gemm_kernel", []))
enc = ivals(coder.TIMING.get("rans_encode_kernel", [])); dec = ivals(coder.TIMING.get("rans_decode_kernel", []))
cp = ivals([(a, b) for _, a, b, _r in copies])
print(f"{NW} in flight, {NREQ} requests: wall {wall:.0f} ms = {wall/NREQ:.0f} ms/request")
print(f"  conv kernels: {len(conv)} launches, sum {sum(b-a for a,b in conv):.0f} ms, union {union(conv):.0f} ms ({100*union(conv)/wall:.0f}% of wall)")
print(f"  coder kernels: enc sum {sum(b-a for a,b in enc):.0f} union {union(enc):.0f}; dec sum {sum(b-a for a,b in dec):.0f} union {union(dec):.0f}; any-coder union {union(sorted(enc+dec)):.0f} ms")
print(f"  copies: sum {sum(b-a for a,b in cp):.0f} union {union(cp):.0f} ms")
allk = sorted(conv + enc + dec)
print(f"  any of ours running: {union(allk):.0f} ms ({100*union(allk)/wall:.0f}%); conv-or-copy union {union(sorted(conv+cp)):.0f}")
# gaps in conv activity > 3 ms: what was running?
gaps, last = [], 0.0
for a, b in conv:
    if a - last > 3.0: gaps.append((last, a))
    last = max(last, b)
print("  conv gaps > 3 ms:", [(round(a), round(b - a)) for a, b in gaps][:40])

# per-request GPU phase spans (ms since base): h2d | analysis convs | encodes | decodes | synthesis convs | d2h
def span(items):
    if not items: return "      -      "
    a = min(base.elapsed_time(p[0]) for p in items); b = max(base.elapsed_time(p[1]) for p in items)
    return f"{a:6.0f}-{b:<6.0f}"
convs = transforms.TIMING.get("conv_gemm_kernel", []); encs = coder.TIMING.get("rans_encode_kernel", []); decs = coder.TIMING.get("rans_decode_kernel", [])
print("  req   h2d           ana-conv      encode        decode        syn-conv(+h_s) d2h")
for rid in range(NREQ):
    h = [(a, b) for k, a, b, r in copies if r == rid and k == "h2d"]; d = [(a, b) for k, a, b, r in copies if r == rid and k == "d2h"]
    print(f"  {rid:3d}  ", span(h), span([p for p in convs if p[2] == rid and p[3] == "c"]), span([p for p in encs if p[2] == rid]),
          span([p for p in decs if p[2] == rid]), span([p for p in convs if p[2] == rid and p[3] == "d"]), span(d))
