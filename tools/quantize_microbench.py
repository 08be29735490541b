"""Microbench for the fused quantize+index kernels (index/CDF path of SURVEY.md 8d): GB/s vs layout."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressai_environment_b200._lib import lib, ptr, check, current_stream
from compressai_environment_b200.kernels import CAI_LAYOUT_NCHW, CAI_LAYOUT_NHWC
from compressai_environment_b200.entropy_models import GaussianConditional
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
C, H, W = 192, 32, 48
dev = "cuda"; torch.manual_seed(0)
tab = torch.exp(torch.linspace(torch.log(torch.tensor(0.11)), torch.log(torch.tensor(256.0)), 64)).to(dev)
y = torch.randn(B, C, H, W, device=dev) * 5; sc = torch.exp(torch.rand(B, C, H, W, device=dev) * 8 - 3); mu = torch.randn(B, C, H, W, device=dev)
sym = torch.empty(B, C * H * W, dtype=torch.int32, device=dev); idx = torch.empty_like(sym)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = lib()
def run(layout, yy, ss, mm, name, bytes_per):
    def go(): check(L.cai_gc_quantize_index(ptr(yy), ptr(ss), ptr(mm), ptr(tab), 64, 0.11, layout, B, C, H * W, ptr(sym), ptr(idx), current_stream()), "qi")
    for _ in range(3): go()
    ts = []
    for _ in range(10):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    n = B * C * H * W
    print(f"{name}: {ms*1e3:.1f} us  {bytes_per*n/ms/1e6:.0f} GB/s ({bytes_per} B/elem, {n/1e6:.1f} M elems)")
    return sym.clone(), idx.clone()
a = run(CAI_LAYOUT_NCHW, y, sc, None, "nchw y+scales", 16)
cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
b = run(CAI_LAYOUT_NHWC, cl(y), cl(sc), None, "nhwc y+scales", 16)
assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
a = run(CAI_LAYOUT_NCHW, y, sc, mu, "nchw y+scales+means", 20)
b = run(CAI_LAYOUT_NHWC, cl(y), cl(sc), cl(mu), "nhwc y+scales+means", 20)
assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
