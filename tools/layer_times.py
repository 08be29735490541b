"""Per-launch time of every transform kernel of one C2 micro-batch (hyperprior q4, 32 x 768x512), each launch alone
on one stream (no overlap with other requests), summed per layer shape over `reps` repetitions.

    python tools/layer_times.py [micro_batch] [reps]

Prints one line per distinct launch shape: launches per micro-batch, mean ms per launch, total ms per micro-batch,
algorithmic TFLOP/s (2 * MAC) -- plus the im2col / col2im / quantize helper kernels measured around the stacks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from compressai_environment_b200 import transforms as T  # noqa: E402
from compressai_environment_b200.zoo import bmshj2018_hyperprior  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = bmshj2018_hyperprior(4)
bench.amplify(net)
net = net.to(dev).eval()
net.update(force=True)
net.micro_batch = mb
x = bench.make_images(mb).to(dev)


def stack_ms(fn, n=reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


with torch.no_grad():
    for _ in range(2):
        enc = net.compress_to_device(x)
        net.decompress_from_device(enc["strings"], enc["shape"])
    torch.cuda.synchronize()
    T.TIMING, T.DETAIL = {}, []
    for _ in range(reps):
        y, y_abs = net.g_a(x, want_abs=True)
        z = net.h_a(y_abs)
        sc = net.h_s(z)
        xh = net.g_s(y, clamp=(0.0, 1.0), nchw_out=True)
    torch.cuda.synchronize()
    det, T.DETAIL, T.TIMING = T.DETAIL, None, None
    agg = {}
    for label, a, b in det:
        agg.setdefault(label, []).append(a.elapsed_time(b))
    total = 0.0
    print(f"micro-batch {mb}, {reps} repetitions, serial launches on one stream")
    for label, ts in agg.items():
        per_mb = sum(ts) / reps
        total += per_mb
        print(f"  {label:70s} n={len(ts) // reps:2d}  {sum(ts) / len(ts):7.3f} ms/launch  {per_mb:7.3f} ms/micro-batch")
    print(f"  conv_gemm total {total:.3f} ms per micro-batch (g_a + h_a + h_s + g_s once each)")
    for name, fn in (("g_a (incl. im2col)", lambda: net.g_a(x, want_abs=True)), ("h_a", lambda: net.h_a(y_abs)),
                     ("h_s", lambda: net.h_s(z)), ("g_s (incl. col2im)", lambda: net.g_s(y, clamp=(0.0, 1.0), nchw_out=True))):
        ms, _ = stack_ms(fn)
        print(f"  stack {name:22s} {ms:7.3f} ms per micro-batch (back-to-back launches)")
