"""Host-side launch cost of one request (no synchronisation inside the timed calls): how long does the Python thread
need to enqueue compress / decompress for 256 images?  If this approaches the GPU time per step, the serving loop is
host (GIL) bound, not GPU bound."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from compressai_environment_b200 import _lib
from compressai_environment_b200.zoo import bmshj2018_hyperprior
B = 256; dev = torch.device("cuda"); torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True)
net.micro_batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = bench.make_images(B).to(dev)
with torch.no_grad():
    for _ in range(3):
        enc = net.compress_to_device(x); dec = net.decompress_from_device(enc["strings"], enc["shape"])
    torch.cuda.synchronize()
    for _ in range(3):
        l0 = _lib.LAUNCHES
        t0 = time.perf_counter(); enc = net.compress_to_device(x); t1 = time.perf_counter()
        dec = net.decompress_from_device(enc["strings"], enc["shape"]); t2 = time.perf_counter()
        torch.cuda.synchronize(); t3 = time.perf_counter()
        print(f"enqueue compress {1e3*(t1-t0):.1f} ms, enqueue decompress {1e3*(t2-t1):.1f} ms, drain {1e3*(t3-t2):.1f} ms, launches {_lib.LAUNCHES-l0}")
    enc = net.compress(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); enc = net.compress(x); t1 = time.perf_counter()
    dec = net.decompress(enc["strings"], enc["shape"]); t2 = time.perf_counter()
    print(f"public API: compress {1e3*(t1-t0):.1f} ms, decompress {1e3*(t2-t1):.1f} ms")
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); enc = net.compress_to_device(x); dec = net.decompress_from_device(enc["strings"], enc["shape"]); pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
