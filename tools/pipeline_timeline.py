"""Diagnostic: per-chunk CUDA-event timeline of one compress+decompress step (device resident)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from compressai_environment_b200 import coder, kernels
from compressai_environment_b200.zoo import bmshj2018_hyperprior
from compressai_environment_b200.models import google as G

B, mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda")
torch.manual_seed(0)
net = bmshj2018_hyperprior(4); bench.amplify(net); net = net.to(dev).eval(); net.update(force=True); net.micro_batch = mb
x = bench.make_images(B).to(dev)
marks = []
def mark(tag):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((tag, e))
orig_gs = net.g_s.forward
def gs(*a, **k):
    r = orig_gs(*a, **k); mark('G'); return r
net.g_s.forward = gs
orig_analysis = net._analysis_chunk
def analysis(xc):
    mark("A0"); r = orig_analysis(xc); mark("A1"); return r
net._analysis_chunk = analysis
orig_enc = coder.encode
def enc(t, s, i):
    r = orig_enc(t, s, i); mark("E" if s.size(1) > 100000 else "e"); return r
G.coder.encode = enc
orig_dec = coder.decode
def dec(t, st, idx, **kw):
    r = orig_dec(t, st, idx, **kw); mark("D" if idx.size(1) > 100000 else "d"); return r
G.coder.decode = dec
import compressai_environment_b200.entropy_models.entropy_models as EM
EM.coder.decode = dec
orig_chunk = None
def dchunk(*a): pass


with torch.no_grad():
    for it in range(3):
        marks.clear()
        t0 = torch.cuda.Event(enable_timing=True); t0.record()
        e = net.compress_to_device(x); d = net.decompress_from_device(e["strings"], e["shape"])
        t1 = torch.cuda.Event(enable_timing=True); t1.record(); torch.cuda.synchronize()
print("total ms", t0.elapsed_time(t1))
print(" ".join(f"{tag}@{t0.elapsed_time(ev):.0f}" for tag, ev in marks))
