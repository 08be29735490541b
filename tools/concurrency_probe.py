"""Probe: do several coder launches on different streams overlap?"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressai_environment_b200 import coder
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "cdf.npz"))
dev = torch.device("cuda")
table = coder.CdfTable(*(torch.from_numpy(g[k]).to(dev) for k in ("gc_cdf", "gc_len", "gc_off")))
gen = torch.Generator(device=dev).manual_seed(1)
tab = torch.from_numpy(g["gc_scale_table"]).to(dev)
B, n, S = 32, 294912, int(sys.argv[1]) if len(sys.argv) > 1 else 4
idx = torch.randint(0, 64, (B, n), generator=gen, device=dev, dtype=torch.int32)
sym = torch.round(torch.randn((B, n), generator=gen, device=dev) * tab[idx.long()] * 4).to(torch.int32)
enc = coder.encode(table, sym, idx); dw = enc.device_words(); torch.cuda.synchronize()
streams = [torch.cuda.Stream(priority=-1) for _ in range(S)]
for what in ("encode", "decode"):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t0.record()
        evs = []
        for st in streams:
            st.wait_event(t0)
            with torch.cuda.stream(st):
                r = coder.encode(table, sym, idx) if what == "encode" else coder.decode(table, None, idx, device_words=dw)
                e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
        torch.cuda.synchronize()
    print(what, S, "streams:", [round(t0.elapsed_time(e), 1) for e in evs])
