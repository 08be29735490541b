"""Probe: one big conv (+fused GDN) launch, timed; used under ncu for the conv kernel capture."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d, run_stack, to_planes
from compressai_environment_b200.layers import GDN
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
kind = sys.argv[2] if len(sys.argv) > 2 else "conv"
dev = "cuda"; torch.manual_seed(0)
if kind == "conv1":
    mods = [Conv2d(3, 128).to(dev), GDN(128).to(dev)]; x = torch.rand(B, 3, 512, 768, device=dev); macs = B * 256 * 384 * (128 * 75 + 128 * 128)
elif kind == "conv":
    mods = [Conv2d(128, 128).to(dev), GDN(128).to(dev)]; x = torch.randn(B, 128, 256, 384, device=dev); macs = B * 128 * 192 * (128 * 128 * 25 + 128 * 128)
else:
    mods = [ConvTranspose2d(128, 128).to(dev), GDN(128, inverse=True).to(dev)]; x = torch.randn(B, 128, 128, 192, device=dev); macs = B * 128 * 192 * 128 * 128 * 25 + B * 256 * 384 * 128 * 128
with torch.no_grad():
    xp = to_planes(x) if kind != 'conv1' else x
    for _ in range(3): y = run_stack(mods + [Conv2d(128, 128, 3, 1).to(dev)][:0], xp) if False else None
    from compressai_environment_b200 import transforms as T
    def go():
        if kind in ("conv", "conv1"): return T._run_conv(mods[0], xp, None, ("planes",), None, gdn=(T._prep_gdn(mods[1]).packed, T._prep_gdn(mods[1]).bias, 1))
        return T._run_deconv(mods[0], xp, None, ("planes",), None, gdn=(T._prep_gdn(mods[1]).packed, T._prep_gdn(mods[1]).bias, 2))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
print(f"{kind} B={B}: {ms:.3f} ms  {2*macs/ms/1e9:.1f} TFLOP/s (algorithmic fp32-equivalent)")

if os.environ.get("CAI_CONV_DEBUG") and int(os.environ["CAI_CONV_DEBUG"]) & 32:
    import ctypes, numpy as np
    from compressai_environment_b200._lib import lib
    L = lib(); buf = (ctypes.c_longlong * 512)()
    L._handle  # noqa
    f = ctypes.CDLL(L._name).cai_debug_conv_trace; f.argtypes = [ctypes.c_void_p]; f(buf)
    a = np.array(buf[:]).reshape(64, 8)
    d = np.diff(a, axis=1)
    print("phase cycles (median over 64 CTAs): prologue, producer-loop, wait-acc, gdnA, wait-acc2, epiB-phase1(TMEM+math+STS), epiB-phase2(copy-out)")
    print(np.median(d, axis=0).astype(int).tolist(), "total", int(np.median(a[:, 7] - a[:, 0])))
