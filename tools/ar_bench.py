"""Autoregressive models (mbt2018 = JointAutoregressiveHierarchicalPriors, cheng2020-anchor / -attn): compress + decompress
time of a batch of 768x512 images on the GPU path, the scan kernel alone for several launch shapes, and (optionally) the
UNMODIFIED reference on the host CPU for one image.

    python tools/ar_bench.py [--model mbt2018] [--quality 3] [--batch 16] [--ref]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="mbt2018")
ap.add_argument("--quality", type=int, default=3)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--ref", action="store_true", help="also time the reference on the host CPU (one image)")
ap.add_argument("--gain", type=float, default=40.0)
a = ap.parse_args()
H, W = 512, 768


def amplify(net, g):
    with torch.no_grad():
        last = net.g_a[6] if a.model == "mbt2018" else net.g_a[len(net.g_a) - 1 if a.model == "cheng2020-anchor" else 7]
        ep = net.entropy_parameters[4]
        for m, k in ((last, g), (ep, 8.0)):
            m.weight.mul_(k)
            m.bias.mul_(k)


out = {"model": a.model, "quality": a.quality, "batch": a.batch, "image": [H, W]}
if torch.cuda.is_available():
    from compressai_environment_b200 import kernels
    from compressai_environment_b200.zoo import models
    torch.manual_seed(0)
    net = models[a.model](a.quality)
    amplify(net, a.gain)
    net = net.cuda().eval()
    net.update(force=True)
    x = torch.rand(a.batch, 3, H, W, generator=torch.Generator().manual_seed(1)).cuda()

    def timed(fn, n=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, r

    ms_c, enc = timed(lambda: net.compress(x))
    ms_d, dec = timed(lambda: net.decompress(enc["strings"], enc["shape"]))
    enc2 = net.compress(x)
    assert [bytes(s) for s in enc2["strings"][0]] == [bytes(s) for s in enc["strings"][0]]
    bits = sum(len(s) for lst in enc["strings"] for s in lst) * 8
    out.update({"compress_ms": ms_c, "decompress_ms": ms_d, "bpp": bits / (a.batch * H * W),
                "mp_per_s": a.batch * H * W / 1e6 / ((ms_c + ms_d) / 1e3)})
    # the scan alone
    with torch.no_grad():
        y = net.g_a(x.contiguous(memory_format=torch.channels_last)); z = net.h_a(y)
        z_hat = net.entropy_bottleneck.decompress(enc["strings"][1], enc["shape"]); params = net.h_s(z_hat)
    gc = net.gaussian_conditional; w = net._ar_weights()
    yn, pn = net._to_nhwc(y), net._to_nhwc(params)
    from compressai_environment_b200 import coder
    words, wb, keep = coder.strings_to_device([bytes(s) for s in enc["strings"][0]], torch.device("cuda"))
    scan = {}
    for cl, gr in ((8, 1), (8, 2), (8, 4), (8, 8), (4, 1), (4, 4), (2, 1), (1, 1)):
        if gr > a.batch:
            continue
        e_ms, _ = timed(lambda: kernels.ar_encode(w, yn, pn, gc.scale_table, gc._bound_scale(), cl, gr), 2)
        d_ms, _ = timed(lambda: kernels.ar_decode(w, gc._table(), words, wb, pn, gc.scale_table, gc._bound_scale(), cl, gr), 2)
        scan[f"cluster{cl}_group{gr}"] = {"encode_scan_ms": e_ms, "decode_scan_ms": d_ms,
                                          "us_per_pixel_decode": d_ms * 1e3 / (y.shape[2] * y.shape[3])}
    out["scan"] = scan
if a.ref:
    from oracle import oracle as orc
    orc.import_ref()
    from compressai.zoo import models as ref_models
    torch.manual_seed(0)
    rnet = ref_models[a.model](a.quality, pretrained=False)
    amplify(rnet, a.gain)
    rnet = rnet.eval()
    rnet.update(force=True)
    x1 = torch.rand(a.batch, 3, H, W, generator=torch.Generator().manual_seed(1))[:1]
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter(); renc = rnet.compress(x1); t1 = time.perf_counter()
    rdec = rnet.decompress(renc["strings"], renc["shape"]); t2 = time.perf_counter()
    out["reference_cpu"] = {"cores": os.cpu_count(), "images": 1, "compress_ms": (t1 - t0) * 1e3, "decompress_ms": (t2 - t1) * 1e3,
                            "mp_per_s": H * W / 1e6 / (t2 - t0)}
print(json.dumps(out))
