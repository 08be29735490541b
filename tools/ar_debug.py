"""Debug helper: plain-torch emulation of the autoregressive scan (the reference's _compress_ar arithmetic, one pixel
at a time) against the fixture strings (CPU) and against kernels.ar_encode / ar_decode (GPU, if available)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
orc.build()
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ar_jarhp.npz"))
sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
y, params = torch.from_numpy(g["y"]), torch.from_numpy(g["params"])
B, M, H, W = y.shape
table = sd["gaussian_conditional.scale_table"].numpy()
wc = (sd["context_prediction.weight"] * sd["context_prediction.mask"])
bc = sd["context_prediction.bias"]
ep = [(sd[f"entropy_parameters.{i}.weight"].flatten(1), sd[f"entropy_parameters.{i}.bias"]) for i in (0, 2, 4)]
def scan(y, params, dev="cpu"):
    y, params = y.to(dev), params.to(dev)
    yh = torch.nn.functional.pad(y.clone(), (2, 2, 2, 2))
    w2 = wc.to(dev).flatten(1); b2 = bc.to(dev); E = [(a.to(dev), b.to(dev)) for a, b in ep]
    syms = torch.zeros((B, H * W * M), dtype=torch.int32); idxs = torch.zeros_like(syms)
    tab = torch.from_numpy(table).to(dev)
    for h in range(H):
        for w in range(W):
            crop = yh[:, :, h:h + 5, w:w + 5].reshape(B, -1)
            ctx = crop @ w2.t() + b2
            v = torch.cat((params[:, :, h, w], ctx), 1)
            v = torch.nn.functional.leaky_relu(v @ E[0][0].t() + E[0][1])
            v = torch.nn.functional.leaky_relu(v @ E[1][0].t() + E[1][1])
            v = v @ E[2][0].t() + E[2][1]
            sc, mu = v[:, :M], v[:, M:]
            sc = torch.clamp(sc, min=0.11)
            idx = (len(table) - 1) - (sc[:, :, None] <= tab[None, None, :-1]).sum(-1)
            q = torch.round(y[:, :, h, w] - mu)
            yh[:, :, h + 2, w + 2] = q + mu
            o = (h * W + w) * M
            syms[:, o:o + M] = q.int().cpu(); idxs[:, o:o + M] = idx.int().cpu()
    return syms, idxs, yh
syms, idxs, yh = scan(y, params)
gt = [sd["gaussian_conditional." + k].numpy() for k in ("_quantized_cdf", "_cdf_length", "_offset")]
for b in range(B):
    ours = orc.rans_encode(syms[b].numpy(), idxs[b].numpy(), *gt)
    ref = g[f"str_0_{b}"].tobytes()
    print("image", b, "emulation bytes == reference bytes:", ours == ref, len(ours), len(ref), "rows used", len(np.unique(idxs[b].numpy())),
          "escapes", float((np.abs(syms[b].numpy()) > 20).mean()))
if torch.cuda.is_available():
    from compressai_environment_b200 import coder, kernels
    from compressai_environment_b200.models import JointAutoregressiveHierarchicalPriors
    net = JointAutoregressiveHierarchicalPriors.from_state_dict(sd).cuda().eval()
    gc = net.gaussian_conditional
    wts = net._ar_weights()
    yn, pn = net._to_nhwc(y.cuda()), net._to_nhwc(params.cuda())
    for cl, gr in ((1, 1), (8, 1), (2, 2)):
        s2, i2, yh2 = kernels.ar_encode(wts, yn, pn, gc.scale_table, gc._bound_scale(), cl, gr)
        torch.cuda.synchronize()
        ds, di = (s2.cpu() != syms), (i2.cpu() != idxs)
        yh_ref = yh.permute(0, 2, 3, 1)
        print(f"cluster {cl} group {gr}: sym diff {int(ds.sum())} idx diff {int(di.sum())} first sym diff at", ds.nonzero()[:3].tolist(),
              "first idx diff", di.nonzero()[:3].tolist(), "y_hat max err", float((yh2.cpu() - yh_ref).abs().max()))
        strings = [g[f"str_0_{b}"].tobytes() for b in range(B)]
        words, wb, keep = coder.strings_to_device(strings, torch.device("cuda"))
        yh3, st, s3 = kernels.ar_decode(wts, gc._table(), words, wb, pn, gc.scale_table, gc._bound_scale(), cl, gr, want_symbols=True)
        torch.cuda.synchronize()
        d3 = s3.cpu() != syms
        print("   decode of reference strings: status", st.tolist(), "sym diff", int(d3.sum()), "first", d3.nonzero()[:3].tolist(),
              "y_hat max err", float((yh3.cpu() - yh_ref).abs().max()))
