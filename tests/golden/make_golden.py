"""Generate the committed golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    ./oracle/build_ref.sh && python tests/golden/make_golden.py

The reference (CompressAI 1.2.0.dev0) is imported from oracle/_ref; nothing from this repo's product
code is used.  Outputs: coder.npz (rANS known-answer streams), cdf.npz (pmf_to_quantized_cdf KATs and
the default GaussianConditional / a seeded EntropyBottleneck table), fp.npz (likelihood / GDN /
quantize / build_indexes fixtures), model_*.npz (tiny seeded models: state_dict, input, reference
strings and reconstructions).
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402

compressai = orc.import_ref()
from compressai import _CXX, ans  # noqa: E402
from compressai.entropy_models import EntropyBottleneck, GaussianConditional  # noqa: E402
from compressai.layers import GDN  # noqa: E402
from compressai.models import FactorizedPrior, MeanScaleHyperprior, ScaleHyperprior  # noqa: E402
from compressai.models.google import get_scale_table  # noqa: E402


def lcg(seed, n):
    out, x = [], seed
    for _ in range(n):
        x = (1103515245 * x + 12345) & 0x7FFFFFFF
        out.append(x)
    return np.array(out, dtype=np.int64)


def ref_encode(sym, idx, cdf, ln, off):
    return ans.RansEncoder().encode_with_indexes(
        [int(v) for v in sym], [int(v) for v in idx], np.asarray(cdf).tolist(),
        [int(v) for v in ln], [int(v) for v in off])


def ref_decode(data, idx, cdf, ln, off):
    return ans.RansDecoder().decode_with_indexes(
        data, [int(v) for v in idx], np.asarray(cdf).tolist(), [int(v) for v in ln], [int(v) for v in off])


def coder_fixtures():
    out = {}
    cdf = np.array([[0, 21845, 65534, 65535, 65536], [0, 32768, 65536, 0, 0]], np.int32)
    ln = np.array([5, 3], np.int32)
    off = np.array([-1, 0], np.int32)
    out["small_cdf"], out["small_len"], out["small_off"] = cdf, ln, off
    # KAT-A
    sym = np.array([0, -1, 1, 0, 5, -3, 0, 0, 1, -1, 0, 2, -2, 0, 40, -17], np.int32)
    idx = np.array([0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1], np.int32)
    b = ref_encode(sym, idx, cdf, ln, off)
    assert ref_decode(b, idx, cdf, ln, off) == sym.tolist()
    out["katA_sym"], out["katA_idx"], out["katA_bytes"] = sym, idx, np.frombuffer(b, np.uint8)
    # KAT-B (LCG stream)
    r = lcg(12345, 8192)
    idx = (r[:4096] & 1).astype(np.int32)
    sym = (((r[4096:] >> 8) % 7) - 3).astype(np.int32)
    b = ref_encode(sym, idx, cdf, ln, off)
    assert ref_decode(b, idx, cdf, ln, off) == sym.tolist()
    out["katB_sym"], out["katB_idx"], out["katB_bytes"] = sym, idx, np.frombuffer(b, np.uint8)
    out["katB_sha256"] = np.frombuffer(hashlib.sha256(b).digest(), np.uint8)

    # default Gaussian-conditional table
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    gcdf, glen, goff = gc._quantized_cdf.numpy(), gc._cdf_length.numpy(), gc._offset.numpy()
    for name, s in (("katF0", [0, 0, 0, 0]), ("katF1", [100000, -100000, 5, -7])):
        b = ref_encode(s, [0] * 4, gcdf, glen, goff)
        assert ref_decode(b, [0] * 4, gcdf, glen, goff) == s
        out[name + "_sym"] = np.array(s, np.int32)
        out[name + "_bytes"] = np.frombuffer(b, np.uint8)
    # Gaussian streams on the default table at several inflation factors t (escape rates 0 / ~9 / ~50 %)
    g = torch.Generator().manual_seed(1234)
    table = gc.scale_table.numpy()
    for t, n in ((1.0, 4096), (4.0, 4096), (16.0, 2048), (1.0, 2), (1.0, 3), (1.0, 31), (1.0, 32), (1.0, 33)):
        idx = torch.randint(0, 64, (n,), generator=g).numpy().astype(np.int32)
        z = torch.randn(n, generator=g).numpy()
        sym = np.rint(z * table[idx] * t).astype(np.int32)
        b = ref_encode(sym, idx, gcdf, glen, goff)
        assert ref_decode(b, idx, gcdf, glen, goff) == sym.tolist()
        key = f"gauss_t{int(t)}_n{n}"
        out[key + "_sym"], out[key + "_idx"], out[key + "_bytes"] = sym, idx, np.frombuffer(b, np.uint8)
    # extreme escapes (8 payload nibbles, both signs) and symbols sitting exactly on the sentinel
    sym = np.array([2**27 - 1, -(2**27), 1565, -1566, 1566, 0, 3, -3, 70000, -70000], np.int32)
    idx = np.array([63, 63, 63, 63, 63, 0, 1, 1, 5, 5], np.int32)
    b = ref_encode(sym, idx, gcdf, glen, goff)
    assert ref_decode(b, idx, gcdf, glen, goff) == sym.tolist()
    out["edge_sym"], out["edge_idx"], out["edge_bytes"] = sym, idx, np.frombuffer(b, np.uint8)
    # buffered encoder fed in two calls == one call; set_stream + 2x decode_stream == one decode
    sym, idx = out["gauss_t4_n4096_sym"], out["gauss_t4_n4096_idx"]
    be = ans.BufferedRansEncoder()
    be.encode_with_indexes(sym[:1000].tolist(), idx[:1000].tolist(), gcdf.tolist(), glen.tolist(), goff.tolist())
    be.encode_with_indexes(sym[1000:].tolist(), idx[1000:].tolist(), gcdf.tolist(), glen.tolist(), goff.tolist())
    assert be.flush() == out["gauss_t4_n4096_bytes"].tobytes()
    d = ans.RansDecoder()
    d.set_stream(out["gauss_t4_n4096_bytes"].tobytes())
    a = d.decode_stream(idx[:1000].tolist(), gcdf.tolist(), glen.tolist(), goff.tolist())
    a += d.decode_stream(idx[1000:].tolist(), gcdf.tolist(), glen.tolist(), goff.tolist())
    assert a == sym.tolist()
    np.savez_compressed(os.path.join(HERE, "coder.npz"), **out)
    return gc


def cdf_fixtures(gc):
    out = {}
    kats = {
        "ref": ([0.1, 0.2, 0.0, 0.0], 16),  # tests/test_ops.py:104-106
        "C": ([0.5, 0.25, 0.125, 0.125, 1e-9], 16),
        "D": ([1e-7] * 6 + [0.9], 16),
        "E": ([0.3, 0, 0, 0.3, 0, 0.4], 12),
    }
    for k, (pmf, prec) in kats.items():
        out[f"kat{k}_pmf"] = np.array(pmf, np.float32)
        out[f"kat{k}_prec"] = np.array(prec)
        out[f"kat{k}_cdf"] = np.array(_CXX.pmf_to_quantized_cdf(pmf, prec), np.int64)
    g = torch.Generator().manual_seed(7)
    for i, m in enumerate((3, 17, 64, 257, 1000)):
        p = torch.rand(m, generator=g) ** 8  # many tiny entries -> steals
        p[torch.rand(m, generator=g) < 0.3] = 0.0
        p[0] = 0.5
        p = (p / p.sum()).float().numpy()
        out[f"rand{i}_pmf"] = p
        out[f"rand{i}_cdf"] = np.array(_CXX.pmf_to_quantized_cdf(p.tolist(), 16), np.int64)
    out["gc_scale_table"] = gc.scale_table.numpy()
    out["gc_cdf"] = gc._quantized_cdf.numpy()
    out["gc_len"] = gc._cdf_length.numpy()
    out["gc_off"] = gc._offset.numpy()
    # the float pmf rows the reference fed to pmf_to_quantized_cdf (entropy_models.py:625-648)
    multiplier = -gc._standardized_quantile(gc.tail_mass / 2)
    pmf_center = torch.ceil(gc.scale_table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = torch.max(pmf_length).item()
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    s = gc.scale_table.unsqueeze(1).float()
    upper = gc._standardized_cumulative((0.5 - samples) / s)
    lower = gc._standardized_cumulative((-0.5 - samples) / s)
    out["gc_pmf"] = (upper - lower).numpy()
    out["gc_pmf_len"] = pmf_length.numpy()
    out["gc_tail"] = (2 * lower[:, :1]).numpy().ravel()
    # a seeded, perturbed entropy bottleneck (so rows differ and factors are non-zero)
    torch.manual_seed(11)
    eb = EntropyBottleneck(6)
    with torch.no_grad():
        for n, p in eb.named_parameters():
            if n == "quantiles":
                p.copy_(torch.tensor([[-7.3, 0.4, 9.1], [-3.0, -0.2, 2.5], [-12.5, 1.1, 11.0],
                                      [-1.2, 0.0, 1.4], [-20.0, 3.0, 25.0], [-5.0, -1.0, 6.0]]).view(6, 1, 3))
            else:
                p.add_(0.3 * torch.randn_like(p))
    eb.update(force=True)
    for n, p in eb.named_parameters():
        out["eb_" + n] = p.detach().numpy()
    out["eb_cdf"], out["eb_len"], out["eb_off"] = (eb._quantized_cdf.numpy(), eb._cdf_length.numpy(),
                                                   eb._offset.numpy())
    np.savez_compressed(os.path.join(HERE, "cdf.npz"), **out)
    return eb


def fp_fixtures(gc, eb):
    out = {}
    g = torch.Generator().manual_seed(5)
    # quantize + build_indexes
    y = (torch.randn(2, 5, 7, 9, generator=g) * 6).float()
    y.view(-1)[:8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.4999, -3.5001])
    mu = torch.randn(2, 5, 7, 9, generator=g)
    sc = torch.exp(torch.rand(2, 5, 7, 9, generator=g) * 9 - 3)
    sc.view(-1)[:4] = torch.tensor([0.0, 0.11, 256.0, 1e9])
    sc.view(-1)[4:68] = gc.scale_table  # values exactly on the thresholds
    out["q_y"], out["q_mu"], out["q_scales"] = y.numpy(), mu.numpy(), sc.numpy()
    out["q_sym"] = gc.quantize(y, "symbols", mu).numpy()
    out["q_sym_nomean"] = gc.quantize(y, "symbols").numpy()
    out["q_deq"] = gc.quantize(y, "dequantize", mu).numpy()
    out["q_idx"] = gc.build_indexes(sc).numpy()
    # Gaussian conditional likelihood forward + autograd backward (eval-mode quantisation excluded:
    # the fixture feeds `outputs` directly to _likelihood + bound, entropy_models.py:650-682)
    yh = (torch.randn(4, 6, 5, 5, generator=g) * 3).requires_grad_()
    s2 = torch.exp(torch.rand(4, 6, 5, 5, generator=g) * 8 - 3).requires_grad_()
    m2 = torch.randn(4, 6, 5, 5, generator=g).requires_grad_()
    lik = gc.likelihood_lower_bound(gc._likelihood(yh, s2, m2))
    gout = torch.randn(lik.shape, generator=g)
    lik.backward(gout)
    out["gcl_y"], out["gcl_s"], out["gcl_m"] = yh.detach().numpy(), s2.detach().numpy(), m2.detach().numpy()
    out["gcl_lik"], out["gcl_gout"] = lik.detach().numpy(), gout.numpy()
    out["gcl_gy"], out["gcl_gs"], out["gcl_gm"] = yh.grad.numpy(), s2.grad.numpy(), m2.grad.numpy()
    # entropy bottleneck likelihood (values laid out N, C, H, W) forward/backward in eval mode pieces
    x = (torch.randn(3, 6, 4, 5, generator=g) * 4).requires_grad_()
    eb.zero_grad()
    v = x.permute(1, 0, 2, 3).reshape(6, 1, -1)
    lik = eb.likelihood_lower_bound(eb._likelihood(v))
    gout = torch.randn(lik.shape, generator=g)
    lik.backward(gout)
    out["ebl_x"] = x.detach().numpy()
    out["ebl_lik"] = lik.detach().reshape(6, 3, 4, 5).permute(1, 0, 2, 3).contiguous().numpy()
    out["ebl_gout"] = gout.reshape(6, 3, 4, 5).permute(1, 0, 2, 3).contiguous().numpy()
    out["ebl_gx"] = x.grad.numpy()
    for n, p in eb.named_parameters():
        if p.grad is not None:
            out["ebl_g" + n] = p.grad.numpy()
    out["eb_loss"] = eb.loss().detach().numpy()
    eb.zero_grad()
    eb.loss().backward()
    out["eb_loss_gquantiles"] = eb.quantiles.grad.numpy()
    # eval forward of the EB (dequantize w/ medians) end to end
    eb.eval()
    xo, lk = eb(x.detach())
    out["ebf_out"], out["ebf_lik"] = xo.detach().numpy(), lk.detach().numpy()
    # GDN / IGDN forward + backward with perturbed parameters
    for inv in (False, True):
        torch.manual_seed(3)
        m = GDN(8, inverse=inv)
        with torch.no_grad():
            m.beta.add_(0.2 * torch.rand_like(m.beta))
            m.gamma.add_(0.05 * torch.rand_like(m.gamma))
        xi = torch.randn(2, 8, 6, 7, generator=g).requires_grad_()
        yo = m(xi)
        go = torch.randn(yo.shape, generator=g)
        yo.backward(go)
        tag = "igdn" if inv else "gdn"
        out[f"{tag}_beta"], out[f"{tag}_gamma"] = m.beta.detach().numpy(), m.gamma.detach().numpy()
        out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_gout"] = xi.detach().numpy(), yo.detach().numpy(), go.numpy()
        out[f"{tag}_gx"], out[f"{tag}_gbeta"], out[f"{tag}_ggamma"] = (xi.grad.numpy(), m.beta.grad.numpy(),
                                                                       m.gamma.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "fp.npz"), **out)


def model_fixtures():
    """Tiny seeded models with amplified last layers (SURVEY.md 8d variant ii) so that symbols and
    scale indexes are not degenerate.  Stores state_dict + input + the reference's outputs."""
    for name, cls, N, M, gy, gs in (("factorized", FactorizedPrior, 8, 12, 48.0, None),
                                    ("hyperprior", ScaleHyperprior, 8, 12, 48.0, 96.0),
                                    ("meanscale", MeanScaleHyperprior, 8, 12, 48.0, 24.0)):
        torch.manual_seed(0)
        net = cls(N, M).eval()
        with torch.no_grad():
            net.g_a[6].weight.mul_(gy)
            net.g_a[6].bias.mul_(gy)
            if gs is not None:
                net.h_s[4].weight.mul_(gs)
                net.h_s[4].bias.mul_(gs)
            for n, p in net.named_parameters():  # break the symmetric GDN / EB init a little
                if "gamma" in n or "beta" in n:
                    p.add_(0.02 * torch.rand_like(p))
        net.update(force=True)
        x = torch.rand(2, 3, 64, 128, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            enc = net.compress(x)
            dec = net.decompress(enc["strings"], enc["shape"])
            fwd = net(x)
        out = {"x": x.numpy(), "x_hat": dec["x_hat"].numpy(), "shape": np.array(enc["shape"]),
               "fwd_x_hat": fwd["x_hat"].numpy(), "N": np.array(N), "M": np.array(M)}
        for k, v in fwd["likelihoods"].items():
            out["fwd_lik_" + k] = v.numpy()
        for li, lst in enumerate(enc["strings"]):
            for bi, s in enumerate(lst):
                out[f"str_{li}_{bi}"] = np.frombuffer(s, np.uint8)
        for k, v in net.state_dict().items():
            out["sd." + k] = v.numpy()
        with torch.no_grad():
            y = net.g_a(x)
            out["y"] = y.numpy()
            if hasattr(net, "h_a"):
                z = net.h_a(torch.abs(y) if cls is ScaleHyperprior else y)
                out["z"] = z.numpy()
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **out)
        print(name, "string bytes:", [[len(s) for s in lst] for lst in enc["strings"]])


if __name__ == "__main__":
    torch.set_num_threads(4)
    gc = coder_fixtures()
    eb = cdf_fixtures(gc)
    fp_fixtures(gc, eb)
    model_fixtures()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
