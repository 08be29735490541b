"""GPU parity tests at BASELINE.json's FULL sizes, through size-independent properties (encode -> decode round
trips, oracle agreement on sampled strings, byte-count invariants) because the CPU oracle cannot code 2^28 symbols
in test time.  Configs: C1 (factorized q1, one 768x512 image), C3 (raw coder 4096 x 65,536 symbols, default Gaussian
table, 0 % and ~9 % escapes), C4 (mbt2018-mean q8 on a 3840x2176 frame: one 10.4 M-symbol string), C5 (hyperprior
training step 16x3x256x256)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _gc_table(golden):
    from compressai_environment_b200 import coder

    c = golden("cdf")
    t = coder.CdfTable(*(torch.from_numpy(c[k]).to(DEV) for k in ("gc_cdf", "gc_len", "gc_off")))
    return t, c


@pytest.mark.parametrize("t,bits_lo,bits_hi", [(1.0, 4.4, 4.7), (4.0, 10.0, 10.8)])
def test_c3_raw_coder_full(golden, orc, t, bits_lo, bits_hi):
    from compressai_environment_b200 import coder

    table, c = _gc_table(golden)
    B, n = 4096, 65536
    g = torch.Generator(device=DEV).manual_seed(1234)
    tab = torch.from_numpy(c["gc_scale_table"]).to(DEV)
    idx = torch.randint(0, 64, (B, n), generator=g, device=DEV, dtype=torch.int32)
    sym = torch.round(torch.randn((B, n), generator=g, device=DEV) * tab[idx.long()] * t).to(torch.int32)
    enc = coder.encode(table, sym, idx)
    assert int(enc.status.max()) == 0
    nw = enc.n_words.cpu().numpy().astype(np.int64)
    bits = nw.sum() * 32 / (B * n)
    assert bits_lo < bits < bits_hi, bits
    # round trip of ALL 2^28 symbols, decoding in place from the encoder's slots
    dec = coder.decode(table, None, idx, device_words=enc.device_words())
    assert torch.equal(dec, sym)
    # oracle agreement on sampled strings (bytes identical), incl. first and last
    strings = enc.to_bytes()
    assert [len(s) for s in strings] == (nw * 4).tolist()
    for b in (0, 1, 777, 2048, 4095):
        ref = orc.rans_encode(sym[b].cpu().numpy(), idx[b].cpu().numpy(), c["gc_cdf"], c["gc_len"], c["gc_off"])
        assert strings[b] == ref
    # decoding from the packed host strings gives the same symbols (spot check to bound H2D time)
    sel = [0, 4095, 1234]
    dec2 = coder.decode(table, [strings[b] for b in sel], idx[sel].contiguous())
    assert torch.equal(dec2, sym[sel])


def test_c1_factorized_full(orc):
    from compressai_environment_b200.zoo import bmshj2018_factorized

    torch.manual_seed(0)
    net = bmshj2018_factorized(1)
    with torch.no_grad():
        net.g_a[6].weight.mul_(64.0)
        net.g_a[6].bias.mul_(64.0)
    net = net.to(DEV).eval()
    assert net.update(force=True)
    x = torch.rand(1, 3, 512, 768, generator=torch.Generator().manual_seed(0)).to(DEV)
    with torch.no_grad():
        enc = net.compress(x)
        dec = net.decompress(enc["strings"], enc["shape"])
        y = net.g_a(x)
        fwd = net(x)
    assert tuple(enc["shape"]) == (32, 48) and dec["x_hat"].shape == x.shape
    eb = net.entropy_bottleneck
    med = eb.quantiles[:, 0, 1].detach().cpu().numpy()
    sym = orc.quantize_symbols(y.cpu().numpy(), med[None, :, None, None])
    assert sym.size == 294912 and np.abs(sym).max() > 3
    idx = np.broadcast_to(np.arange(192, dtype=np.int32)[None, :, None, None], sym.shape)
    tabs = [t.cpu().numpy() for t in (eb._quantized_cdf, eb._cdf_length, eb._offset)]
    assert enc["strings"][0][0] == orc.rans_encode(sym[0], idx[0], *tabs)
    assert (dec["x_hat"] - fwd["x_hat"].clamp(0, 1)).abs().max() <= 2e-3


def test_c4_meanscale_4k_frame(orc):
    """One 3840x2160 frame padded to 3840x2176 (examples/codec.py:227-240): a single 10,444,800-symbol y string."""
    from compressai_environment_b200.zoo import mbt2018_mean

    torch.manual_seed(0)
    net = mbt2018_mean(8)
    with torch.no_grad():
        net.g_a[6].weight.mul_(64.0)
        net.g_a[6].bias.mul_(64.0)
        net.h_s[4].weight.mul_(64.0)
        net.h_s[4].bias.mul_(64.0)
    net = net.to(DEV).eval()
    net.update(force=True)
    x = torch.rand(1, 3, 2176, 3840, generator=torch.Generator().manual_seed(0)).to(DEV)
    with torch.no_grad():
        enc = net.compress(x)
        dec = net.decompress(enc["strings"], enc["shape"])
        y = net.g_a(x)
        z_hat = net.entropy_bottleneck.decompress(enc["strings"][1], enc["shape"])
        scales, means = net._params(z_hat)
    assert tuple(enc["shape"]) == (34, 60) and dec["x_hat"].shape == x.shape
    gc = net.gaussian_conditional
    tabs = [t.cpu().numpy() for t in (gc._quantized_cdf, gc._cdf_length, gc._offset)]
    sym = orc.quantize_symbols(y.cpu().numpy(), means.cpu().numpy())
    idx = orc.gc_build_indexes(scales.cpu().numpy(), gc.scale_table.cpu().numpy())
    assert sym.size == 10444800
    got = orc.rans_decode(enc["strings"][0][0], idx[0], *tabs)      # the ORACLE decodes OUR 10.4 M-symbol string
    assert np.array_equal(got, sym[0].ravel())
    escapes = np.mean((sym - gc._offset.cpu().numpy()[idx] < 0) |
                      (sym - gc._offset.cpu().numpy()[idx] >= gc._cdf_length.cpu().numpy()[idx] - 2))
    assert escapes >= 0.05, escapes                                  # "escape-heavy" as BASELINE asks
    assert enc["strings"][0][0] == orc.rans_encode(sym[0], idx[0], *tabs)
    assert torch.isfinite(dec["x_hat"]).all() and float(dec["x_hat"].min()) >= 0 and float(dec["x_hat"].max()) <= 1


def test_c5_training_step():
    """bmshj2018-hyperprior training-mode forward / backward on 16x3x256x256 (examples/train.py:132-165 step)."""
    from compressai_environment_b200.zoo import bmshj2018_hyperprior

    torch.manual_seed(0)
    net = bmshj2018_hyperprior(4).to(DEV).train()
    params = [p for n, p in net.named_parameters() if not n.endswith(".quantiles")]
    aux_params = [p for n, p in net.named_parameters() if n.endswith(".quantiles")]
    opt = torch.optim.Adam(params, lr=1e-4)
    aux_opt = torch.optim.Adam(aux_params, lr=1e-3)
    x = torch.rand(16, 3, 256, 256, device=DEV)
    losses = []
    for _ in range(2):
        opt.zero_grad()
        aux_opt.zero_grad()
        out = net(x)
        assert out["likelihoods"]["y"].shape == (16, 192, 16, 16) and out["likelihoods"]["z"].shape == (16, 128, 4, 4)
        bpp = sum(torch.log(l).sum() / (-np.log(2) * 16 * 256 * 256) for l in out["likelihoods"].values())
        loss = 0.018 * 255 ** 2 * torch.nn.functional.mse_loss(out["x_hat"], x) + bpp
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        aux = net.aux_loss()
        aux.backward()
        aux_opt.step()
        losses.append(float(loss))
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
    assert np.isfinite(losses).all()
