"""ssf2020 hyperprior (compressai/models/video/google.py:150-196 + QReLU layers.py:247-296) against
a fixture produced by the reference on CPU (tests/golden/make_golden_video.py).  Byte parity is defined at the coder
boundary as for the image models (tests/test_models_gpu.py): latents and Gaussian parameters within a stated fp
tolerance; wherever our symbols / indexes equal the reference's the byte strings are identical; the oracle coder fed
OUR symbols reproduces OUR bytes; decompress(compress(y)) returns the encoder's y_hat exactly."""
import numpy as np
import pytest
import torch

DEV = "cuda"
RTOL = 1e-3


def _load(golden):
    from compressai_environment_b200.models.video import Hyperprior

    g = golden("video_hyperprior")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    planes = sd["hyper_encoder.0.weight"].shape[1]
    net = Hyperprior(planes=planes, mid_planes=sd["hyper_encoder.0.weight"].shape[0])
    net.load_state_dict(sd)
    return net, g, sd


def test_qrelu_matches_reference_fixture(golden):
    from compressai_environment_b200.layers import QReLU

    g = golden("video_hyperprior")
    x = torch.from_numpy(g["qrelu_x"]).requires_grad_(True)
    y = QReLU.apply(x, 8, 100)
    y.backward(torch.from_numpy(g["qrelu_g"]))
    assert np.array_equal(y.detach().numpy(), g["qrelu_y"])
    np.testing.assert_allclose(x.grad.numpy(), g["qrelu_gx"], rtol=1e-6, atol=1e-30)


def test_state_dict_keys_match_reference(golden):
    net, g, sd = _load(golden)
    assert set(net.state_dict().keys()) == set(sd.keys())


@pytest.mark.gpu
def test_video_hyperprior_vs_reference(golden, orc):
    net, g, sd = _load(golden)
    net = net.to(DEV).eval()
    y = torch.from_numpy(g["y"]).to(DEV)
    with torch.no_grad():
        y_hat, enc = net.compress(y)
        dec = net.decompress(enc["strings"], enc["shape"])
        z = net.hyper_encoder(y.contiguous(memory_format=torch.channels_last))
        z_hat = net.entropy_bottleneck.decompress(enc["strings"][1], enc["shape"])
        scales, means = net._params(z_hat)
        fwd_y_hat, lik = net(y)
    assert tuple(enc["shape"]) == tuple(g["shape"])
    assert torch.equal(dec, y_hat), "decompress(compress(y)) must return the encoder's y_hat"

    def close(a, ref, what):
        err = np.abs(a.cpu().numpy() - ref).max()
        assert err <= RTOL * np.abs(ref).max(), (what, err)

    close(z, g["z"], "z")
    ref_strings = [[g[f"str_{li}_{bi}"].tobytes() for bi in range(2)] for li in range(2)]
    med = sd["entropy_bottleneck.quantiles"][:, 0, 1].numpy()[None, :, None, None]
    z_sym = orc.quantize_symbols(z.cpu().numpy(), med)
    z_same = np.array_equal(z_sym, orc.quantize_symbols(g["z"], med))
    ztabs = [sd["entropy_bottleneck." + k].numpy() for k in ("_quantized_cdf", "_cdf_length", "_offset")]
    zidx = np.broadcast_to(np.arange(z_sym.shape[1], dtype=np.int32)[None, :, None, None], z_sym.shape)
    for b in range(2):
        assert enc["strings"][1][b] == orc.rans_encode(z_sym[b], zidx[b], *ztabs)
    if z_same:
        assert [bytes(s) for s in enc["strings"][1]] == ref_strings[1]
        close(scales, g["scales"], "scales")
        close(means, g["means"], "means")
    # y strings: the oracle codes OUR symbols / indexes to OUR bytes, and decodes them back
    gtabs = [sd["gaussian_conditional." + k].numpy() for k in ("_quantized_cdf", "_cdf_length", "_offset")]
    table = sd["gaussian_conditional.scale_table"].numpy()
    y_sym = np.rint(g["y"] - means.cpu().numpy()).astype(np.int32)
    y_idx = orc.gc_build_indexes(scales.cpu().numpy(), table)
    for b in range(2):
        ours = bytes(enc["strings"][0][b])
        assert ours == orc.rans_encode(y_sym[b], y_idx[b], *gtabs)
        assert np.array_equal(orc.rans_decode(ours, y_idx[b], *gtabs), y_sym[b].ravel())
    ref_sym = np.rint(g["y"] - g["means"]).astype(np.int32)
    ref_idx = orc.gc_build_indexes(g["scales"], table)
    if z_same and np.array_equal(y_sym, ref_sym) and np.array_equal(y_idx, ref_idx):
        assert [bytes(s) for s in enc["strings"][0]] == ref_strings[0]
        assert np.abs(y_hat.cpu().numpy() - g["y_hat"]).max() <= 1e-3
    else:  # report how far from identical we are: rounding-boundary flips only
        assert (y_sym != ref_sym).mean() <= 0.002 and (y_idx != ref_idx).mean() <= 0.002
    # training-mode forward (likelihoods) against the reference's
    assert (np.abs(lik["z"].cpu().numpy() - g["fwd_lik_z"]) > 1e-3).mean() <= 0.01
    assert (np.abs(lik["y"].cpu().numpy() - g["fwd_lik_y"]) > 1e-3).mean() <= 0.01
    assert (np.abs(fwd_y_hat.cpu().numpy() - g["fwd_y_hat"]) > 1e-3).mean() <= 0.002


@pytest.mark.gpu
def test_video_hyperprior_training_backward():
    """QReLU + deconv stacks in training mode: gradients flow to every parameter through our kernels."""
    from compressai_environment_b200.models.video import Hyperprior

    torch.manual_seed(0)
    net = Hyperprior(32, 32).to(DEV).train()
    y = torch.randn(2, 32, 32, 32, device=DEV, requires_grad=True)
    y_hat, lik = net(y)
    loss = sum(torch.log(v).sum() for v in lik.values()) * -1e-3 + (y_hat ** 2).mean()
    loss.backward()
    missing = [n for n, p in net.named_parameters() if p.grad is None and not n.endswith("quantiles")]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)
