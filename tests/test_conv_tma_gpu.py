"""GPU tests of the persistent TMA-fed transform kernel (csrc/conv_tma.cu) against a plain PyTorch fp32 reference of
the same op (conv2d / conv_transpose2d with TF32 off, GDN / IGDN by the definition of layers/gdn.py:77-92), at grid
widths the kernel takes (>= 64 pixels per row) -- the narrow shapes of test_transforms_gpu.py stay on the per-tile
kernel.  Every case asserts that the launches really went to the TMA kernel (``mode == 1``).  Tolerance as for the
per-tile kernel: max-abs error <= 2e-4 of max |ref| (split-bf16 operands, fp32 accumulation)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 2e-4


def _check(got, ref, rtol=RTOL):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= rtol * scale, (err, scale)


def _run(mods, x, expect_tma, **kw):
    from compressai_environment_b200 import transforms as T

    T.TIMING, T.DETAIL = {}, []
    policy, T.TMA_POLICY = T.TMA_POLICY, "all"
    try:
        out = T.run_stack(mods, x, **kw)
        torch.cuda.synchronize()
        labels = [d[0] for d in T.DETAIL]
    finally:
        T.TIMING, T.DETAIL, T.TMA_POLICY = None, None, policy
    n_tma = sum(lbl.endswith(" tma") for lbl in labels)
    assert n_tma == expect_tma, labels
    return out


def _gdn(C, inverse, seed):
    from compressai_environment_b200.layers import GDN

    torch.manual_seed(seed)
    m = GDN(C, inverse=inverse).to(DEV)
    with torch.no_grad():
        m.beta.add_(0.2 * torch.rand_like(m.beta))
        m.gamma.add_(0.05 * torch.rand_like(m.gamma))
    return m


def _gdn_ref(m, x):
    beta, gamma = m.effective_params()
    C = beta.numel()
    norm = F.conv2d(x.double() ** 2, gamma.double().reshape(C, C, 1, 1), beta.double())
    return (x.double() * (torch.sqrt(norm) if m.inverse else torch.rsqrt(norm))).float()


# (cin, cout, k, stride, input H x W): output widths 96 (one segment), 192 (two of 96), 100 (ragged), 64
@pytest.mark.parametrize("cin,cout,k,s,hw", [(128, 128, 5, 2, (20, 192)), (128, 128, 5, 2, (12, 384)),
                                             (64, 96, 5, 2, (9, 200)), (128, 128, 3, 1, (7, 96)),
                                             (192, 128, 3, 1, (5, 70)), (48, 64, 5, 2, (6, 128))])
def test_conv_tma(cin, cout, k, s, hw):
    from compressai_environment_b200.transforms import Conv2d

    torch.manual_seed(0)
    m = Conv2d(cin, cout, k, s).to(DEV)
    x = torch.randn(3, cin, *hw, device=DEV)
    with torch.no_grad():
        ref = F.conv2d(x, m.weight, m.bias, stride=s, padding=k // 2)
        _check(_run([m], x, 1), ref)                                   # fp32 output (KIND 3)
        _check(_run([m, torch.nn.ReLU()], x, 1), F.relu(ref))
        g = _gdn(cout, False, 5)
        _check(_run([m, g], x, 0), _gdn_ref(g, ref))                   # GDN as the LAST module: fp32 out, per-tile kernel
        # as an inner layer: planes out (KIND 1 / 2) feeding a second conv
        m2 = Conv2d(cout, 32, 3, 1).to(DEV)
        ref2 = F.conv2d(_gdn_ref(g, ref), m2.weight, m2.bias, stride=1, padding=1)
        _check(_run([m, g, m2], x, 2 if ref.shape[-1] >= 64 else 1), ref2, rtol=4e-4)
        ref3 = F.conv2d(F.leaky_relu(ref), m2.weight, m2.bias, stride=1, padding=1)
        _check(_run([m, torch.nn.LeakyReLU(), m2], x, 2 if ref.shape[-1] >= 64 else 1), ref3, rtol=4e-4)


@pytest.mark.parametrize("cin,cout,hw", [(128, 128, (6, 96)), (192, 128, (5, 64)), (128, 128, (4, 192)), (64, 48, (7, 100))])
def test_deconv_tma(cin, cout, hw):
    from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d

    torch.manual_seed(1)
    m = ConvTranspose2d(cin, cout, 5, 2).to(DEV)
    x = torch.randn(2, cin, *hw, device=DEV)
    with torch.no_grad():
        ref = F.conv_transpose2d(x, m.weight, m.bias, stride=2, padding=2, output_padding=1)
        _check(_run([m], x, 4), ref)                                   # four phases, fp32 output
        g = _gdn(cout, True, 7)
        m2 = Conv2d(cout, 32, 3, 1).to(DEV)
        ref2 = F.conv2d(_gdn_ref(g, ref), m2.weight, m2.bias, stride=1, padding=1)
        _check(_run([m, g, m2], x, 5), ref2, rtol=4e-4)                # fused IGDN -> planes -> 3x3 conv
        ref3 = F.conv2d(F.relu(ref), m2.weight, m2.bias, stride=1, padding=1)
        _check(_run([m, torch.nn.ReLU(), m2], x, 5), ref3, rtol=4e-4)


def test_first_and_last_layer_tma():
    """The 3-channel first layer (im2col + 1x1 GEMM + fused GDN) and last layer (1x1 GEMM to 80 columns + col2im)."""
    from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d

    torch.manual_seed(2)
    c1 = Conv2d(3, 128, 5, 2).to(DEV)
    g = _gdn(128, False, 9)
    x = torch.rand(2, 3, 64, 256, device=DEV)
    d1 = ConvTranspose2d(128, 3, 5, 2).to(DEV)
    z = torch.randn(2, 128, 16, 128, device=DEV)
    m2 = Conv2d(128, 16, 3, 1).to(DEV)
    with torch.no_grad():
        ref = _gdn_ref(g, F.conv2d(x, c1.weight, c1.bias, stride=2, padding=2))
        ref = F.conv2d(ref, m2.weight, m2.bias, stride=1, padding=1)
        _check(_run([c1, g, m2], x, 2), ref, rtol=4e-4)
        refd = F.conv_transpose2d(z, d1.weight, d1.bias, stride=2, padding=2, output_padding=1).clamp(0, 1)
        _check(_run([d1], z, 1, clamp=(0.0, 1.0), nchw_out=True), refd)


def test_many_tiles_persistent_loop():
    """More tiles than SMs (several tiles per persistent CTA, both accumulator sets reused many times)."""
    from compressai_environment_b200.transforms import Conv2d

    torch.manual_seed(3)
    m = Conv2d(128, 128, 5, 2).to(DEV)
    g = _gdn(128, False, 11)
    m2 = Conv2d(128, 16, 3, 1).to(DEV)
    x = torch.randn(4, 128, 128, 192, device=DEV)   # 4 x 64 rows x 1 segment = 256 tiles... x 96 wide
    with torch.no_grad():
        a = F.conv2d(x, m.weight, m.bias, stride=2, padding=2)
        ref = F.conv2d(_gdn_ref(g, a), m2.weight, m2.bias, stride=1, padding=1)
        _check(_run([m, g, m2], x, 2), ref, rtol=4e-4)
