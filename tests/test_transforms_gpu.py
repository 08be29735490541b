"""GPU tests of the tcgen05 implicit-GEMM transforms against a plain PyTorch fp32 reference of the same op
(torch.nn.functional conv2d / conv_transpose2d with TF32 disabled, GDN formula of layers/gdn.py:77-92).
Tolerance: max-abs error <= 2e-4 of max |ref| (operands carry ~16 mantissa bits: bf16 hi + lo planes)."""
import numpy as np
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


@pytest.mark.parametrize("cin,cout,k,s,hw", [(128, 128, 5, 2, (16, 24)), (192, 128, 3, 1, (8, 12)), (64, 192, 5, 2, (9, 7)),
                                             (128, 320, 5, 2, (12, 12)), (3, 128, 5, 2, (32, 48)), (16, 16, 3, 1, (5, 5)),
                                             (320, 192, 3, 1, (6, 10))])
def test_conv(cin, cout, k, s, hw):
    from compressai_environment_b200.transforms import Conv2d, run_stack

    torch.manual_seed(0)
    m = Conv2d(cin, cout, k, s).to(DEV)
    x = torch.randn(2, cin, *hw, device=DEV)
    with torch.no_grad():
        ref = F.conv2d(x, m.weight, m.bias, stride=s, padding=k // 2)
        _check(run_stack([m], x), ref)
        _check(run_stack([m], x.contiguous(memory_format=torch.channels_last)), ref)
        _check(run_stack([m, torch.nn.ReLU()], x), F.relu(ref))
        _check(run_stack([m, torch.nn.LeakyReLU()], x), F.leaky_relu(ref))


@pytest.mark.parametrize("cin,cout,hw", [(128, 128, (8, 12)), (192, 128, (4, 6)), (128, 3, (16, 24)), (64, 96, (5, 3)),
                                         (128, 480, (4, 4))])
def test_deconv(cin, cout, hw):
    from compressai_environment_b200.transforms import ConvTranspose2d, run_stack

    torch.manual_seed(1)
    m = ConvTranspose2d(cin, cout, 5, 2).to(DEV)
    x = torch.randn(2, cin, *hw, device=DEV)
    with torch.no_grad():
        ref = F.conv_transpose2d(x, m.weight, m.bias, stride=2, padding=2, output_padding=1)
        _check(run_stack([m], x), ref)
        if cout == 3:
            _check(run_stack([m], x, clamp=(0.0, 1.0), nchw_out=True), ref.clamp(0, 1))
        else:
            _check(run_stack([m, torch.nn.ReLU()], x), F.relu(ref))


@pytest.mark.parametrize("hw", [(67, 131), (5, 9), (64, 130), (33, 200)])
def test_first_and_last_layer_tiled_patch_kernels(hw, monkeypatch):
    """Cin = 3 / Cout = 3 layers go through the shared-memory tiled im2col / col2im kernels; ragged tile edges, both
    input layouts, both output layouts, and agreement with the generic gather kernels (CAI_PATCH_GENERIC=1)."""
    from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d, run_stack

    torch.manual_seed(3)
    c1 = Conv2d(3, 32, 5, 2).to(DEV)
    d1 = ConvTranspose2d(32, 3, 5, 2).to(DEV)
    x = torch.rand(3, 3, *hw, device=DEV)
    z = torch.randn(3, 32, *hw, device=DEV)
    with torch.no_grad():
        ref_c = F.conv2d(x, c1.weight, c1.bias, stride=2, padding=2)
        ref_d = F.conv_transpose2d(z, d1.weight, d1.bias, stride=2, padding=2, output_padding=1)
        got_c = run_stack([c1], x)
        got_c_cl = run_stack([c1], x.contiguous(memory_format=torch.channels_last))
        got_d = run_stack([d1], z)
        got_d_clamp = run_stack([d1], z, clamp=(0.0, 1.0), nchw_out=True)
        _check(got_c, ref_c)
        _check(got_c_cl, ref_c)
        _check(got_d, ref_d)
        _check(got_d_clamp, ref_d.clamp(0, 1))
        assert got_d_clamp.is_contiguous()
        monkeypatch.setenv("CAI_PATCH_GENERIC", "1")
        gen_c = run_stack([c1], x)
        gen_d = run_stack([d1], z)
        assert torch.equal(gen_c, got_c)          # same planes -> same GEMM -> identical
        _check(gen_d, got_d, rtol=1e-6)           # same addends, different summation order


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("C", [128, 192, 32])
def test_gdn(C, inverse):
    from compressai_environment_b200.layers import GDN

    torch.manual_seed(2)
    m = GDN(C, inverse=inverse).to(DEV)
    with torch.no_grad():
        m.beta.add_(0.2 * torch.rand_like(m.beta))
        m.gamma.add_(0.05 * torch.rand_like(m.gamma))
    x = torch.randn(2, C, 6, 7, device=DEV) * 2
    with torch.no_grad():
        beta, gamma = m.effective_params()
        norm = F.conv2d(x.double() ** 2, gamma.double().reshape(C, C, 1, 1), beta.double())
        ref = x.double() * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))
        _check(m(x), ref.float())


def test_gdn_golden(golden):
    from compressai_environment_b200.layers import GDN

    f = golden("fp")
    for tag, inv in (("gdn", False), ("igdn", True)):
        m = GDN(8, inverse=inv).to(DEV)
        with torch.no_grad():
            m.beta.copy_(torch.from_numpy(f[tag + "_beta"]))
            m.gamma.copy_(torch.from_numpy(f[tag + "_gamma"]))
            # also exercise explicit zero-padding to 16 channels through a wider layer
            m16 = GDN(16, inverse=inv).to(DEV)
            m16.beta[:8] = m.beta
            m16.gamma.fill_(m16.gamma_reparam.init(torch.zeros(1, device=DEV)).item())
            m16.gamma[:8, :8] = m.gamma
            x = torch.zeros(2, 16, 6, 7, device=DEV)
            x[:, :8] = torch.from_numpy(f[tag + "_x"]).to(DEV)
            y = m16(x)[:, :8]
        _check(y, torch.from_numpy(f[tag + "_y"]).to(DEV))
        with torch.no_grad():
            _check(m(torch.from_numpy(f[tag + "_x"]).to(DEV)), torch.from_numpy(f[tag + "_y"]).to(DEV))


def test_odd_channel_counts():
    """Channel counts that are not multiples of 16 are zero padded by the host layer."""
    from compressai_environment_b200.transforms import Conv2d, ConvTranspose2d, run_stack

    torch.manual_seed(5)
    c1, c2, d1 = Conv2d(3, 8).to(DEV), Conv2d(8, 12).to(DEV), ConvTranspose2d(12, 8).to(DEV)
    x = torch.rand(2, 3, 32, 48, device=DEV)
    with torch.no_grad():
        ref = F.conv2d(F.relu(F.conv2d(x, c1.weight, c1.bias, stride=2, padding=2)), c2.weight, c2.bias, stride=2, padding=2)
        got = run_stack([c1, torch.nn.ReLU(), c2], x)
        _check(got, ref)
        _check(run_stack([d1], ref), F.conv_transpose2d(ref, d1.weight, d1.bias, stride=2, padding=2, output_padding=1))


def test_full_stack_matches_torch():
    """g_a-like stack (conv, GDN, conv, GDN, conv) + abs planes + a following h_a-like stack fed by planes."""
    from compressai_environment_b200.layers import GDN
    from compressai_environment_b200.transforms import Conv2d, TransformStack, run_stack

    torch.manual_seed(3)
    ga = TransformStack(Conv2d(3, 32), GDN(32), Conv2d(32, 32), GDN(32), Conv2d(32, 48)).to(DEV)
    ha = TransformStack(Conv2d(48, 32, 3, 1), torch.nn.ReLU(inplace=True), Conv2d(32, 32)).to(DEV)
    x = torch.rand(2, 3, 64, 96, device=DEV)
    with torch.no_grad():
        y, yabs = run_stack(list(ga), x, want_abs=True)
        z = run_stack(list(ha), yabs)
        # torch reference with the same parameters
        r = x
        for m in ga:
            if isinstance(m, GDN):
                beta, gamma = m.effective_params()
                r = r * torch.rsqrt(F.conv2d(r * r, gamma.reshape(32, 32, 1, 1), beta))
            else:
                r = F.conv2d(r, m.weight, m.bias, stride=m.stride, padding=m.padding)
        rz = F.conv2d(F.relu(F.conv2d(r.abs(), ha[0].weight, ha[0].bias, padding=1)), ha[2].weight, ha[2].bias, stride=2,
                      padding=2)
    _check(y, r, 5e-4)
    _check(z, rz, 5e-4)


def test_gdn_backward_golden(golden):
    """GDN / IGDN forward + backward (through the non-negative reparametrisation) vs the reference autograd fixture."""
    from compressai_environment_b200.layers import GDN

    f = golden("fp")
    for tag, inv in (("gdn", False), ("igdn", True)):
        m = GDN(8, inverse=inv).to(DEV)
        with torch.no_grad():
            m.beta.copy_(torch.from_numpy(f[tag + "_beta"]))
            m.gamma.copy_(torch.from_numpy(f[tag + "_gamma"]))
        x = torch.from_numpy(f[tag + "_x"]).to(DEV).requires_grad_()
        y = m(x)
        _check(y, torch.from_numpy(f[tag + "_y"]).to(DEV))
        y.backward(torch.from_numpy(f[tag + "_gout"]).to(DEV))
        _check(x.grad, torch.from_numpy(f[tag + "_gx"]).to(DEV), 5e-4)
        _check(m.beta.grad, torch.from_numpy(f[tag + "_gbeta"]).to(DEV), 5e-4)
        _check(m.gamma.grad, torch.from_numpy(f[tag + "_ggamma"]).to(DEV), 5e-4)


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_backward_vs_torch_autograd(inverse):
    from compressai_environment_b200.layers import GDN

    torch.manual_seed(7)
    C = 128
    m = GDN(C, inverse=inverse).to(DEV)
    with torch.no_grad():
        m.beta.add_(0.2 * torch.rand_like(m.beta))
        m.gamma.add_(0.05 * torch.rand_like(m.gamma))
    x = (torch.randn(2, C, 9, 11, device=DEV) * 2).requires_grad_()
    go = torch.randn(2, C, 9, 11, device=DEV)
    m(x).backward(go)
    got = (x.grad.clone(), m.beta.grad.clone(), m.gamma.grad.clone())
    x.grad = None
    m.zero_grad()
    beta, gamma = m.effective_params()
    xd = x.double()
    norm = F.conv2d(xd * xd, gamma.double().reshape(C, C, 1, 1), beta.double())
    (xd * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))).backward(go.double())
    _check(got[0], x.grad, 5e-4)
    _check(got[1], m.beta.grad, 5e-4)
    _check(got[2], m.gamma.grad, 5e-4)
