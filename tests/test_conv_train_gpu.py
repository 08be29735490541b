"""GPU tests of the training-mode convolutions (transforms._ConvFunction): forward on the implicit-GEMM kernel, data
gradient as the other layer kind's forward, weight gradient on the tcgen05 pixel-reduction GEMM (csrc/wgrad.cu,
``cai_conv_wgrad``) -- against torch autograd of F.conv2d / F.conv_transpose2d in fp32 with TF32 off (what the
reference runs: nn.Conv2d / nn.ConvTranspose2d built by compressai/models/utils.py:128-146).
Tolerance: max-abs error <= 3e-4 of max |ref| per tensor (split-bf16 operands, fp32 accumulation)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 3e-4


def _close(got, ref, what, rtol=RTOL):
    got, ref = got.float(), ref.float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= rtol * max(scale, 1e-20), (what, err, scale)


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


# small tensor = dY [N, Cs, Hs, Ws], big tensor = X [N, Cb, Hb, Wb]: every box shape (Ws >= 64, 32, 16, 8, ragged),
# both strides, 1 / 2 m-tiles, 1 / 2 n-tiles, 3-channel operands on either side
@pytest.mark.parametrize("N,Cs,Cb,k,s,hw", [(2, 128, 128, 5, 2, (16, 64)), (3, 128, 128, 5, 2, (8, 32)),
                                            (2, 192, 128, 5, 2, (16, 16)), (2, 128, 192, 3, 1, (16, 16)),
                                            (1, 128, 3, 5, 2, (24, 40)), (2, 3, 128, 5, 2, (12, 20)),
                                            (2, 64, 48, 5, 2, (3, 5)), (1, 320, 272, 3, 1, (9, 7)),
                                            (4, 128, 128, 5, 2, (4, 4)), (1, 128, 128, 5, 2, (70, 130))])
def test_wgrad_kernel(N, Cs, Cb, k, s, hw):
    from compressai_environment_b200 import transforms as T

    torch.manual_seed(Cs * 7 + Cb + k + s + hw[0])
    p = k // 2
    Hs, Ws = hw
    Hb, Wb = (Hs - 1) * s + k - 2 * p + (s - 1), (Ws - 1) * s + k - 2 * p + (s - 1)   # what a transposed conv emits
    small = torch.randn(N, Cs, Hs, Ws, device=DEV)
    big = torch.randn(N, Cb, Hb, Wb, device=DEV)
    got = T.conv_wgrad(small, big, k, s, p)
    # reference: weight gradient of conv2d(big, W[Cs, Cb, k, k]) for upstream `small` (float64)
    w = torch.zeros(Cs, Cb, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(big.double(), w, None, stride=s, padding=p)
    assert out.shape[-2:] == small.shape[-2:], (out.shape, small.shape)
    (ref,) = torch.autograd.grad(out, w, small.double())
    _close(got, ref, "wgrad")
    again = T.conv_wgrad(small, big, k, s, p)
    assert torch.equal(got, again), "split-K reduction must be deterministic"


@pytest.mark.parametrize("cin,cout,k,s,hw,xgrad", [(128, 128, 5, 2, (32, 32), True), (3, 128, 5, 2, (64, 48), False),
                                                   (128, 192, 5, 2, (16, 24), True), (192, 128, 3, 1, (16, 16), True),
                                                   (128, 128, 5, 2, (7, 9), True), (64, 96, 5, 2, (20, 12), True)])
def test_conv_training_matches_torch(cin, cout, k, s, hw, xgrad):
    from compressai_environment_b200.transforms import Conv2d

    torch.manual_seed(cin + cout + hw[0])
    m = Conv2d(cin, cout, kernel_size=k, stride=s).to(DEV)
    x = torch.randn(2, cin, *hw, device=DEV, requires_grad=xgrad)
    y = m(x)
    g = torch.randn_like(y)
    grads = torch.autograd.grad(y, ([x] if xgrad else []) + [m.weight, m.bias], g)
    xr = x.detach().clone().requires_grad_(xgrad)
    yr = F.conv2d(xr, m.weight, m.bias, stride=s, padding=k // 2)
    refs = torch.autograd.grad(yr, ([xr] if xgrad else []) + [m.weight, m.bias], g)
    _close(y, yr, "forward")
    for name, a, b in zip((["dx"] if xgrad else []) + ["dw", "db"], grads, refs):
        _close(a, b, name)


@pytest.mark.parametrize("cin,cout,k,s,hw", [(192, 128, 5, 2, (8, 8)), (128, 128, 5, 2, (16, 12)),
                                             (128, 3, 5, 2, (32, 24)), (128, 128, 5, 2, (4, 4)),
                                             (96, 64, 5, 2, (5, 11))])
def test_deconv_training_matches_torch(cin, cout, k, s, hw):
    from compressai_environment_b200.transforms import ConvTranspose2d

    torch.manual_seed(cin + cout + hw[0])
    m = ConvTranspose2d(cin, cout, kernel_size=k, stride=s, output_padding=s - 1).to(DEV)
    x = torch.randn(2, cin, *hw, device=DEV, requires_grad=True)
    y = m(x)
    g = torch.randn_like(y)
    grads = torch.autograd.grad(y, [x, m.weight, m.bias], g)
    xr = x.detach().clone().requires_grad_(True)
    yr = F.conv_transpose2d(xr, m.weight, m.bias, stride=s, padding=k // 2, output_padding=s - 1)
    refs = torch.autograd.grad(yr, [xr, m.weight, m.bias], g)
    _close(y, yr, "forward")
    for name, a, b in zip(["dx", "dw", "db"], grads, refs):
        _close(a, b, name)


def test_hyperprior_training_step_gradients_match_torch():
    """Whole-model check (config C5 at reduced size): loss and EVERY parameter gradient of one RateDistortionLoss
    forward/backward through our kernels == the same model with its conv layers on torch (cuDNN fp32) autograd, same
    noise (same CUDA generator state)."""
    from compressai_environment_b200 import transforms as T
    from compressai_environment_b200.models import ScaleHyperprior
    from compressai_environment_b200.training import RateDistortionLoss

    torch.manual_seed(3)
    net = ScaleHyperprior(64, 96).to(DEV).train()
    x = torch.rand(4, 3, 128, 128, device=DEV)
    crit = RateDistortionLoss(1e-2)

    def run(use_torch_convs):
        saved = (T.Conv2d.forward, T.ConvTranspose2d.forward)
        if use_torch_convs:
            T.Conv2d.forward = lambda self, x: F.conv2d(x, self.weight, self.bias, stride=self.stride, padding=self.padding)
            T.ConvTranspose2d.forward = lambda self, x: F.conv_transpose2d(
                x, self.weight, self.bias, stride=self.stride, padding=self.padding, output_padding=self.output_padding)
        try:
            net.zero_grad(set_to_none=True)
            torch.manual_seed(11)  # same quantisation noise in both runs
            out = crit(net(x), x)
            out["loss"].backward()
            return out["loss"].item(), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
        finally:
            T.Conv2d.forward, T.ConvTranspose2d.forward = saved

    loss_a, g_a = run(False)
    loss_b, g_b = run(True)
    assert abs(loss_a - loss_b) <= 1e-4 * abs(loss_b), (loss_a, loss_b)
    assert set(g_a) == set(g_b) and len(g_a) > 20
    for n in g_b:
        _close(g_a[n], g_b[n], n, rtol=2e-3)
