"""GPU tests of the HEADLINE configuration (BASELINE.json configs[1], "C2": bmshj2018-hyperprior q4, 256 images of
768x512 per GPU, micro-batches of 32, several requests in flight) and of the C ABI's re-entrancy.

* ``test_c2_hyperprior_batch256`` runs the very gate ``bench.py`` runs before timing (``bench.parity_gate``): oracle
  bytes on sampled images, all 512 strings decoded and compared with the encoder's symbols, device-path statuses,
  host path == device path (strings and reconstruction), reconstruction against float64 / float32 torch synthesis.
  Then three requests run concurrently from host threads and must reproduce the single-request result bit for bit
  (reference behaviour being restated: compressai/models/google.py:302-321).
* ``test_abi_reentrancy_mixed_shapes`` hammers compress / decompress from 8 host threads over models whose layers and
  tables need DIFFERENT dynamic shared memory sizes -- the interleaving that broke when the shared-memory attribute
  was set per launch.
"""
import os
import sys
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _c2_model():
    import bench
    from compressai_environment_b200.zoo import bmshj2018_hyperprior

    torch.manual_seed(0)
    net = bmshj2018_hyperprior(quality=4)
    bench.amplify(net)
    net = net.to(DEV).eval()
    net.update(force=True)
    net.micro_batch = 32
    return net


def test_c2_hyperprior_batch256():
    import bench
    from compressai_environment_b200 import coder

    net = _c2_model()
    B = 256
    x_host = bench.make_images(B, seed=0).pin_memory()
    x_dev = x_host.to(DEV)
    rep = bench.parity_gate(net, x_dev, x_host, 32)
    print("C2 gate:", rep)
    assert rep["strings_decoded"] == 2 * B
    assert rep["decoded_symbols_differing"] == 0
    assert rep["oracle_strings_differing"] == 0 and rep["oracle_images"] >= 4
    assert rep["device_path_statuses_nonzero"] == 0
    assert rep["host_vs_device_strings_differing"] == 0
    assert rep["host_vs_device_x_hat_max_abs"] == 0.0
    assert rep["x_hat_vs_fp64_max_abs"] <= rep["x_hat_tolerance"], rep
    assert rep["ok"]
    assert min(rep["oracle_y_bytes"]) > 100_000  # the headline stream is not degenerate (~440 KB per image)

    # >= 3 concurrent requests through the public API, host tensors in / host tensors out
    with torch.no_grad():
        ref = net.compress(x_host)
        ref_y = [bytes(s) for s in ref["strings"][0]]
        ref_hat = net.decompress(ref["strings"], ref["shape"])["x_hat"].cpu()
    outs, errs = [None] * 3, []

    def work(k):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()), torch.no_grad():
                out = torch.empty((B, 3, 512, 768), dtype=torch.float32).pin_memory()
                for _ in range(2):
                    enc = net.compress(x_host)
                    dec = net.decompress(enc["strings"], enc["shape"], out=out)
                torch.cuda.current_stream().synchronize()
                outs[k] = ([bytes(s) for s in enc["strings"][0]], dec["x_hat"])
        except BaseException as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(k,)) for k in range(3)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for ys, xh in outs:
        assert ys == ref_y
        assert torch.equal(xh, ref_hat)


def test_abi_reentrancy_mixed_shapes():
    """8 host threads, 4 different models (channel counts 32/48, 64/96, 128/192 factorized, 48/80 mean-scale): conv
    tiles of different widths and stage counts, EB tables of different sizes and the 186 KB Gaussian table all in
    flight at once.  Every thread's result must equal the single-threaded result of its model."""
    from compressai_environment_b200.models import FactorizedPrior, MeanScaleHyperprior, ScaleHyperprior

    specs = [(ScaleHyperprior, 32, 48, (2, 3, 128, 192)), (ScaleHyperprior, 64, 96, (3, 3, 192, 128)),
             (FactorizedPrior, 128, 192, (1, 3, 256, 256)), (MeanScaleHyperprior, 48, 80, (2, 3, 128, 128))]
    nets, xs, refs = [], [], []
    for i, (cls, N, M, shp) in enumerate(specs):
        torch.manual_seed(i)
        net = cls(N, M)
        with torch.no_grad():
            net.g_a[6].weight.mul_(32.0)
            net.g_a[6].bias.mul_(32.0)
            if hasattr(net, "h_s"):
                net.h_s[4].weight.mul_(64.0)
                net.h_s[4].bias.mul_(64.0)
        net = net.to(DEV).eval()
        net.update(force=True)
        x = torch.rand(*shp, generator=torch.Generator().manual_seed(100 + i)).to(DEV)
        with torch.no_grad():
            enc = net.compress(x)
            dec = net.decompress(enc["strings"], enc["shape"])
        nets.append(net)
        xs.append(x)
        refs.append(([[bytes(s) for s in lst] for lst in enc["strings"]], dec["x_hat"].clone()))
    torch.cuda.synchronize()
    errs = []

    def work(k):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream()), torch.no_grad():
                for it in range(12):
                    j = (k + it) % len(nets)
                    enc = nets[j].compress(xs[j])
                    dec = nets[j].decompress(enc["strings"], enc["shape"])
                    got = [[bytes(s) for s in lst] for lst in enc["strings"]]
                    torch.cuda.current_stream().synchronize()
                    if got != refs[j][0] or not torch.equal(dec["x_hat"], refs[j][1]):
                        errs.append(f"thread {k} iteration {it} model {j}: result differs")
                        return
        except BaseException as e:  # noqa: BLE001
            errs.append(f"thread {k}: {e!r}")

    ts = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs[:3]


def test_truncated_and_corrupt_streams_raise():
    """A string cut short must not decode silently (status 6 used to be dropped)."""
    from compressai_environment_b200.models import ScaleHyperprior

    torch.manual_seed(0)
    net = ScaleHyperprior(32, 48)
    with torch.no_grad():
        net.g_a[6].weight.mul_(48.0)
        net.g_a[6].bias.mul_(48.0)
        net.h_s[4].weight.mul_(96.0)
        net.h_s[4].bias.mul_(96.0)
    net = net.to(DEV).eval()
    net.update(force=True)
    x = torch.rand(2, 3, 128, 192, device=DEV)
    with torch.no_grad():
        enc = net.compress(x)
        ys = [bytes(s) for s in enc["strings"][0]]
        zs = [bytes(s) for s in enc["strings"][1]]
        assert len(ys[0]) > 64
        with pytest.raises(ValueError):
            net.decompress([[ys[0][:len(ys[0]) // 2 // 4 * 4], ys[1]], zs], enc["shape"])
        # intact strings still decode after the failure
        a = net.decompress([ys, zs], enc["shape"])["x_hat"]
        b = net.decompress(enc["strings"], enc["shape"])["x_hat"]
        assert torch.equal(a, b)


def test_cache_invalidation_on_data_edit():
    """Edits through ``.data`` bypass the version counter; update() / invalidate_caches() must still make the fused
    inference path see the new weights (stale packed weights -> silent encoder / decoder mismatch)."""
    import compressai_environment_b200 as cai
    from compressai_environment_b200.models import FactorizedPrior

    torch.manual_seed(0)
    net = FactorizedPrior(32, 48).to(DEV).eval()
    net.update(force=True)
    x = torch.rand(1, 3, 64, 64, device=DEV)
    with torch.no_grad():
        y0 = net.g_a(x).clone()
        net.g_a[6].weight.data.mul_(2.0)
        net.g_a[6].bias.data.mul_(2.0)
        cai.invalidate_caches(net)
        y1 = net.g_a(x)
        assert torch.allclose(y1, 2.0 * y0, rtol=1e-4, atol=1e-5)
        net.g_a[6].weight.data.mul_(0.5)
        net.g_a[6].bias.data.mul_(0.5)
        net.update(force=True)  # update() drops the caches as well
        assert torch.allclose(net.g_a(x), y0, rtol=1e-4, atol=1e-5)
