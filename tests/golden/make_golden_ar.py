"""Golden fixtures for the autoregressive models, generated from the UNMODIFIED reference:

    ./oracle/build_ref.sh && python tests/golden/make_golden_ar.py

* ``ar_jarhp.npz``: JointAutoregressiveHierarchicalPriors(N=16, M=32) (compressai/models/google.py:395-661), two
  128x192 images: state_dict, input, the strings of ``compress`` (``_compress_ar``), ``decompress``'s x_hat, eval-mode
  ``forward`` outputs.
* ``ar_cheng.npz``: Cheng2020Attention(N=16) (compressai/models/waseda.py:113-153), one 64x128 image, same content.
* ``cheng_blocks.npz``: each block of compressai/layers/layers.py:98-244 alone (input, state_dict, output).

Random init gives near-zero latents and constant Gaussian parameters, so -- as for the other fixtures (SURVEY 8d ii) --
the last layers of g_a, h_s and entropy_parameters are scaled by fixed constants first.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402

orc.import_ref()
from compressai.layers import (AttentionBlock, ResidualBlock, ResidualBlockUpsample,  # noqa: E402
                               ResidualBlockWithStride)
from compressai.models import Cheng2020Attention, JointAutoregressiveHierarchicalPriors  # noqa: E402


def amplify(mods_gains):
    with torch.no_grad():
        for m, g in mods_gains:
            m.weight.mul_(g)
            m.bias.mul_(g)


def run_model(net, x, name, seed):
    net.update(force=True)
    with torch.no_grad():
        enc = net.compress(x)
        raw = []
        hook = net.g_s.register_forward_hook(lambda m, i, o: raw.append(float(o.abs().max())))  # before clamp_
        dec = net.decompress(enc["strings"], enc["shape"])
        hook.remove()
        fwd = net(x)
        # intermediate tensors for diagnostics
        y = net.g_a(x)
        z = net.h_a(y)
        z_hat = net.entropy_bottleneck.decompress(enc["strings"][1], enc["shape"])
        params = net.h_s(z_hat)
    # the input is not stored: the test regenerates it from the seed (torch's CPU generator is deterministic)
    out = {"x_shape": np.array(x.shape), "x_seed": np.array(seed), "shape": np.array(enc["shape"]),
           "x_hat": dec["x_hat"].numpy(), "x_hat_raw_absmax": np.array(raw[0]), "y": y.numpy(), "z": z.numpy(), "params": params.numpy(),
           "fwd_lik_y": fwd["likelihoods"]["y"].numpy(), "fwd_lik_z": fwd["likelihoods"]["z"].numpy(),
           "fwd_x_hat_mean": np.array(float(fwd["x_hat"].mean())), "fwd_x_hat_std": np.array(float(fwd["x_hat"].std()))}
    for li, lst in enumerate(enc["strings"]):
        for bi, s in enumerate(lst):
            out[f"str_{li}_{bi}"] = np.frombuffer(s, np.uint8)
    for k, v in net.state_dict().items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    bits = sum(len(s) for lst in enc["strings"] for s in lst) * 8
    print(name, "strings:", [[len(s) for s in lst] for lst in enc["strings"]], "bpp", bits / (x.shape[0] * x.shape[2] * x.shape[3]),
          "y std", float(y.std()), "params std", float(params.std()),
          "raw absmax", raw[0], "x_hat saturated", float(((dec["x_hat"] <= 0) | (dec["x_hat"] >= 1)).float().mean()))


torch.manual_seed(0)
net = JointAutoregressiveHierarchicalPriors(N=16, M=32).eval()
amplify([(net.g_a[6], 120.0), (net.h_a[4], 30.0), (net.h_s[4], 20.0), (net.entropy_parameters[4], 25.0),
         (net.context_prediction, 3.0)])
x = torch.rand(2, 3, 128, 192, generator=torch.Generator().manual_seed(1))
run_model(net, x, "ar_jarhp.npz", 1)

torch.manual_seed(2)
net = Cheng2020Attention(N=16).eval()
amplify([(net.g_a[7], 30.0), (net.h_a[8], 30.0), (net.h_s[8], 20.0), (net.entropy_parameters[4], 25.0),
         (net.context_prediction, 3.0)])
x = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(3))
run_model(net, x, "ar_cheng.npz", 3)

# the blocks alone
out = {}
gen = torch.Generator().manual_seed(4)
for name, blk, shape in (("rbws", ResidualBlockWithStride(3, 16, stride=2), (2, 3, 32, 48)),
                         ("rbws16", ResidualBlockWithStride(16, 32, stride=2), (2, 16, 16, 24)),
                         ("rbu", ResidualBlockUpsample(16, 16, 2), (2, 16, 8, 12)),
                         ("rb", ResidualBlock(16, 32), (2, 16, 16, 24)),
                         ("rb_same", ResidualBlock(16, 16), (2, 16, 16, 24)),
                         ("attn", AttentionBlock(16), (2, 16, 16, 24))):
    blk = blk.eval()
    xin = torch.randn(*shape, generator=gen)
    with torch.no_grad():
        yout = blk(xin)
    out[f"{name}.x"], out[f"{name}.y"] = xin.numpy(), yout.numpy()
    for k, v in blk.state_dict().items():
        out[f"{name}.sd.{k}"] = v.numpy()
np.savez_compressed(os.path.join(HERE, "cheng_blocks.npz"), **out)
print("blocks written")
