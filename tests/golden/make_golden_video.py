"""Golden fixture for the ssf2020 hyperprior, generated from the UNMODIFIED reference:

    ./oracle/build_ref.sh && python tests/golden/make_golden_video.py

The reference's ``Hyperprior`` class is local to ``ScaleSpaceFlow.__init__`` (compressai/models/video/google.py:150);
it is taken from an instance (``type(ssf.img_hyperprior)``) and re-instantiated with 16 planes so that the fixture stays
small.  Stores state_dict, a synthetic latent, the reference's strings / y_hat / likelihoods, and a QReLU
forward / backward fixture (compressai/layers/layers.py:247-296).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402

orc.import_ref()
from compressai.layers import QReLU  # noqa: E402
from compressai.models.video import ScaleSpaceFlow  # noqa: E402

torch.manual_seed(0)
cls = type(ScaleSpaceFlow().img_hyperprior)
torch.manual_seed(1)
net = cls(planes=16, mid_planes=16).eval()
with torch.no_grad():  # random init gives constant scales / means: amplify the last decoder layers (SURVEY 8d ii)
    for m, g in ((net.hyper_decoder_scale.deconv3, 400.0), (net.hyper_decoder_mean[4], 60.0), (net.hyper_encoder[4], 30.0)):
        m.weight.mul_(g)
        m.bias.mul_(g)
from compressai.models.google import get_scale_table  # noqa: E402

net.gaussian_conditional.update_scale_table(get_scale_table(), force=True)  # as ScaleSpaceFlow.update does
net.update(force=True)
y = torch.randn(2, 16, 32, 48, generator=torch.Generator().manual_seed(2)) * 3.0
with torch.no_grad():
    y_hat, enc = net.compress(y)
    dec = net.decompress(enc["strings"], enc["shape"])
    fwd_y_hat, lik = net(y)
    z = net.hyper_encoder(y)
    z_hat = net.entropy_bottleneck.decompress(enc["strings"][1], enc["shape"])
    scales, means = net.hyper_decoder_scale(z_hat), net.hyper_decoder_mean(z_hat)
assert torch.equal(dec, y_hat)
out = {"y": y.numpy(), "y_hat": y_hat.numpy(), "shape": np.array(enc["shape"]), "z": z.numpy(), "scales": scales.numpy(),
       "means": means.numpy(), "fwd_y_hat": fwd_y_hat.numpy(), "fwd_lik_y": lik["y"].numpy(), "fwd_lik_z": lik["z"].numpy()}
for li, lst in enumerate(enc["strings"]):
    for bi, s in enumerate(lst):
        out[f"str_{li}_{bi}"] = np.frombuffer(s, np.uint8)
for k, v in net.state_dict().items():
    out["sd." + k] = v.numpy()
# QReLU fixture
x = (torch.rand(4, 37, generator=torch.Generator().manual_seed(3)) * 400 - 70).requires_grad_(True)
g = torch.randn(4, 37, generator=torch.Generator().manual_seed(4))
q = QReLU.apply(x, 8, 100)
q.backward(g)
out["qrelu_x"], out["qrelu_g"], out["qrelu_y"], out["qrelu_gx"] = x.detach().numpy(), g.numpy(), q.detach().numpy(), x.grad.numpy()
np.savez_compressed(os.path.join(HERE, "video_hyperprior.npz"), **out)
print("strings:", [[len(s) for s in lst] for lst in enc["strings"]], "scale range", float(scales.min()), float(scales.max()),
      "rows used", len(np.unique(net.gaussian_conditional.build_indexes(scales).numpy())))
