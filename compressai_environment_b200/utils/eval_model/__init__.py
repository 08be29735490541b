"""Evaluation protocol of ``python -m compressai.utils.eval_model`` -- SURVEY.md 8(f) rank 1, the caller of the hot
path whose timings the reference publishes.  Mirrors (reference tree, ``compressai/utils/eval_model/__main__.py``):

  :81-83    psnr            -10 log10(mse)
  :92-139   inference       centre pad to 64, compress (timed), decompress (timed), crop, PSNR / MS-SSIM / bpp
  :142-160  inference_entropy_estimation   forward(), bpp from the likelihoods
  :173-189  eval_model      mean of the per-image metrics

Differences, all deliberate: timings bracket the calls with a device synchronise (the reference's CUDA numbers only
include whatever the ``.tolist()`` round trips happen to wait for); ``eval_model`` also accepts image tensors, not only
file paths; ``half=True`` is rejected (the transforms compute in split bf16 with fp32-grade accuracy; there is no fp16
path).  MS-SSIM: the reference calls the third-party ``pytorch_msssim`` (not vendored, not installed here); ``ms_ssim``
below restates the published algorithm (Wang et al. 2003: 11-tap sigma-1.5 Gaussian window, five scales with weights
0.0448 / 0.2856 / 0.3001 / 0.2363 / 0.1333, K1 = 0.01, K2 = 0.03, 2x average-pool between scales).
"""
import math
import os
import time
from collections import defaultdict
from typing import Dict, Iterable, Union

import torch
import torch.nn.functional as F

from ... import codec_io

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")
_MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def collect_images(rootpath: str):
    return [os.path.join(rootpath, f) for f in os.listdir(rootpath) if os.path.splitext(f)[-1].lower() in IMG_EXTENSIONS]


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = F.mse_loss(a, b).item()
    return -10 * math.log10(mse)


def _gauss_filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    C = x.size(1)
    k = win.to(x.dtype).to(x.device)
    x = F.conv2d(x, k.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)   # "valid": no padding
    return F.conv2d(x, k.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)


def _ssim_cs(x, y, win, data_range):
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _gauss_filter(x, win), _gauss_filter(y, win)
    s11 = _gauss_filter(x * x, win) - mu1 * mu1
    s22 = _gauss_filter(y * y, win) - mu2 * mu2
    s12 = _gauss_filter(x * y, win) - mu1 * mu2
    cs = (2 * s12 + c2) / (s11 + s22 + c2)
    ssim = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
    return ssim.flatten(2).mean(-1), cs.flatten(2).mean(-1)   # per (image, channel)


def ms_ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """Multi-scale SSIM of two [N, C, H, W] tensors (mean over images and channels); min(H, W) must exceed 160."""
    if x.shape != y.shape or x.dim() != 4:
        raise ValueError("ms_ssim expects two [N, C, H, W] tensors of equal shape")
    if min(x.shape[-2:]) <= (11 - 1) * 2 ** 4:
        raise ValueError("image too small for 5-scale MS-SSIM with an 11-tap window (needs min(H, W) > 160)")
    coords = torch.arange(11, dtype=torch.float32) - 5
    win = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2))
    win = win / win.sum()
    x, y = x.float(), y.float()
    mcs = []
    for i in range(5):
        ssim, cs = _ssim_cs(x, y, win, data_range)
        if i < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            x = F.avg_pool2d(x, kernel_size=2, padding=pad)
            y = F.avg_pool2d(y, kernel_size=2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(ssim)], dim=0)            # [5, N, C]
    w = torch.tensor(_MS_WEIGHTS, dtype=vals.dtype, device=vals.device).view(-1, 1, 1)
    return torch.prod(vals ** w, dim=0).mean()


def _sync(x):
    if x.is_cuda:
        torch.cuda.synchronize(x.device)


@torch.no_grad()
def inference(model, x: torch.Tensor) -> Dict[str, float]:
    """One image [3, h, w] in [0, 1] through the real codec path; same keys as the reference."""
    x = x.unsqueeze(0)
    h, w = x.size(2), x.size(3)
    x_padded = codec_io.pad(x, 64)
    _sync(x)
    start = time.time()
    out_enc = model.compress(x_padded)
    _sync(x)
    enc_time = time.time() - start
    start = time.time()
    out_dec = model.decompress(out_enc["strings"], out_enc["shape"])
    _sync(x)
    dec_time = time.time() - start
    x_hat = codec_io.crop(out_dec["x_hat"], (h, w))
    num_pixels = x.size(0) * h * w
    bpp = sum(len(s[0]) for s in out_enc["strings"]) * 8.0 / num_pixels
    rv = {"psnr": psnr(x, x_hat), "bpp": bpp, "encoding_time": enc_time, "decoding_time": dec_time}
    if min(h, w) > 160:
        rv["ms-ssim"] = ms_ssim(x, x_hat, data_range=1.0).item()
    return rv


@torch.no_grad()
def inference_entropy_estimation(model, x: torch.Tensor) -> Dict[str, float]:
    x = x.unsqueeze(0)
    _sync(x)
    start = time.time()
    out_net = model.forward(x)
    _sync(x)
    elapsed = time.time() - start
    num_pixels = x.size(0) * x.size(2) * x.size(3)
    bpp = sum((torch.log(lk).sum() / (-math.log(2) * num_pixels)) for lk in out_net["likelihoods"].values())
    return {"psnr": psnr(x, out_net["x_hat"]), "bpp": bpp.item(), "encoding_time": elapsed / 2.0,
            "decoding_time": elapsed / 2.0}


def read_image(filepath: str) -> torch.Tensor:
    from PIL import Image
    import numpy as np

    img = np.asarray(Image.open(filepath).convert("RGB"), dtype=np.uint8)
    return torch.from_numpy(img.copy()).permute(2, 0, 1).float().div(255.0)   # == torchvision ToTensor for uint8 RGB


def eval_model(model, images: Iterable[Union[str, torch.Tensor]], entropy_estimation: bool = False, half: bool = False):
    if half:
        raise ValueError("half precision is not supported: the transforms already run on tensor cores (split bf16)")
    device = next(model.parameters()).device
    metrics, n = defaultdict(float), 0
    for item in images:
        x = (read_image(item) if isinstance(item, str) else item).to(device)
        rv = inference_entropy_estimation(model, x) if entropy_estimation else inference(model, x)
        for k, v in rv.items():
            metrics[k] += v
        n += 1
    return {k: v / n for k, v in metrics.items()}
