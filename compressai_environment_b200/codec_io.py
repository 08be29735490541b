"""On-disk / wire container of the reference codec script -- SURVEY.md 8(f) rank 1, the data format either side of
the hot path.  What it mirrors (reference tree, ``examples/codec.py``):

  :60-62, :154-185  header     1 byte model id (position of the name in ``compressai.zoo.models``),
                               1 byte ``(metric_id << 4) | (quality - 1)``
  :227-254          pad / crop centre padding of H and W to a multiple of 64, zeros; crop undoes it
  :272-307          image file  header | >2I original (h, w) | B bit depth | body
  :187-205          body        >2I latent shape | >I number of latents | per latent: >I length, bytes of image 0

All integers are big-endian.  A file written by the reference is readable here and vice versa (golden vectors:
``tests/golden/container.npz``, produced by the reference's own functions).  Only the image container is covered;
the video container (``:310-345``) belongs to ssf2020, which is outside the hot path.
"""
import struct
from typing import BinaryIO, Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

# name -> id = position in the reference registry compressai/zoo/__init__.py:41-57 (image models, then video models)
MODEL_IDS = {"bmshj2018-factorized": 0, "bmshj2018-hyperprior": 1, "mbt2018-mean": 2, "mbt2018": 3,
             "cheng2020-anchor": 4, "cheng2020-attn": 5, "ssf2020": 6}
METRIC_IDS = {"mse": 0, "ms-ssim": 1}
_MODEL_NAMES = {v: k for k, v in MODEL_IDS.items()}
_METRIC_NAMES = {v: k for k, v in METRIC_IDS.items()}


def pack_header(model: str, metric: str, quality: int) -> bytes:
    if model not in MODEL_IDS:
        raise ValueError(f'Invalid architecture name "{model}"')
    if metric not in METRIC_IDS:
        raise ValueError(f'Invalid metric "{metric}"')
    if not 1 <= int(quality) <= 16:
        raise ValueError(f'Invalid quality value "{quality}"')
    return struct.pack(">2B", MODEL_IDS[model], (METRIC_IDS[metric] << 4) | ((int(quality) - 1) & 0x0F))


def unpack_header(two_bytes: bytes) -> Tuple[str, str, int]:
    model_id, code = struct.unpack(">2B", two_bytes)
    if model_id not in _MODEL_NAMES or (code >> 4) not in _METRIC_NAMES:
        raise ValueError("not a CompressAI bitstream header")
    return _MODEL_NAMES[model_id], _METRIC_NAMES[code >> 4], (code & 0x0F) + 1


def _margins(h: int, w: int, H: int, W: int):
    left, top = (W - w) // 2, (H - h) // 2
    return left, W - w - left, top, H - h - top


def pad(x: torch.Tensor, p: int = 64) -> torch.Tensor:
    """Centre-pad H and W up to the next multiple of ``p`` with zeros (works on device tensors)."""
    h, w = x.size(2), x.size(3)
    H, W = (h + p - 1) // p * p, (w + p - 1) // p * p
    return F.pad(x, _margins(h, w, H, W), mode="constant", value=0)


def crop(x: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """Inverse of :func:`pad`: cut the centre ``size = (h, w)`` window back out."""
    l, r, t, b = _margins(int(size[0]), int(size[1]), x.size(2), x.size(3))
    return F.pad(x, (-l, -r, -t, -b), mode="constant", value=0)


def write_body(fd: BinaryIO, shape: Sequence[int], strings: List[List[bytes]]) -> int:
    """``strings`` as returned by ``model.compress`` for a batch of ONE image: one list per latent."""
    n = fd.write(struct.pack(">3I", int(shape[0]), int(shape[1]), len(strings)))
    for per_latent in strings:
        s = per_latent[0]
        n += fd.write(struct.pack(">I", len(s)))
        n += fd.write(s)
    return n


def _read_exact(fd: BinaryIO, n: int) -> bytes:
    b = fd.read(n)
    if len(b) != n:
        raise ValueError("truncated CompressAI bitstream")
    return b


def read_body(fd: BinaryIO) -> Tuple[List[List[bytes]], Tuple[int, int]]:
    h, w, n = struct.unpack(">3I", _read_exact(fd, 12))
    strings = []
    for _ in range(n):
        (length,) = struct.unpack(">I", _read_exact(fd, 4))
        strings.append([_read_exact(fd, length)])
    return strings, (h, w)


def write_image(fd: BinaryIO, model: str, metric: str, quality: int, original_size: Sequence[int], shape: Sequence[int],
                strings: List[List[bytes]], bitdepth: int = 8) -> int:
    n = fd.write(pack_header(model, metric, quality))
    n += fd.write(struct.pack(">2I", int(original_size[0]), int(original_size[1])))
    n += fd.write(struct.pack(">B", int(bitdepth)))
    return n + write_body(fd, shape, strings)


def read_image(fd: BinaryIO) -> Dict:
    model, metric, quality = unpack_header(_read_exact(fd, 2))
    h, w = struct.unpack(">2I", _read_exact(fd, 8))
    (bitdepth,) = struct.unpack(">B", _read_exact(fd, 1))
    strings, shape = read_body(fd)
    return {"model": model, "metric": metric, "quality": quality, "original_size": (h, w), "bitdepth": bitdepth,
            "shape": shape, "strings": strings}


@torch.no_grad()
def encode_image(net, x: torch.Tensor, fd: BinaryIO, model: str, metric: str = "mse", quality: int = 1) -> Dict:
    """Reference ``encode_image`` (:272-307) for an image tensor [1, 3, h, w] in [0, 1]: pad, compress on the GPU,
    write the container.  Returns ``{"bpp": ...}`` like the reference."""
    if x.dim() != 4 or x.size(0) != 1:
        raise ValueError("encode_image takes one image [1, 3, h, w]")
    h, w = x.size(2), x.size(3)
    out = net.compress(pad(x, 64))
    size = write_image(fd, model, metric, quality, (h, w), out["shape"], out["strings"])
    return {"bpp": size * 8.0 / (h * w)}


@torch.no_grad()
def decode_image(net, fd: BinaryIO) -> Dict:
    """Reference ``decode_image`` (:408-425) minus the PIL conversion: returns the cropped reconstruction tensor and
    the parsed header.  ``net`` must be the model the header names (``info["model"], info["quality"]``)."""
    info = read_image(fd)
    out = net.decompress(info["strings"], info["shape"])
    info["x_hat"] = crop(out["x_hat"], info["original_size"])
    return info
