"""Container format against vectors written by the reference's own examples/codec.py
(tests/golden/make_golden_container.py).  CPU only: the format is host-side byte handling."""
import io

import numpy as np
import pytest
import torch


def test_headers_match_reference(golden):
    from compressai_environment_b200 import codec_io

    g = golden("container")
    for hb, name, metric, q in zip(g["header_bytes"], g["header_names"], g["header_metrics"], g["header_quality"]):
        assert codec_io.pack_header(str(name), str(metric), int(q)) == hb.tobytes()
        assert codec_io.unpack_header(hb.tobytes()) == (str(name), str(metric), int(q))
    with pytest.raises(ValueError):
        codec_io.pack_header("no-such-model", "mse", 1)
    with pytest.raises(ValueError):
        codec_io.unpack_header(bytes([200, 0]))


def test_image_stream_is_byte_identical_and_round_trips(golden):
    from compressai_environment_b200 import codec_io

    g = golden("container")
    strings = [[g["stream_y"].tobytes()], [g["stream_z"].tobytes()]]
    buf = io.BytesIO()
    n = codec_io.write_image(buf, "bmshj2018-hyperprior", "mse", 4, (501, 763), (8, 12), strings, bitdepth=8)
    assert buf.getvalue() == g["stream"].tobytes() and n == len(g["stream"])
    info = codec_io.read_image(io.BytesIO(g["stream"].tobytes()))   # a file the reference wrote
    assert info["model"] == "bmshj2018-hyperprior" and info["metric"] == "mse" and info["quality"] == 4
    assert info["original_size"] == (501, 763) and info["bitdepth"] == 8 and info["shape"] == (8, 12)
    assert info["strings"] == strings
    with pytest.raises(ValueError):
        codec_io.read_image(io.BytesIO(g["stream"].tobytes()[:-3]))


def test_pad_crop_geometry_matches_reference(golden):
    from compressai_environment_b200 import codec_io

    for h, w, H, W, y0, x0 in golden("container")["pad_geometry"].tolist():
        x = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w) + 1.0
        p = codec_io.pad(x, 64)
        assert tuple(p.shape[2:]) == (H, W)
        assert p[0, 0, y0, x0] == 1.0 and int(torch.count_nonzero(p)) == h * w
        assert torch.equal(codec_io.crop(p, (h, w)), x)


def test_ms_ssim_restatement_sanity():
    from compressai_environment_b200.utils.eval_model import ms_ssim, psnr

    torch.manual_seed(0)
    x = torch.rand(2, 3, 176, 200)
    assert abs(float(ms_ssim(x, x)) - 1.0) < 1e-6
    a = float(ms_ssim(x, (x + 0.05 * torch.randn_like(x)).clamp(0, 1)))
    b = float(ms_ssim(x, (x + 0.20 * torch.randn_like(x)).clamp(0, 1)))
    assert 0.0 < b < a < 1.0
    assert abs(psnr(x, x + 0.1) - 20.0) < 1e-3
    with pytest.raises(ValueError):
        ms_ssim(x[..., :100, :100], x[..., :100, :100])


def test_container_round_trip_random_strings():
    """Property: any (model, metric, quality, size, shape, strings) survives write_image -> read_image."""
    import random

    from compressai_environment_b200 import codec_io

    rnd = random.Random(11)
    for _ in range(50):
        model = rnd.choice(sorted(codec_io.MODEL_IDS))
        metric = rnd.choice(sorted(codec_io.METRIC_IDS))
        q = rnd.randint(1, 8)
        size = (rnd.randint(1, 5000), rnd.randint(1, 5000))
        shape = (rnd.randint(1, 300), rnd.randint(1, 300))
        strings = [[bytes(rnd.getrandbits(8) for _ in range(rnd.choice([0, 8, 12, 4096])))] for _ in range(rnd.randint(1, 3))]
        buf = io.BytesIO()
        n = codec_io.write_image(buf, model, metric, q, size, shape, strings, bitdepth=rnd.choice([8, 10]))
        assert n == len(buf.getvalue()) == 2 + 8 + 1 + 12 + sum(4 + len(s[0]) for s in strings)
        info = codec_io.read_image(io.BytesIO(buf.getvalue()))
        assert (info["model"], info["metric"], info["quality"]) == (model, metric, q)
        assert info["original_size"] == size and info["shape"] == shape and info["strings"] == strings
