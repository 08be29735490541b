"""Golden vectors for the on-disk container of the reference codec script.

Imports /root/reference/examples/codec.py (with the reference package from oracle/_ref on the path) and records what
ITS functions write: headers, a complete image bitstream for synthetic strings, and centre pad / crop geometry.
Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden_container.py"""
import importlib.util, io, os, sys
import numpy as np, torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle

oracle.import_ref()
spec = importlib.util.spec_from_file_location("ref_codec", "/root/reference/examples/codec.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
combos = [("bmshj2018-factorized", "mse", 1), ("bmshj2018-hyperprior", "mse", 4), ("mbt2018-mean", "ms-ssim", 8),
          ("bmshj2018-hyperprior", "ms-ssim", 6)]
hdrs = []
for name, metric, q in combos:
    h = ref.get_header(name, metric, q, 0, ref.CodecType.IMAGE_CODEC)
    buf = io.BytesIO(); ref.write_uchars(buf, h); hdrs.append(np.frombuffer(buf.getvalue(), dtype=np.uint8))
    assert ref.parse_header(h) == (name, metric, q)
out["header_bytes"] = np.stack(hdrs)
out["header_names"] = np.array([c[0] for c in combos]); out["header_metrics"] = np.array([c[1] for c in combos])
out["header_quality"] = np.array([c[2] for c in combos])

rng = np.random.default_rng(7)
strings = [[rng.integers(0, 256, 37, dtype=np.uint8).tobytes()], [rng.integers(0, 256, 8, dtype=np.uint8).tobytes()]]
buf = io.BytesIO()
ref.write_uchars(buf, ref.get_header("bmshj2018-hyperprior", "mse", 4, 0, ref.CodecType.IMAGE_CODEC))
ref.write_uints(buf, (501, 763)); ref.write_uchars(buf, (8,)); ref.write_body(buf, (8, 12), strings)
out["stream"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
out["stream_y"] = np.frombuffer(strings[0][0], dtype=np.uint8); out["stream_z"] = np.frombuffer(strings[1][0], dtype=np.uint8)
# read back with the reference reader
rd = io.BytesIO(buf.getvalue()); rd.read(2 + 8 + 1)
s2, shp = ref.read_body(rd)
assert s2 == strings and tuple(shp) == (8, 12)

sizes = [(501, 763), (64, 64), (1, 1), (2160, 3840), (65, 127)]
geo = []
for h, w in sizes:
    x = torch.arange(h * w, dtype=torch.float32).reshape(1, 1, h, w) + 1.0
    p = ref.pad(x, 64)
    ys, xs = torch.nonzero(p[0, 0] == 1.0)[0].tolist()          # where element (0, 0) landed
    c = ref.crop(p, (h, w))
    assert torch.equal(c, x)
    geo.append([h, w, p.size(2), p.size(3), ys, xs])
out["pad_geometry"] = np.array(geo, dtype=np.int64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "container.npz"), **out)
print("wrote container.npz", {k: v.shape for k, v in out.items()})
