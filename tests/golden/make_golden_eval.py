"""Golden numbers for the eval_model protocol: the reference's own ``inference`` and
``inference_entropy_estimation`` (compressai/utils/eval_model/__main__.py) run on the seeded tiny hyperprior fixture.
``pytorch_msssim`` is not installed, so it is stubbed for the import; MS-SSIM is therefore not part of the fixture.
Run in the build container only:  python tests/golden/make_golden_eval.py"""
import importlib, os, sys, types
import numpy as np, torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle

oracle.import_ref()
stub = types.ModuleType("pytorch_msssim")
stub.ms_ssim = lambda a, b, data_range=1.0: torch.zeros(())
sys.modules["pytorch_msssim"] = stub
ref = importlib.import_module("compressai.utils.eval_model.__main__")
from compressai.models import ScaleHyperprior

g = np.load(os.path.join(ROOT, "tests", "golden", "model_hyperprior.npz"))
net = ScaleHyperprior(int(g["N"]), int(g["M"]))
net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")})
net.eval()
x = torch.from_numpy(g["x"]).float()[0, :, :61, :100].contiguous()
a = ref.inference(net, x)
b = ref.inference_entropy_estimation(net, torch.from_numpy(g["x"]).float()[0])  # forward() needs sizes the strides divide
np.savez(os.path.join(ROOT, "tests", "golden", "eval.npz"), inf_psnr=a["psnr"], inf_bpp=a["bpp"], est_psnr=b["psnr"],
         est_bpp=b["bpp"], crop=np.array([61, 100]))
print("inference", a, "\nestimation", b)
