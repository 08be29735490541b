"""Architecture table of compressai/zoo/image.py (:52-59 name -> class, :189-219 quality -> (N, M)) for the
model families on the hot path.  Pretrained weights are NOT fetched (no network): ``pretrained=True`` raises."""
from ..models import FactorizedPrior, MeanScaleHyperprior, ScaleHyperprior

model_architectures = {
    "bmshj2018-factorized": FactorizedPrior,
    "bmshj2018-hyperprior": ScaleHyperprior,
    "mbt2018-mean": MeanScaleHyperprior,
}

cfgs = {
    "bmshj2018-factorized": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "bmshj2018-hyperprior": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "mbt2018-mean": {q: (128, 192) if q <= 4 else (192, 320) for q in range(1, 9)},
}


def _load_model(architecture, metric, quality, pretrained=False, progress=True, **kwargs):
    if architecture not in model_architectures:
        raise ValueError(f'Invalid architecture name "{architecture}"')
    if quality not in cfgs[architecture]:
        raise ValueError(f'Invalid quality value "{quality}"')
    if pretrained:
        raise RuntimeError("pretrained weights are not available offline; load a reference state_dict instead")
    return model_architectures[architecture](*cfgs[architecture][quality], **kwargs)


def bmshj2018_factorized(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    return _load_model("bmshj2018-factorized", metric, quality, pretrained, progress, **kwargs)


def bmshj2018_hyperprior(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    return _load_model("bmshj2018-hyperprior", metric, quality, pretrained, progress, **kwargs)


def mbt2018_mean(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    return _load_model("mbt2018-mean", metric, quality, pretrained, progress, **kwargs)


models = {
    "bmshj2018-factorized": bmshj2018_factorized,
    "bmshj2018-hyperprior": bmshj2018_hyperprior,
    "mbt2018-mean": mbt2018_mean,
}

# the name the reference's zoo exports (compressai/zoo/__init__.py) and its examples import
image_models = models
