"""Architecture table of compressai/zoo/image.py (:52-59 name -> class, :189-219 quality -> (N, M)) for the
model families on the hot path.  Pretrained weights are NOT fetched (no network): ``pretrained=True`` raises."""
from ..models import (Cheng2020Anchor, Cheng2020Attention, FactorizedPrior, JointAutoregressiveHierarchicalPriors,
                      MeanScaleHyperprior, ScaleHyperprior)

model_architectures = {
    "bmshj2018-factorized": FactorizedPrior,
    "bmshj2018-hyperprior": ScaleHyperprior,
    "mbt2018-mean": MeanScaleHyperprior,
    "mbt2018": JointAutoregressiveHierarchicalPriors,
    "cheng2020-anchor": Cheng2020Anchor,
    "cheng2020-attn": Cheng2020Attention,
}

cfgs = {
    "bmshj2018-factorized": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "bmshj2018-hyperprior": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "mbt2018-mean": {q: (128, 192) if q <= 4 else (192, 320) for q in range(1, 9)},
    "mbt2018": {q: (192, 192) if q <= 4 else (192, 320) for q in range(1, 9)},
    "cheng2020-anchor": {q: (128,) if q <= 3 else (192,) for q in range(1, 7)},
    "cheng2020-attn": {q: (128,) if q <= 3 else (192,) for q in range(1, 7)},
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


def mbt2018(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    return _load_model("mbt2018", metric, quality, pretrained, progress, **kwargs)


def cheng2020_anchor(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    return _load_model("cheng2020-anchor", metric, quality, pretrained, progress, **kwargs)


def cheng2020_attn(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    return _load_model("cheng2020-attn", metric, quality, pretrained, progress, **kwargs)


models = {
    "bmshj2018-factorized": bmshj2018_factorized,
    "bmshj2018-hyperprior": bmshj2018_hyperprior,
    "mbt2018-mean": mbt2018_mean,
    "mbt2018": mbt2018,
    "cheng2020-anchor": cheng2020_anchor,
    "cheng2020-attn": cheng2020_attn,
}

# the name the reference's zoo exports (compressai/zoo/__init__.py) and its examples import
image_models = models
