from .google import Hyperprior, HyperDecoder, HyperDecoderWithQReLU, HyperEncoder  # noqa: F401

__all__ = ["Hyperprior", "HyperEncoder", "HyperDecoder", "HyperDecoderWithQReLU"]
