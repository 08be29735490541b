from .gdn import GDN
from .layers import QReLU

__all__ = ["GDN", "QReLU"]
