from .gdn import GDN

__all__ = ["GDN"]
