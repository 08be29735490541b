"""B200-native (sm_100a) implementation of CompressAI's learned-codec hot path.

Mirrors the reference package's public surface for that path only (compressai/__init__.py:37-77):
the entropy-coder registry, ``ans`` / ``_CXX`` native modules, ``entropy_models``, ``layers``, ``ops``,
``models`` and the ``zoo`` configuration table.  Everything heavy runs in ``lib/libcai_b200.so``
(C ABI in include/cai_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

_entropy_coder = "ans"
_available_entropy_coders = [_entropy_coder]


def set_entropy_coder(entropy_coder):
    """Specifies the default entropy coder used to encode the bit-streams (compressai/__init__.py:48-63)."""
    global _entropy_coder
    if entropy_coder not in _available_entropy_coders:
        raise ValueError(
            f'Invalid entropy coder "{entropy_coder}", choose from' f'({", ".join(_available_entropy_coders)}).'
        )
    _entropy_coder = entropy_coder


def get_entropy_coder():
    """Return the name of the default entropy coder used to encode the bit-streams."""
    return _entropy_coder


def available_entropy_coders():
    """Return the list of available entropy coders."""
    return _available_entropy_coders


def invalidate_caches(module):
    """Drop the packed weights / tables cached below ``module`` (needed only after editing parameters through
    ``tensor.data``, which bypasses the version counters the caches are keyed on)."""
    from ._cache import invalidate_caches as _inv

    _inv(module)
