"""Derived device data cached on modules: packed weights (transforms), packed CDF tables and host copies of
one-element buffers (entropy models).

Validity.  An entry is keyed on ``(data_ptr, _version, device)`` of the tensors it was derived from, which catches
optimizer steps, ``copy_``, ``load_state_dict`` into existing storage and re-assignment.  Edits made through
``tensor.data`` bypass the version counter, so every entry is ALSO dropped by the events after which such edits are
expected to be visible: ``module.train()`` / ``.eval()``, ``module._apply`` (``.to()``, ``.cuda()``, ``.float()``),
``load_state_dict()`` and the model-level ``update()``.  ``invalidate_caches(model)`` does the same on demand.

Publication.  An entry built on stream A is consumed by launches on other streams (the codec pipeline runs
analysis, hyper-synthesis, synthesis and coder streams, and several request threads).  Each entry carries a CUDA
event recorded after its producing work; a consumer on a different stream waits on that event until the event has
been observed complete once (after which the data is visible to all later work on any stream).  Fills are
serialised by a lock so two request threads never build the same entry twice.
"""
from __future__ import annotations

import threading

import torch

_LOCK = threading.RLock()


class _Entry:
    __slots__ = ("key", "value", "event", "stream", "visible")

    def __init__(self, key, value, event, stream):
        self.key, self.value, self.event, self.stream = key, value, event, stream
        self.visible = event is None


def tensor_key(*tensors):
    return tuple((t.data_ptr(), t._version, tuple(t.shape), str(t.device)) for t in tensors)


def cached(owner, slot: str, key, build, device=None, synchronous: bool = False):
    """Return ``owner``'s entry ``slot`` for ``key``, building it with ``build()`` (under the fill lock) if it is
    missing or stale.  ``synchronous``: ``build`` synchronises its stream itself, no event is needed."""
    store = owner.__dict__.setdefault("_cai_cache", {})
    ent = store.get(slot)
    if ent is None or ent.key != key:
        with _LOCK:
            ent = store.get(slot)
            if ent is None or ent.key != key:
                value = build()
                if synchronous or device is None or not torch.cuda.is_available():
                    ent = _Entry(key, value, None, None)
                else:
                    cur = torch.cuda.current_stream(device)
                    ent = _Entry(key, value, cur.record_event(), cur.cuda_stream)
                store[slot] = ent
    if not ent.visible:
        cur = torch.cuda.current_stream(device)
        if ent.stream != cur.cuda_stream:
            if ent.event.query():
                ent.visible = True
            else:
                cur.wait_event(ent.event)
    return ent.value


def drop(owner):
    store = owner.__dict__.get("_cai_cache")
    if store:
        with _LOCK:
            store.clear()


class CacheOwner:
    """Mixin for ``nn.Module`` subclasses holding ``cached`` entries: drops them on train() / _apply() /
    load_state_dict()."""

    def _drop_caches(self):
        drop(self)

    def train(self, mode: bool = True):
        drop(self)
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        drop(self)
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        drop(self)
        return super()._load_from_state_dict(*args, **kwargs)

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_cai_cache", None)
        return state


def invalidate_caches(module) -> None:
    """Drop every cached derived tensor below ``module`` (call after editing parameters through ``.data``)."""
    for m in module.modules():
        drop(m)
