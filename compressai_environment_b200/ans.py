"""Drop-in for the reference's ``compressai.ans`` pybind11 module
(compressai/cpp_exts/rans/rans_interface.cpp:361-381): ``RansEncoder``, ``RansDecoder`` and
``BufferedRansEncoder`` with the same list-based signatures, argument meaning and return types.

Every call runs the sm_100a coder kernels through the C ABI (``cai_rans_*``); there is no host coder.
The list API exists for compatibility (the reference's tests, JARHP-style callers); the fast path is
``EntropyModel.compress / decompress`` which hands whole device tensors to the same kernels.

Defined here but undefined in the reference: streams of 0 or 1 symbols (the reference under-runs its
output buffer) -- N = 0 yields the 8 flush bytes of the initial state.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import List

import numpy as np
import torch

from . import coder
from ._lib import CaiError

__name__ = "compressai.ans"  # noqa: A001  (rans_interface.cpp:362 sets the same attribute)

_TABLE_CACHE: "OrderedDict[tuple, coder.CdfTable]" = OrderedDict()
_TABLE_CACHE_SIZE = 8


def _device():
    if not torch.cuda.is_available():
        raise CaiError("compressai.ans needs a CUDA device: the coder runs on the GPU and has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_table_arrays(cdfs, cdfs_sizes, offsets):
    sizes = np.asarray(cdfs_sizes, dtype=np.int32).reshape(-1)
    offs = np.asarray(offsets, dtype=np.int32).reshape(-1)
    K = len(cdfs)
    if sizes.size != K or offs.size != K:
        raise ValueError("cdfs, cdfs_sizes and offsets must have the same number of rows")
    try:
        arr = np.asarray(cdfs, dtype=np.int32)
        if arr.ndim != 2:
            raise ValueError
    except ValueError:  # ragged rows
        L = max((len(r) for r in cdfs), default=0)
        arr = np.zeros((K, L), dtype=np.int32)
        for i, r in enumerate(cdfs):
            arr[i, :len(r)] = r
    if arr.shape[1] < 2:
        arr = np.pad(arr, ((0, 0), (0, 2 - arr.shape[1])))
    return np.ascontiguousarray(arr), sizes, offs


def _table_for(arr, sizes, offs) -> coder.CdfTable:
    dev = _device()
    h = hashlib.blake2b(digest_size=16)
    h.update(arr.tobytes())
    h.update(sizes.tobytes())
    h.update(offs.tobytes())
    key = (dev.index, arr.shape, h.digest())
    t = _TABLE_CACHE.get(key)
    if t is None:
        t = coder.CdfTable(torch.from_numpy(arr).to(dev), torch.from_numpy(sizes).to(dev),
                           torch.from_numpy(offs).to(dev))
        _TABLE_CACHE[key] = t
        while len(_TABLE_CACHE) > _TABLE_CACHE_SIZE:
            _TABLE_CACHE.popitem(last=False)
    else:
        _TABLE_CACHE.move_to_end(key)
    return t


def _i32(x, dev):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(1, -1))).to(dev)


class BufferedRansEncoder:
    """rans_interface.hpp:49-67: accumulate symbols over several calls, emit one string on flush()."""

    def __init__(self):
        self._calls = []  # (symbols i32, indexes i32, (cdf, sizes, offsets))

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> None:
        sym = np.asarray(symbols, dtype=np.int32).reshape(-1)
        idx = np.asarray(indexes, dtype=np.int32).reshape(-1)
        if sym.size != idx.size:
            raise ValueError("symbols and indexes must have the same length")
        self._calls.append((sym, idx, _as_table_arrays(cdfs, cdfs_sizes, offsets)))

    def flush(self) -> bytes:
        calls, self._calls = self._calls, []
        dev = _device()
        if not calls:
            arr, sizes, offs = np.array([[0, 65536]], np.int32), np.array([2], np.int32), np.array([0], np.int32)
            sym = idx = np.zeros(0, np.int32)
        else:
            arr, sizes, offs = calls[0][2]
            same = all(c[2][0].shape == arr.shape and np.array_equal(c[2][0], arr) and np.array_equal(c[2][1], sizes)
                       and np.array_equal(c[2][2], offs) for c in calls[1:])
            if same:
                idx = np.concatenate([c[1] for c in calls])
            else:  # different tables per call: stack the rows and shift each call's indexes
                L = max(c[2][0].shape[1] for c in calls)
                rows, szs, ofs, idxs, base = [], [], [], [], 0
                for _, ix, (a, s, o) in calls:
                    rows.append(np.pad(a, ((0, 0), (0, L - a.shape[1]))))
                    szs.append(s)
                    ofs.append(o)
                    idxs.append(ix + base)
                    base += a.shape[0]
                arr, sizes, offs = np.concatenate(rows), np.concatenate(szs), np.concatenate(ofs)
                idx = np.concatenate(idxs).astype(np.int32)
            sym = np.concatenate([c[0] for c in calls])
        table = _table_for(arr, sizes, offs)
        enc = coder.encode(table, _i32(sym, dev), _i32(idx, dev))
        return enc.to_bytes()[0]


class RansEncoder:
    """rans_interface.hpp:69-83 / .cpp:202-213."""

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
        b = BufferedRansEncoder()
        b.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return b.flush()


class RansDecoder:
    """rans_interface.hpp:85-113 / .cpp:215-359.  Stateful for set_stream / decode_stream."""

    def __init__(self):
        self._words = None
        self._state = None

    def decode_with_indexes(self, encoded: bytes, indexes, cdfs, cdfs_sizes, offsets) -> List[int]:
        dev = _device()
        table = _table_for(*_as_table_arrays(cdfs, cdfs_sizes, offsets))
        out = coder.decode(table, [bytes(encoded)], _i32(indexes, dev))
        return out.reshape(-1).cpu().tolist()

    def set_stream(self, encoded: bytes) -> None:
        dev = _device()
        words, wb, keep = coder.strings_to_device([bytes(encoded)], dev)
        torch.cuda.current_stream(dev).synchronize()
        self._words = (words, wb)
        # state = (x, next word) as two uint64; initialised by the first decode_stream call
        self._state = None

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets) -> List[int]:
        if self._words is None:
            raise RuntimeError("decode_stream called before set_stream")
        dev = self._words[0].device
        table = _table_for(*_as_table_arrays(cdfs, cdfs_sizes, offsets))
        resume = self._state is not None
        if not resume:
            self._state = torch.zeros(2, dtype=torch.int64, device=dev)
        out = coder.decode(table, None, _i32(indexes, dev), state=self._state, resume=resume,
                           device_words=self._words)
        return out.reshape(-1).cpu().tolist()
