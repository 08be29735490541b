"""Tensor-level front end of the batched rANS coder (``cai_rans_*`` / ``cai_table_*`` in the C ABI).

Host mirror of what ``EntropyModel.compress`` / ``decompress`` do around the coder in the reference
(compressai/entropy_models/entropy_models.py:259-267, :313-323) -- minus the per-image Python loop and
the five ``.tolist()`` round trips: all strings of a batch are coded by one kernel launch and the byte
strings come back in a single device-to-host copy.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, current_stream, lib, ptr, require_cuda


# Optional profiling hook used by bench.py: when set to a dict, every coder launch appends a
# (start_event, end_event) pair under its kernel name, recorded on the launching stream.
TIMING = None


class _Timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if TIMING is not None:
            self.e1.record()
            TIMING.setdefault(self.name, []).append((self.e0, self.e1))
        return False


class CdfTable:
    """Packed CDF tables resident in HBM (``cai_table_t``).  Built once per ``update()``."""

    def __init__(self, quantized_cdf: torch.Tensor, cdf_length: torch.Tensor, offset: torch.Tensor):
        require_cuda(quantized_cdf, "_quantized_cdf")
        if quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {tuple(quantized_cdf.size())}")
        cdf = quantized_cdf.detach().to(torch.int32).contiguous()
        ln = cdf_length.detach().reshape(-1).to(device=cdf.device, dtype=torch.int32).contiguous()
        off = offset.detach().reshape(-1).to(device=cdf.device, dtype=torch.int32).contiguous()
        if ln.numel() != cdf.size(0) or off.numel() != cdf.size(0):
            raise ValueError("cdf, cdf_length and offset disagree on the number of rows")
        self.device = cdf.device
        self.K, self.Lmax = int(cdf.size(0)), int(cdf.size(1))
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().cai_table_create(ptr(cdf), ptr(ln), ptr(off), self.K, self.Lmax, current_stream(),
                                         ctypes.byref(self._h)), "cai_table_create")

    @property
    def handle(self):
        return self._h

    def info(self):
        K, nb, inl = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        by = ctypes.c_int64()
        check(lib().cai_table_info(self._h, ctypes.byref(K), ctypes.byref(by), ctypes.byref(nb), ctypes.byref(inl)))
        return {"K": K.value, "blob_bytes": by.value, "lut_buckets": nb.value, "in_smem": bool(inl.value)}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().cai_table_destroy(h)
            except Exception:
                pass


def slot_words(n: int) -> int:
    return int(lib().cai_rans_slot_words(int(n)))


def wait_stream(stream=None) -> None:
    """Block the calling host thread until ``stream`` (default: the current one) has drained, WITHOUT spinning.
    ``stream.synchronize()`` busy-polls by default: with several request threads per process and several processes
    per box (one per GPU) the pollers eat the cores the launching threads need (measured: ~0.5 s of CPU time per
    122 ms request).  A blocking-sync event puts the thread to sleep until the GPU signals."""
    stream = torch.cuda.current_stream() if stream is None else stream
    ev = torch.cuda.Event(blocking=True)
    ev.record(stream)
    ev.synchronize()


def to_host(t: torch.Tensor) -> torch.Tensor:
    """Device -> host read that only blocks THIS thread.  ``tensor.cpu()`` copies into pageable memory: the driver
    runs that copy synchronously and other host threads' CUDA calls queue behind it until the producing kernels have
    finished (measured in the serving loop: every other request stalled for the length of an encode).  A pinned
    destination + stream synchronise has no such side effect."""
    if not t.is_cuda:
        return t
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    wait_stream(torch.cuda.current_stream(t.device))
    return host


class PackedStrings(list):
    """The byte strings of a batch as zero-copy views of ONE pinned host buffer.

    A ``list`` whose items are read-only ``memoryview`` slices (``len()``, ``==`` against ``bytes``, ``file.write``,
    ``b"".join`` and ``bytes(s)`` all work as with the ``bytes`` objects the reference returns); the strings sit
    back to back, word aligned, in ``host`` (int32, pinned) at word offsets ``begin[i] .. begin[i+1]``.  Building it
    costs no per-string copy, and ``decode`` uploads an intact ``PackedStrings`` (or a slice of one) with a single
    asynchronous copy straight from the pinned buffer -- no repacking on the host."""

    def __init__(self, host: torch.Tensor, begin: np.ndarray):
        self.host, self.begin = host, np.asarray(begin, dtype=np.int64)
        raw = memoryview(host.numpy().view(np.uint8)).toreadonly()
        super().__init__(raw[int(a) * 4:int(e) * 4] for a, e in zip(self.begin[:-1], self.begin[1:]))

    def __getitem__(self, key):
        if isinstance(key, slice):
            a, e, step = key.indices(len(self))
            if step == 1 and self.intact():
                e = max(a, e)
                return PackedStrings(self.host, self.begin[a:e + 1])
            return list(self)[key]
        return super().__getitem__(key)

    def intact(self) -> bool:
        """False once the list was edited so that it no longer mirrors ``host`` / ``begin``."""
        n = len(self)
        if n != len(self.begin) - 1:
            return False
        for i in ((0, n - 1) if n else ()):
            v = super().__getitem__(i)
            if not isinstance(v, memoryview) or v.nbytes != int(self.begin[i + 1] - self.begin[i]) * 4:
                return False
        return True

    def to_bytes(self) -> List[bytes]:
        return [bytes(v) for v in self]


def as_strings(strings):
    """A sequence of byte strings as something ``decode`` indexes cheaply (PackedStrings kept as they are)."""
    return strings if isinstance(strings, PackedStrings) else list(strings)


def _raise_status(status: torch.Tensor, what: str):
    """Every non-zero per-string status is an error -- including 6 (the decoder consumed more words than the string
    holds): a valid stream is consumed exactly, so running past the end means truncation or corruption."""
    st = to_host(status)
    bad = torch.nonzero(st != 0).reshape(-1)
    if bad.numel():
        i = int(bad[0])
        raise ValueError(f"{what}: string {i}: {_lib.STATUS_TEXT.get(int(st[i]), int(st[i]))}")


class EncodedBatch:
    """Device-side result of one encode launch (slots + lengths); ``to_bytes()`` brings strings to host."""

    def __init__(self, slots, n_words, status, slot_w):
        self.slots, self.n_words, self.status, self.slot_w = slots, n_words, status, slot_w

    def device_words(self):
        """(words, word_begin) usable as ``decode(..., device_words=...)`` without leaving the device: the
        strings are read in place from the slots (begin[b] = end of slot b - n_words[b])."""
        B = self.n_words.numel()
        begin = torch.arange(1, B + 1, device=self.slots.device, dtype=torch.int64) * self.slot_w
        begin -= self.n_words.to(torch.int64)
        return self.slots.reshape(-1), begin, self.n_words

    def to_bytes(self) -> List[bytes]:
        B = self.n_words.numel()
        if B == 0:
            return []
        dev = self.slots.device
        nw = to_host(self.n_words)  # sync point: sizes are needed to allocate the packed buffer
        _raise_status(self.status, "rANS encode")
        total = int(nw.sum())
        begin = torch.empty(B + 1, dtype=torch.int64, device=dev)
        packed = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib().cai_rans_compact(ptr(self.slots), self.slot_w, ptr(self.n_words), B, ptr(begin), ptr(packed),
                                         total, current_stream()), "cai_rans_compact")
        host = torch.empty(max(total, 1), dtype=torch.int32, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        wait_stream(torch.cuda.current_stream(dev))
        raw = host.numpy().view(np.uint8)
        ends = np.cumsum(nw.numpy().astype(np.int64)) * 4
        out, a = [], 0
        for e in ends:
            out.append(raw[a:e].tobytes())
            a = int(e)
        return out


def batches_to_host(batches: Sequence[EncodedBatch]):
    """Bring the strings of several EncodedBatches to the host with TWO host synchronisations in total: all sizes
    come back in one copy, every batch is compacted into its own range of one packed device buffer, and one
    device->host copy lands all strings in one pinned buffer.  Batches may live on different streams
    (``batch.stream``).  Returns ``(host int32 pinned tensor, [word-offset array (n_b + 1) per batch])``."""
    batches = list(batches)
    live = [b for b in batches if b.n_words.numel()]
    if not live:
        return torch.empty(0, dtype=torch.int32), [np.zeros(1, np.int64) for _ in batches]
    dev = live[0].slots.device
    cur = torch.cuda.current_stream(dev)
    for b in live:
        st = getattr(b, "stream", None)
        if st is not None and st != cur:
            cur.wait_event(st.record_event())
            for t in (b.slots, b.n_words, b.status):
                t.record_stream(cur)
    counts = [int(b.n_words.numel()) for b in live]
    both = to_host(torch.cat([torch.cat([b.n_words for b in live]), torch.cat([b.status for b in live])]))  # sync 1
    nw_all, st_all = both[:sum(counts)].numpy().astype(np.int64), both[sum(counts):]
    if int(st_all.abs().max()) != 0:
        for b in live:
            _raise_status(b.status, "rANS encode")
    totals, a = [], 0
    for c in counts:
        totals.append(int(nw_all[a:a + c].sum()))
        a += c
    grand = sum(totals)
    packed = torch.empty(max(grand, 1), dtype=torch.int32, device=dev)
    begin = torch.empty(sum(counts) + len(live), dtype=torch.int64, device=dev)
    off = boff = 0
    with torch.cuda.device(dev):
        for b, c, tot in zip(live, counts, totals):
            check(lib().cai_rans_compact(ptr(b.slots), b.slot_w, ptr(b.n_words), c, ptr(begin[boff:boff + c + 1]),
                                         ptr(packed[off:]) if tot else ptr(packed), tot, current_stream()),
                  "cai_rans_compact")
            off += tot
            boff += c + 1
    host = torch.empty(max(grand, 1), dtype=torch.int32, pin_memory=True)
    host.copy_(packed, non_blocking=True)
    wait_stream(cur)                                                                                       # sync 2
    ends = np.concatenate([[0], np.cumsum(nw_all)])
    begins, a = [], 0
    it = iter(counts)
    for b in batches:
        if b.n_words.numel():
            c = next(it)
            begins.append(ends[a:a + c + 1].copy())
            a += c
        else:
            begins.append(ends[a:a + 1].copy())
    return host, begins


def batches_to_bytes(batches: Sequence[EncodedBatch]) -> List[List[bytes]]:
    """``to_bytes()`` for several EncodedBatches (see :func:`batches_to_host`), as real ``bytes`` objects."""
    host, begins = batches_to_host(batches)
    return [PackedStrings(host, bg).to_bytes() for bg in begins]


def encode(table: CdfTable, symbols: torch.Tensor, indexes: torch.Tensor) -> EncodedBatch:
    """Encode B equal-length strings.  ``symbols`` / ``indexes``: int32 [B, n] in coder order."""
    require_cuda(symbols, "symbols")
    require_cuda(indexes, "indexes")
    assert symbols.dtype == torch.int32 and indexes.dtype == torch.int32
    assert symbols.is_contiguous() and indexes.is_contiguous() and symbols.shape == indexes.shape
    B = int(symbols.size(0)) if symbols.dim() > 0 else 0
    n = int(symbols.numel() // B) if B else 0
    dev = symbols.device
    sw = slot_words(n)
    slots = torch.empty((max(B, 1), sw), dtype=torch.int32, device=dev)
    n_words = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _Timed("rans_encode_kernel"):
        check(lib().cai_rans_encode_batch(table.handle, ptr(symbols), ptr(indexes), None, n, B, ptr(slots), sw,
                                          ptr(n_words), ptr(status), current_stream()), "cai_rans_encode_batch")
    return EncodedBatch(slots, n_words, status, sw)


def strings_to_device(strings: Sequence[bytes], device) -> tuple:
    """Concatenate byte strings (zero padded to whole words) -> (uint32-as-int32 words, int64 begins).
    An intact :class:`PackedStrings` is uploaded as it is: one asynchronous copy from its pinned buffer."""
    if isinstance(strings, PackedStrings) and strings.intact() and len(strings):
        lo, hi = int(strings.begin[0]), int(strings.begin[-1])
        words = strings.host[lo:max(hi, lo + 1)].to(device, non_blocking=True)
        wb_host = torch.empty(len(strings.begin), dtype=torch.int64, pin_memory=torch.cuda.is_available())
        wb_host.numpy()[:] = strings.begin - lo
        return words, wb_host.to(device, non_blocking=True), (strings.host, wb_host)
    lens = np.array([(len(s) + 3) // 4 for s in strings], dtype=np.int64)
    begin = np.zeros(len(strings) + 1, dtype=np.int64)
    np.cumsum(lens, out=begin[1:])
    total = int(begin[-1])
    host = torch.empty(max(total, 1) * 4, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    hv = host.numpy()
    for s, a, nwords in zip(strings, begin[:-1], lens):
        n = len(s)
        hv[a * 4:a * 4 + n] = np.frombuffer(s, dtype=np.uint8)
        if n != nwords * 4:  # zero the padding of a string that is not a whole number of words
            hv[a * 4 + n:(a + nwords) * 4] = 0
    if total == 0:
        hv[:] = 0
    words = host.view(torch.int32).to(device, non_blocking=True)
    wb_host = torch.empty(len(begin), dtype=torch.int64, pin_memory=torch.cuda.is_available())
    wb_host.numpy()[:] = begin
    return words, wb_host.to(device, non_blocking=True), (host, wb_host)


def decode(table: CdfTable, strings: Sequence[bytes], indexes: torch.Tensor, state: Optional[torch.Tensor] = None,
           resume: bool = False, device_words=None, status_out: Optional[list] = None) -> torch.Tensor:
    """Decode B strings; ``indexes`` int32 [B, n] in coder order.  Returns int32 [B, n].
    With ``status_out`` (a list) the per-string status tensor is appended instead of being checked here, so the call
    stays asynchronous (the pinned staging buffer is kept alive by torch's caching host allocator)."""
    require_cuda(indexes, "indexes")
    assert indexes.dtype == torch.int32 and indexes.is_contiguous()
    B = int(indexes.size(0)) if indexes.dim() > 0 else 0
    n = int(indexes.numel() // B) if B else 0
    dev = indexes.device
    out = torch.empty_like(indexes)
    if B == 0:
        return out
    if device_words is None:
        if len(strings) != B:
            raise ValueError("Invalid strings or indexes parameters")
        words, wb, keep = strings_to_device(strings, dev)
        wcount = None
    else:
        words, wb = device_words[0], device_words[1]
        wcount = device_words[2] if len(device_words) > 2 else None
        keep = None
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _Timed("rans_decode_kernel"):
        check(lib().cai_rans_decode_batch(table.handle, ptr(words), ptr(wb), ptr(wcount), ptr(indexes), None, n, B, ptr(out),
                                          ptr(state), 1 if resume else 0, ptr(status), current_stream()),
              "cai_rans_decode_batch")
    if status_out is not None:
        status_out.append(status)
    else:  # no deferred check requested: surface decoder errors here (host strings and device words alike)
        _raise_status(status, "rANS decode")
    return out


def check_status(statuses, what="rANS decode"):
    """Deferred check of the status tensors collected with ``status_out`` (synchronises)."""
    statuses = [st for st in statuses if st.numel()]
    if not statuses:
        return
    allst = to_host(torch.cat([st.reshape(-1) for st in statuses]))
    if int(allst.abs().max()) == 0:
        return
    for st in statuses:
        _raise_status(st, what)
