"""CPU ORACLE front end (test infrastructure, NOT product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  It wraps ``libcai_oracle.so`` (plain-C restatement of the reference's
integer arithmetic, see ``cai_oracle.c``) and restates the small floating point steps that feed the
integer path with numpy, each citing the reference lines it follows (paths relative to
``/root/reference``).  Parity status: PINNED by ``tests/test_oracle.py`` against ``tests/golden/``
(vectors produced by the compiled reference) and, when ``oracle/_ref`` exists, the reference itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    """Compile ``libcai_oracle.so`` (gcc, seconds) and, if /root/reference is mounted, oracle/_ref."""
    so = os.path.join(_HERE, "libcai_oracle.so")
    src = os.path.join(_HERE, "cai_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libcai_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref() -> bool:
    """(Re)install the reference into oracle/_ref when its sources are mounted; True if usable."""
    if os.path.isdir("/root/reference/compressai") and not have_ref():
        subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return have_ref()


def have_ref() -> bool:
    d = os.path.join(_HERE, "_ref", "compressai")
    return os.path.isdir(d) and any(f.startswith("ans.") for f in os.listdir(d))


def import_ref():
    """Import the UNMODIFIED reference package (``compressai``) from oracle/_ref."""
    if not have_ref():
        raise RuntimeError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    p = os.path.join(_HERE, "_ref")
    if p not in sys.path:
        sys.path.insert(0, p)
    import compressai  # noqa: F401  (the reference)

    return compressai


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_rans_encode.restype = ctypes.c_int64
        L.orc_rans_encode.argtypes = [_i32p, _i32p, ctypes.c_int64, _i32p, _i32p, _i32p, ctypes.c_int32,
                                      ctypes.c_int32, _u32p, ctypes.c_int64]
        L.orc_rans_decode.restype = ctypes.c_int
        L.orc_rans_decode.argtypes = [_u32p, ctypes.c_int64, _i32p, ctypes.c_int64, _i32p, _i32p, _i32p,
                                      ctypes.c_int32, ctypes.c_int32, _i32p]
        L.orc_rans_encode_batch.restype = ctypes.c_int64
        L.orc_rans_encode_batch.argtypes = [_i32p, _i32p, ctypes.c_int64, ctypes.c_int64, _i32p, _i32p,
                                            _i32p, ctypes.c_int32, ctypes.c_int32, _u32p, ctypes.c_int64,
                                            _i64p]
        L.orc_rans_decode_batch.restype = ctypes.c_int64
        L.orc_rans_decode_batch.argtypes = [_u32p, ctypes.c_int64, _i64p, _i32p, ctypes.c_int64,
                                            ctypes.c_int64, _i32p, _i32p, _i32p, ctypes.c_int32,
                                            ctypes.c_int32, _i32p]
        L.orc_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.orc_pmf_to_quantized_cdf.argtypes = [_f32p, ctypes.c_int32, ctypes.c_int32, _u32p]
        L.orc_pmf_rows_to_cdf.restype = ctypes.c_int
        L.orc_pmf_rows_to_cdf.argtypes = [_f32p, _i32p, _f32p, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int32, _i32p]
        _LIB = L
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _p(a, t):
    return a.ctypes.data_as(t)


def _tables(cdfs, cdf_len, offsets):
    cdfs = _c(cdfs, np.int32)
    assert cdfs.ndim == 2
    return cdfs, _c(cdf_len, np.int32), _c(offsets, np.int32)


def worst_case_words(n: int) -> int:
    """Upper bound on the words one string can need: 16 + 4*9 bits per symbol, + 31 bits of slack,
    + the 2 flush words."""
    return (52 * int(n) + 31 + 31) // 32 + 2


def rans_encode(symbols, indexes, cdfs, cdf_len, offsets) -> bytes:
    """rans_interface.cpp:202-213 (RansEncoder.encode_with_indexes) for one string."""
    sym, idx = _c(symbols, np.int32).ravel(), _c(indexes, np.int32).ravel()
    cdfs, cdf_len, offsets = _tables(cdfs, cdf_len, offsets)
    cap = worst_case_words(sym.size)
    out = np.empty(cap, np.uint32)
    nw = lib().orc_rans_encode(_p(sym, _i32p), _p(idx, _i32p), sym.size, _p(cdfs, _i32p),
                               _p(cdf_len, _i32p), _p(offsets, _i32p), cdfs.shape[0], cdfs.shape[1],
                               _p(out, _u32p), cap)
    if nw < 0:
        raise ValueError("oracle encode: invalid index/table or overflow")
    return out[cap - nw:].astype("<u4").tobytes()


def rans_decode(data: bytes, indexes, cdfs, cdf_len, offsets) -> np.ndarray:
    """rans_interface.cpp:215-284 (RansDecoder.decode_with_indexes) for one string."""
    idx = _c(indexes, np.int32).ravel()
    cdfs, cdf_len, offsets = _tables(cdfs, cdf_len, offsets)
    w = np.frombuffer(data, dtype="<u4").astype(np.uint32)
    out = np.empty(idx.size, np.int32)
    rc = lib().orc_rans_decode(_p(w, _u32p), w.size, _p(idx, _i32p), idx.size, _p(cdfs, _i32p),
                               _p(cdf_len, _i32p), _p(offsets, _i32p), cdfs.shape[0], cdfs.shape[1],
                               _p(out, _i32p))
    if rc != 0:
        raise ValueError(f"oracle decode failed rc={rc}")
    return out


def rans_encode_batch(symbols, indexes, cdfs, cdf_len, offsets, threads: int = 1):
    """Encode B equal-length strings; returns (list[bytes], seconds-free raw arrays).  Used by the CPU
    baseline; ``threads`` > 1 splits the strings over a thread pool (ctypes releases the GIL)."""
    sym, idx = _c(symbols, np.int32), _c(indexes, np.int32)
    B, n = sym.shape
    cdfs, cdf_len, offsets = _tables(cdfs, cdf_len, offsets)
    cap = worst_case_words(n)
    out = np.empty((B, cap), np.uint32)
    nw = np.zeros(B, np.int64)

    def work(lo, hi):
        return lib().orc_rans_encode_batch(
            _p(sym[lo:hi], _i32p), _p(idx[lo:hi], _i32p), n, hi - lo, _p(cdfs, _i32p), _p(cdf_len, _i32p),
            _p(offsets, _i32p), cdfs.shape[0], cdfs.shape[1], _p(out[lo:hi], _u32p), cap,
            _p(nw[lo:hi], _i64p))

    _fan_out(work, B, threads)
    return out, nw


def rans_decode_batch(slots, nw, indexes, cdfs, cdf_len, offsets, threads: int = 1) -> np.ndarray:
    idx = _c(indexes, np.int32)
    B, n = idx.shape
    cdfs, cdf_len, offsets = _tables(cdfs, cdf_len, offsets)
    out = np.empty((B, n), np.int32)
    cap = slots.shape[1]

    def work(lo, hi):
        return lib().orc_rans_decode_batch(
            _p(slots[lo:hi], _u32p), cap, _p(nw[lo:hi], _i64p), _p(idx[lo:hi], _i32p), n, hi - lo,
            _p(cdfs, _i32p), _p(cdf_len, _i32p), _p(offsets, _i32p), cdfs.shape[0], cdfs.shape[1],
            _p(out[lo:hi], _i32p))

    _fan_out(work, B, threads)
    return out


def _fan_out(work, B, threads):
    threads = max(1, min(int(threads), B))
    if threads == 1:
        bad = work(0, B)
    else:
        cuts = np.linspace(0, B, threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            bad = sum(ex.map(lambda t: work(int(cuts[t]), int(cuts[t + 1])), range(threads)))
    if bad:
        raise ValueError("oracle batch coder: invalid input or overflow")


def slots_to_bytes(slots, nw):
    cap = slots.shape[1]
    return [slots[b, cap - int(nw[b]):].astype("<u4").tobytes() for b in range(slots.shape[0])]


def pmf_to_quantized_cdf(pmf, precision: int = 16) -> list:
    """ops.cpp:40-109.  Raises ValueError like the pybind-translated std::domain_error."""
    p = _c(pmf, np.float32).ravel()
    out = np.empty(p.size + 1, np.uint32)
    rc = lib().orc_pmf_to_quantized_cdf(_p(p, _f32p), p.size, precision, _p(out, _u32p))
    if rc == -1:
        raise ValueError("Invalid `pmf`, non-finite or negative element found")
    if rc == -2:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    if rc != 0:
        raise ValueError("Invalid `pmf`: no symbol to steal frequency from")
    return out.astype(np.int64).tolist()


def pmf_rows_to_cdf(pmf, pmf_len, tail, precision: int = 16) -> np.ndarray:
    """entropy_models.py:204-212 (EntropyModel._pmf_to_cdf): per-row cat(pmf[:len], tail) -> cdf."""
    pmf = _c(pmf, np.float32)
    K, Lp = pmf.shape
    pl, tl = _c(pmf_len, np.int32), _c(tail, np.float32).ravel()
    out = np.zeros((K, Lp + 2), np.int32)
    rc = lib().orc_pmf_rows_to_cdf(_p(pmf, _f32p), _p(pl, _i32p), _p(tl, _f32p), K, Lp, precision,
                                   _p(out, _i32p))
    if rc != 0:
        raise ValueError(f"oracle pmf_rows_to_cdf failed rc={rc}")
    return out


# ------------------------------------------------------------------------------------------------
# float steps that feed the integer path (numpy, fp32 unless noted)
# ------------------------------------------------------------------------------------------------
F32_BOUND_SCALE = np.float32(0.11)  # LowerBound(0.11).bound as fp32 (bound_ops.py:69-71)
F32_BOUND_LIK = np.float32(1e-9)


def quantize_symbols(x, means=None) -> np.ndarray:
    """entropy_models.py:167-180: round-half-even(x - mu) -> int32."""
    v = np.asarray(x, np.float32)
    if means is not None:
        v = v - np.asarray(means, np.float32)
    return np.rint(v).astype(np.int32)


def dequantize(sym, means=None) -> np.ndarray:
    """entropy_models.py:188-197."""
    out = np.asarray(sym).astype(np.float32)
    if means is not None:
        out = out + np.asarray(means, np.float32)
    return out


def gc_build_indexes(scales, scale_table, bound=F32_BOUND_SCALE) -> np.ndarray:
    """entropy_models.py:684-689: idx = (T-1) - #{j < T-1 : max(s, bound) <= table[j]}."""
    s = np.maximum(np.asarray(scales, np.float32), np.float32(bound))
    t = np.asarray(scale_table, np.float32)
    idx = np.full(s.shape, t.size - 1, np.int32)
    for v in t[:-1]:
        idx -= (s <= v).astype(np.int32)
    return idx


def get_scale_table(lo=0.11, hi=256, levels=64) -> np.ndarray:
    """models/google.py:195-201 (torch.exp(torch.linspace(...)) in fp32)."""
    import torch

    import math
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels)).numpy()


def gdn(x, beta, gamma, inverse=False) -> np.ndarray:
    """layers/gdn.py:77-92 on NCHW fp32 with already-reparametrised beta [C], gamma [C, C] (float64
    accumulation so that it can serve as the accuracy reference for an fp32 kernel)."""
    x64 = np.asarray(x, np.float64)
    norm = np.einsum("ij,njhw->nihw", np.asarray(gamma, np.float64), x64 * x64)
    norm = norm + np.asarray(beta, np.float64)[None, :, None, None]
    norm = np.sqrt(norm) if inverse else 1.0 / np.sqrt(norm)
    return (x64 * norm).astype(np.float32)
