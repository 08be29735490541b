"""ctypes binding of ``libcai_b200.so`` (C ABI declared in ``include/cai_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is
raised.  The library is built in-tree by ``__graft_entry__.build()`` /
``make -C compressai_environment_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int8, c_int32, c_int64, c_void_p

# The codec pipeline runs ~10 streams (analysis, synthesis, one coder stream per in-flight chunk); with the default
# of 8 hardware queues unrelated streams would alias and serialise.  Only effective before CUDA is initialised.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcai_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

CAI_LAYOUT_NCHW = 0
CAI_LAYOUT_NHWC = 1

STATUS_TEXT = {
    0: "ok",
    1: "encoder slot overflow",
    2: "cdf index out of range",
    3: "Invalid `pmf`, non-finite or negative element found",
    4: "Invalid `pmf`: at least one element must have a non-zero probability.",
    5: "Invalid `pmf`: no symbol to steal frequency from",
    6: "decoder ran past the end of the string",
}

class ConvDesc(Structure):
    """Mirror of ``cai_conv_desc`` (include/cai_b200.h)."""
    _fields_ = ([(n, c_void_p) for n in ("a_hi", "a_lo", "w_packed", "bias", "aux_hi", "aux_lo", "out_f32", "out_hi",
                                         "out_lo", "sq_hi", "sq_lo", "abs_hi", "abs_lo")]
                + [(n, c_int32) for n in ("N", "H", "W", "Cin", "Ho", "Wo", "Cout", "Hp", "Wp", "os", "o0y", "o0x", "is_",
                                          "ntaps", "BN", "epilogue")]
                + [("clamp_lo", c_float), ("clamp_hi", c_float), ("dy", c_int8 * 32), ("dx", c_int8 * 32), ("glen", c_int8 * 32),
                   ("gdn_w", c_void_p), ("gdn_beta", c_void_p), ("gdn_mode", c_int32), ("mode", c_int32)])


class ArDesc(Structure):
    """Mirror of ``cai_ar_desc`` (include/cai_b200.h)."""
    _fields_ = ([(n, c_void_p) for n in ("w_ctx", "b_ctx", "w1", "b1", "w2", "b2", "w3", "b3", "params", "scale_table")]
                + [("scale_bound", c_float), ("slope", c_float)]
                + [(n, c_int32) for n in ("T", "B", "H", "W", "M", "P", "n_ctx", "n1", "n2", "n3", "ksize", "cluster",
                                          "group", "flags")])


# name -> (restype, argtypes).  Must list every symbol declared in include/cai_b200.h
# (tests/test_abi.py cross-checks this table against the header and the built library).
SIGNATURES = {
    "cai_abi_version": (c_int, []),
    "cai_last_error": (c_char_p, []),
    "cai_device_info": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "cai_table_create": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, POINTER(c_void_p)]),
    "cai_table_destroy": (None, [c_void_p]),
    "cai_table_info": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int64), POINTER(c_int32), POINTER(c_int32)]),
    "cai_rans_slot_words": (c_int64, [c_int64]),
    "cai_rans_encode_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64,
                                      c_void_p, c_void_p, c_void_p]),
    "cai_rans_compact": (c_int, [c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "cai_rans_decode_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                      c_void_p, c_int32, c_void_p, c_void_p]),
    "cai_gc_quantize_index": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_int32, c_int64,
                                      c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "cai_eb_quantize_index": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                      c_void_p]),
    "cai_dequantize": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "cai_pixels_u8_to_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cai_pixels_f32_to_u8": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cai_pmf_to_quantized_cdf": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                         c_void_p]),
    "cai_gc_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_float, c_float, c_int64, c_void_p,
                               c_void_p, c_void_p]),
    "cai_gc_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int64, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "cai_eb_forward": (c_int, [c_void_p, c_void_p, POINTER(c_int32), c_int32, c_void_p, c_void_p, c_int32, c_float,
                               c_int32, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "cai_eb_backward": (c_int, [c_void_p, c_void_p, POINTER(c_int32), c_int32, c_void_p, c_float, c_int32, c_int64,
                                c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "cai_eb_logits": (c_int, [c_void_p, c_void_p, POINTER(c_int32), c_int32, c_void_p, c_int64, c_int64, c_void_p,
                              c_void_p, c_void_p]),
    "cai_conv_gemm": (c_int, [POINTER(ConvDesc), c_void_p]),
    "cai_conv_tma_eligible": (c_int, [POINTER(ConvDesc)]),
    "cai_conv_wgrad_workspace": (c_int64, [c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "cai_conv_wgrad": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                               c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "cai_split_planes": (c_int, [c_void_p, c_int32, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "cai_im2col_split": (c_int, [c_void_p] + [c_int32] * 11 + [c_void_p, c_void_p, c_void_p]),
    "cai_col2im": (c_int, [c_void_p, c_void_p] + [c_int32] * 11 + [c_float, c_float, c_void_p, c_void_p]),
    "cai_gdn_bwd_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "cai_gdn_bwd_finish": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    "cai_gdn_bwd_params": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "cai_ar_encode": (c_int, [POINTER(ArDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cai_ar_decode": (c_int, [POINTER(ArDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class CaiError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libcai_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC_DIR, "-j8"], stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CaiError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header / library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if L.cai_abi_version() != 2:
            raise CaiError("libcai_b200.so ABI version mismatch")
        _lib = L
    return _lib


# kernels launched per successful C-ABI call (bench.py reports the total as "gpu_launches")
KERNELS_PER_CALL = {"cai_table_create": 3, "cai_rans_encode_batch": 1, "cai_rans_compact": 2,
                    "cai_rans_decode_batch": 1, "cai_gc_quantize_index": 1, "cai_eb_quantize_index": 1,
                    "cai_dequantize": 1, "cai_pmf_to_quantized_cdf": 1, "cai_gc_forward": 1, "cai_gc_backward": 1,
                    "cai_eb_forward": 1, "cai_eb_backward": 1, "cai_eb_logits": 1, "cai_conv_gemm": 1,
                    "cai_split_planes": 1, "cai_im2col_split": 1, "cai_col2im": 1, "cai_gdn_bwd_prepare": 1,
                    "cai_gdn_bwd_finish": 1, "cai_gdn_bwd_params": 1, "cai_conv_wgrad": 4}
LAUNCHES = 0


def check(rc: int, what: str = "") -> None:
    global LAUNCHES
    LAUNCHES += KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        msg = lib().cai_last_error().decode("utf-8", "replace")
        raise CaiError(f"{what or 'libcai_b200'} failed (rc={rc}): {msg}")


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise CaiError(f"{name} must live on a CUDA device (got {t.device}); there is no CPU fallback")
    return t


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def current_stream():
    import torch

    return c_void_p(torch.cuda.current_stream().cuda_stream)
