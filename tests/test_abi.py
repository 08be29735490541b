"""CPU suite: the C-ABI library loads and exports every symbol include/cai_b200.h declares, and the
ctypes table in _lib.py covers exactly that set.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cai_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cai_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header():
    from compressai_environment_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), f"{s} declared in cai_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.lib().cai_abi_version() == 2


def test_slot_words_bound():
    from compressai_environment_b200 import _lib

    L = _lib.lib()
    for n in (0, 1, 2, 31, 32, 33, 65536, 10_444_800):
        w = L.cai_rans_slot_words(n)
        assert w % 32 == 0 and w * 32 >= 52 * n + 31 + 64


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from compressai_environment_b200 import _CXX, ans
    from compressai_environment_b200._lib import CaiError

    with pytest.raises(CaiError):
        ans.RansEncoder().encode_with_indexes([0], [0], [[0, 65536]], [2], [0])
    with pytest.raises(CaiError):
        _CXX.pmf_to_quantized_cdf([0.5, 0.5], 16)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "compressai_environment_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_conv_desc_layout_matches_header(tmp_path):
    """The ctypes mirror of cai_conv_desc must have the C compiler's size and field offsets."""
    import ctypes, subprocess

    from compressai_environment_b200._lib import ConvDesc

    fields = ["a_hi", "bias", "out_f32", "abs_lo", "N", "Cout", "is", "ntaps", "BN", "epilogue", "clamp_lo", "dy", "dx",
              "glen", "gdn_w", "gdn_beta", "gdn_mode"]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cai_b200.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(cai_conv_desc));\n'
                   + "".join(f'  printf(" %zu", offsetof(cai_conv_desc, {f}));\n' for f in fields)
                   + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got[0] == ctypes.sizeof(ConvDesc)
    for f, off in zip(fields, got[1:]):
        assert getattr(ConvDesc, "is_" if f == "is" else f).offset == off, f


def test_ar_desc_layout_matches_header(tmp_path):
    """Same check for the ctypes mirror of cai_ar_desc."""
    import ctypes, subprocess

    from compressai_environment_b200._lib import ArDesc

    fields = [n for n, _ in ArDesc._fields_]
    src = tmp_path / "layout_ar.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cai_b200.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(cai_ar_desc));\n'
                   + "".join(f'  printf(" %zu", offsetof(cai_ar_desc, {f}));\n' for f in fields)
                   + "  return 0;\n}\n")
    exe = tmp_path / "layout_ar"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got[0] == ctypes.sizeof(ArDesc)
    for f, off in zip(fields, got[1:]):
        assert getattr(ArDesc, f).offset == off, f


def test_interval_union():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.interval_union([]) == 0.0
    assert bench.interval_union([(0, 1), (2, 3)]) == 2.0
    assert bench.interval_union([(2, 5), (0, 3), (4, 4.5), (10, 11)]) == 6.0
