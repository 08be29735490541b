"""CPU suite part 1: pin the ORACLE (oracle/cai_oracle.c + oracle/oracle.py) against the golden vectors
produced by the compiled reference (tests/golden/make_golden.py) and, when oracle/_ref is present,
against the reference itself on fresh random inputs."""
import hashlib

import numpy as np
import pytest

CODER_KEYS = ["katA", "katB", "gauss_t1_n4096", "gauss_t4_n4096", "gauss_t16_n2048", "gauss_t1_n2",
              "gauss_t1_n3", "gauss_t1_n31", "gauss_t1_n32", "gauss_t1_n33", "edge"]


def _tables(g, c, key):
    if key.startswith("kat") and key[3] in "AB":
        return g["small_cdf"], g["small_len"], g["small_off"]
    return c["gc_cdf"], c["gc_len"], c["gc_off"]


@pytest.mark.parametrize("key", CODER_KEYS)
def test_coder_golden(orc, golden, key):
    g, c = golden("coder"), golden("cdf")
    cdf, ln, off = _tables(g, c, key)
    sym, idx, ref = g[key + "_sym"], g[key + "_idx"], g[key + "_bytes"].tobytes()
    got = orc.rans_encode(sym, idx, cdf, ln, off)
    assert got == ref
    assert np.array_equal(orc.rans_decode(ref, idx, cdf, ln, off), sym)


def test_coder_survey_kats(orc, golden):
    """Literal known answers quoted in SURVEY.md 8(c) (independent of the npz files)."""
    g, c = golden("coder"), golden("cdf")
    assert g["katA_bytes"].tobytes().hex() == "d2a25d5cf114e60161ffffad31103142f0ff1f02317749b2"
    b = orc.rans_encode(g["katB_sym"], g["katB_idx"], g["small_cdf"], g["small_len"], g["small_off"])
    assert len(b) == 5908
    assert hashlib.sha256(b).hexdigest() == "8e362f68c17ec9741833fb7a39f90bf210a50d77ac89084c304207480efd1dff"
    z = np.zeros(4, np.int32)
    assert orc.rans_encode([0, 0, 0, 0], z, c["gc_cdf"], c["gc_len"], c["gc_off"]).hex() == "2e00068000000000"
    assert orc.rans_encode([100000, -100000, 5, -7], z, c["gc_cdf"], c["gc_len"], c["gc_off"]).hex() == \
        "ffff008000000000c5d30000f3ff5f3d0df3ff1ff6ff1f0b"


def test_coder_empty_and_single(orc, golden):
    """N = 0 / 1 are undefined in the reference (heap under-run); the oracle defines them."""
    c = golden("cdf")
    e = orc.rans_encode([], [], c["gc_cdf"], c["gc_len"], c["gc_off"])
    assert e == (0x80000000).to_bytes(4, "little") + (0).to_bytes(4, "little")
    one = orc.rans_encode([3], [10], c["gc_cdf"], c["gc_len"], c["gc_off"])
    assert len(one) == 8 and orc.rans_decode(one, [10], c["gc_cdf"], c["gc_len"], c["gc_off"]).tolist() == [3]


@pytest.mark.parametrize("key", ["katref", "katC", "katD", "katE", "rand0", "rand1", "rand2", "rand3", "rand4"])
def test_pmf_golden(orc, golden, key):
    c = golden("cdf")
    prec = int(c[key + "_prec"]) if key + "_prec" in c else 16
    assert orc.pmf_to_quantized_cdf(c[key + "_pmf"], prec) == c[key + "_cdf"].tolist()


def test_pmf_reference_kat_literal(orc):
    assert orc.pmf_to_quantized_cdf([0.1, 0.2, 0, 0], 16) == [0, 21845, 65534, 65535, 65536]  # test_ops.py:104-106
    for bad in ([-0.1, 0.5], [float("inf"), 0.5], [float("nan"), 0.5], [0.0, 0.0]):
        with pytest.raises(ValueError):
            orc.pmf_to_quantized_cdf(bad, 16)


def test_gc_table_golden(orc, golden):
    c = golden("cdf")
    got = orc.pmf_rows_to_cdf(c["gc_pmf"], c["gc_pmf_len"], c["gc_tail"], 16)
    assert np.array_equal(got, c["gc_cdf"])


def test_float_steps_golden(orc, golden):
    f, c = golden("fp"), golden("cdf")
    assert np.array_equal(orc.quantize_symbols(f["q_y"], f["q_mu"]), f["q_sym"])
    assert np.array_equal(orc.quantize_symbols(f["q_y"]), f["q_sym_nomean"])
    assert np.array_equal(orc.gc_build_indexes(f["q_scales"], c["gc_scale_table"]), f["q_idx"])
    assert np.array_equal(orc.get_scale_table(), c["gc_scale_table"])
    for tag, inv in (("gdn", False), ("igdn", True)):
        ped = np.float32(2.0 ** -36)
        beta = np.maximum(f[tag + "_beta"], np.float32((1e-6 + 2.0 ** -36) ** 0.5)) ** 2 - ped
        gamma = np.maximum(f[tag + "_gamma"], np.float32(2.0 ** -18)) ** 2 - ped
        assert np.abs(orc.gdn(f[tag + "_x"], beta, gamma, inv) - f[tag + "_y"]).max() < 2e-6


@pytest.mark.ref
def test_coder_vs_reference_random(orc, golden):
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built")
    orc.import_ref()
    from compressai import ans

    c = golden("cdf")
    cdf, ln, off = c["gc_cdf"], c["gc_len"], c["gc_off"]
    lists = (cdf.tolist(), ln.tolist(), off.tolist())
    rng = np.random.default_rng(0)
    tab = c["gc_scale_table"]
    for t, n in ((1, 5000), (3, 3000), (40, 700), (1, 2), (1, 17)):
        idx = rng.integers(0, 64, n).astype(np.int32)
        sym = np.rint(rng.standard_normal(n) * tab[idx] * t).astype(np.int32)
        ref = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), *lists)
        assert orc.rans_encode(sym, idx, cdf, ln, off) == ref
        assert ans.RansDecoder().decode_with_indexes(ref, idx.tolist(), *lists) == \
            orc.rans_decode(ref, idx, cdf, ln, off).tolist()


@pytest.mark.ref
def test_pmf_vs_reference_random(orc):
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built")
    orc.import_ref()
    from compressai import _CXX

    rng = np.random.default_rng(1)
    for m in (2, 5, 33, 200, 900):
        for _ in range(3):
            p = rng.random(m).astype(np.float32) ** rng.integers(1, 12)
            p[rng.random(m) < 0.4] = 0
            p[rng.integers(0, m)] = 1.0
            p = (p / p.sum()).astype(np.float32)
            assert orc.pmf_to_quantized_cdf(p, 16) == _CXX.pmf_to_quantized_cdf(p.tolist(), 16)
