"""GPU parity tests of the rANS coder / table build / quantise kernels, through the C ABI, against the
oracle and the committed golden vectors (bit-exact)."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CODER_KEYS = ["katA", "katB", "gauss_t1_n4096", "gauss_t4_n4096", "gauss_t16_n2048", "gauss_t1_n2",
              "gauss_t1_n3", "gauss_t1_n31", "gauss_t1_n32", "gauss_t1_n33", "edge"]


def _tables(g, c, key):
    if key.startswith("kat") and key[3] in "AB":
        return g["small_cdf"], g["small_len"], g["small_off"]
    return c["gc_cdf"], c["gc_len"], c["gc_off"]


@pytest.fixture(scope="module")
def gc_table(golden):
    from compressai_environment_b200 import coder

    c = golden("cdf")
    dev = torch.device("cuda")
    return coder.CdfTable(torch.from_numpy(c["gc_cdf"]).to(dev), torch.from_numpy(c["gc_len"]).to(dev),
                          torch.from_numpy(c["gc_off"]).to(dev))


@pytest.mark.parametrize("key", CODER_KEYS)
def test_list_api_golden(golden, key):
    from compressai_environment_b200 import ans

    g, c = golden("coder"), golden("cdf")
    cdf, ln, off = _tables(g, c, key)
    sym, idx, ref = g[key + "_sym"], g[key + "_idx"], g[key + "_bytes"].tobytes()
    lists = (cdf.tolist(), ln.tolist(), off.tolist())
    got = ans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), *lists)
    assert got == ref
    assert ans.RansDecoder().decode_with_indexes(ref, idx.tolist(), *lists) == sym.tolist()


def test_survey_kats(golden):
    from compressai_environment_b200 import ans

    g, c = golden("coder"), golden("cdf")
    small = (g["small_cdf"].tolist(), g["small_len"].tolist(), g["small_off"].tolist())
    b = ans.RansEncoder().encode_with_indexes(g["katA_sym"].tolist(), g["katA_idx"].tolist(), *small)
    assert b.hex() == "d2a25d5cf114e60161ffffad31103142f0ff1f02317749b2"
    b = ans.RansEncoder().encode_with_indexes(g["katB_sym"].tolist(), g["katB_idx"].tolist(), *small)
    assert len(b) == 5908
    assert hashlib.sha256(b).hexdigest() == "8e362f68c17ec9741833fb7a39f90bf210a50d77ac89084c304207480efd1dff"
    gct = (c["gc_cdf"].tolist(), c["gc_len"].tolist(), c["gc_off"].tolist())
    assert ans.RansEncoder().encode_with_indexes([0] * 4, [0] * 4, *gct).hex() == "2e00068000000000"
    assert ans.RansEncoder().encode_with_indexes([100000, -100000, 5, -7], [0] * 4, *gct).hex() == \
        "ffff008000000000c5d30000f3ff5f3d0df3ff1ff6ff1f0b"


def test_empty_and_single(golden, orc):
    from compressai_environment_b200 import ans

    c = golden("cdf")
    gct = (c["gc_cdf"].tolist(), c["gc_len"].tolist(), c["gc_off"].tolist())
    assert ans.RansEncoder().encode_with_indexes([], [], *gct) == bytes.fromhex("0000008000000000")
    one = ans.RansEncoder().encode_with_indexes([3], [10], *gct)
    assert one == orc.rans_encode([3], [10], c["gc_cdf"], c["gc_len"], c["gc_off"])
    assert ans.RansDecoder().decode_with_indexes(one, [10], *gct) == [3]
    assert ans.RansDecoder().decode_with_indexes(one, [], *gct) == []


def test_buffered_and_streaming(golden):
    from compressai_environment_b200 import ans

    g, c = golden("coder"), golden("cdf")
    gct = (c["gc_cdf"].tolist(), c["gc_len"].tolist(), c["gc_off"].tolist())
    sym, idx = g["gauss_t4_n4096_sym"].tolist(), g["gauss_t4_n4096_idx"].tolist()
    ref = g["gauss_t4_n4096_bytes"].tobytes()
    be = ans.BufferedRansEncoder()
    be.encode_with_indexes(sym[:1000], idx[:1000], *gct)
    be.encode_with_indexes(sym[1000:], idx[1000:], *gct)
    assert be.flush() == ref
    d = ans.RansDecoder()
    d.set_stream(ref)
    got = d.decode_stream(idx[:1000], *gct)
    got += d.decode_stream(idx[1000:1001], *gct)
    got += d.decode_stream(idx[1001:], *gct)
    assert got == sym


def test_buffered_mixed_tables(golden, orc):
    """Two calls with different tables == one string over the stacked table."""
    from compressai_environment_b200 import ans

    g, c = golden("coder"), golden("cdf")
    small = (g["small_cdf"], g["small_len"], g["small_off"])
    big = (c["gc_cdf"][:8], c["gc_len"][:8], c["gc_off"][:8])
    s1, i1 = g["katA_sym"], g["katA_idx"]
    s2, i2 = g["gauss_t1_n33_sym"], g["gauss_t1_n33_idx"] % 8
    be = ans.BufferedRansEncoder()
    be.encode_with_indexes(s1.tolist(), i1.tolist(), *[a.tolist() for a in small])
    be.encode_with_indexes(s2.tolist(), i2.tolist(), *[a.tolist() for a in big])
    L = big[0].shape[1]
    cdf = np.concatenate([np.pad(small[0], ((0, 0), (0, L - small[0].shape[1]))), big[0]])
    ref = orc.rans_encode(np.concatenate([s1, s2]), np.concatenate([i1, i2 + 2]), cdf,
                          np.concatenate([small[1], big[1]]), np.concatenate([small[2], big[2]]))
    assert be.flush() == ref


@pytest.mark.parametrize("B,n,t", [(1, 1, 1.0), (7, 100, 1.0), (300, 257, 4.0), (64, 4096, 1.0), (33, 1000, 30.0),
                                   (2, 70001, 2.0)])
def test_batch_vs_oracle(golden, orc, gc_table, B, n, t):
    from compressai_environment_b200 import coder

    c = golden("cdf")
    cdf, ln, off, tab = c["gc_cdf"], c["gc_len"], c["gc_off"], c["gc_scale_table"]
    rng = np.random.default_rng(B * 1000 + n)
    idx = rng.integers(0, 64, (B, n)).astype(np.int32)
    sym = np.rint(rng.standard_normal((B, n)) * tab[idx] * t).astype(np.int32)
    dev = torch.device("cuda")
    enc = coder.encode(gc_table, torch.from_numpy(sym).to(dev), torch.from_numpy(idx).to(dev))
    got = enc.to_bytes()
    slots, nw = orc.rans_encode_batch(sym, idx, cdf, ln, off)
    ref = orc.slots_to_bytes(slots, nw)
    assert [len(b) for b in got] == [len(b) for b in ref]
    assert got == ref
    dec = coder.decode(gc_table, ref, torch.from_numpy(idx).to(dev))
    assert np.array_equal(dec.cpu().numpy(), sym)


def test_many_rows_table_global_fallback(orc):
    """A table too large for shared memory (in_smem = False) exercises the global-memory variant."""
    from compressai_environment_b200 import coder

    rng = np.random.default_rng(3)
    K, L = 600, 300
    cdf = np.zeros((K, L), np.int32)
    ln = rng.integers(3, L + 1, K).astype(np.int32)
    for k in range(K):
        m = ln[k] - 1
        cuts = np.sort(rng.choice(np.arange(1, 65536), m - 1, replace=False))
        cdf[k, 1:m] = cuts
        cdf[k, m] = 65536
    off = rng.integers(-50, 5, K).astype(np.int32)
    dev = torch.device("cuda")
    t = coder.CdfTable(torch.from_numpy(cdf).to(dev), torch.from_numpy(ln).to(dev), torch.from_numpy(off).to(dev))
    assert t.info()["in_smem"] is False
    B, n = 40, 3000
    idx = rng.integers(0, K, (B, n)).astype(np.int32)
    sym = rng.integers(-60, 320, (B, n)).astype(np.int32)
    got = coder.encode(t, torch.from_numpy(sym).to(dev), torch.from_numpy(idx).to(dev)).to_bytes()
    slots, nw = orc.rans_encode_batch(sym, idx, cdf, ln, off)
    assert got == orc.slots_to_bytes(slots, nw)
    dec = coder.decode(t, got, torch.from_numpy(idx).to(dev))
    assert np.array_equal(dec.cpu().numpy(), sym)


def test_bad_index_reports(gc_table):
    from compressai_environment_b200 import coder

    dev = torch.device("cuda")
    sym = torch.zeros((1, 10), dtype=torch.int32, device=dev)
    idx = torch.full((1, 10), 64, dtype=torch.int32, device=dev)
    with pytest.raises(ValueError):
        coder.encode(gc_table, sym, idx).to_bytes()


# ---- table build -------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["katref", "katC", "katD", "katE", "rand0", "rand1", "rand2", "rand3", "rand4"])
def test_pmf_golden(golden, key):
    from compressai_environment_b200 import _CXX

    c = golden("cdf")
    prec = int(c[key + "_prec"]) if key + "_prec" in c else 16
    assert _CXX.pmf_to_quantized_cdf(c[key + "_pmf"].tolist(), prec) == c[key + "_cdf"].tolist()


def test_pmf_errors():
    from compressai_environment_b200 import _CXX

    assert _CXX.pmf_to_quantized_cdf([0.1, 0.2, 0, 0], 16) == [0, 21845, 65534, 65535, 65536]
    for bad in ([-0.1, 0.5], [float("inf"), 0.5], [float("nan"), 0.5], [0.0, 0.0]):
        with pytest.raises(ValueError):
            _CXX.pmf_to_quantized_cdf(bad, 16)


def test_pmf_gc_table_rows(golden):
    from compressai_environment_b200 import _CXX

    c = golden("cdf")
    dev = torch.device("cuda")
    cdf = _CXX.pmf_rows_to_quantized_cdf(torch.from_numpy(c["gc_pmf"]).to(dev), torch.from_numpy(c["gc_pmf_len"]).to(dev),
                                         torch.from_numpy(c["gc_tail"]).to(dev), 16)
    assert np.array_equal(cdf.cpu().numpy(), c["gc_cdf"])


def test_pmf_random_vs_oracle(orc):
    from compressai_environment_b200 import _CXX

    rng = np.random.default_rng(5)
    dev = torch.device("cuda")
    K, Lp = 96, 700
    pmf = np.zeros((K, Lp), np.float32)
    ln = rng.integers(1, Lp + 1, K).astype(np.int32)
    for k in range(K):
        p = rng.random(ln[k]).astype(np.float32) ** rng.integers(1, 14)
        p[rng.random(ln[k]) < 0.5] = 0
        p[rng.integers(0, ln[k])] = 1.0
        pmf[k, :ln[k]] = p / p.sum()
    tail = (rng.random(K) * 1e-4).astype(np.float32)
    tail[::7] = 0
    ref = orc.pmf_rows_to_cdf(pmf, ln, tail, 16)
    got = _CXX.pmf_rows_to_quantized_cdf(torch.from_numpy(pmf).to(dev), torch.from_numpy(ln).to(dev),
                                         torch.from_numpy(tail).to(dev), 16)
    assert np.array_equal(got.cpu().numpy(), ref)


# ---- quantise / index ----------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
@pytest.mark.parametrize("shape", [(2, 5, 7, 9), (3, 192, 4, 6), (1, 33, 1, 65), (2, 64, 32, 48)])
def test_gc_quantize_index(golden, orc, layout, shape):
    from compressai_environment_b200 import kernels

    c = golden("cdf")
    rng = np.random.default_rng(sum(shape))
    y = (rng.standard_normal(shape) * 6).astype(np.float32)
    y.reshape(-1)[:8] = [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.4999, -3.5001]
    mu = rng.standard_normal(shape).astype(np.float32)
    sc = np.exp(rng.random(shape) * 9 - 3).astype(np.float32)
    flat = sc.reshape(-1)
    flat[:4] = [0.0, 0.11, 256.0, 1e9]
    flat[4:68] = c["gc_scale_table"][: min(64, flat.size - 4)] if flat.size >= 68 else flat[4:68]
    flat[-1] = np.nan
    dev = torch.device("cuda")
    mf = torch.channels_last if layout == "nhwc" else torch.contiguous_format
    ty, tm, ts = (torch.from_numpy(a).to(dev).contiguous(memory_format=mf) for a in (y, mu, sc))
    tab = torch.from_numpy(c["gc_scale_table"]).to(dev)
    sym, idx = kernels.gc_quantize_index(ty, ts, tm, tab, 0.11)
    assert np.array_equal(sym.cpu().numpy().reshape(shape), orc.quantize_symbols(y, mu))
    assert np.array_equal(idx.cpu().numpy().reshape(shape), orc.gc_build_indexes(sc, c["gc_scale_table"]))
    sym2, _ = kernels.gc_quantize_index(ty, None, None, tab, 0.11)
    assert np.array_equal(sym2.cpu().numpy().reshape(shape), orc.quantize_symbols(y))
    out = kernels.dequantize(sym, tm, None, shape, mf)
    assert np.array_equal(out.cpu().numpy(), orc.dequantize(orc.quantize_symbols(y, mu), mu))


def test_quantize_golden(golden):
    from compressai_environment_b200 import kernels

    f, c = golden("fp"), golden("cdf")
    dev = torch.device("cuda")
    y, mu, sc = (torch.from_numpy(f[k]).to(dev) for k in ("q_y", "q_mu", "q_scales"))
    tab = torch.from_numpy(c["gc_scale_table"]).to(dev)
    sym, idx = kernels.gc_quantize_index(y, sc, mu, tab, 0.11)
    assert np.array_equal(sym.cpu().numpy().reshape(y.shape), f["q_sym"])
    assert np.array_equal(idx.cpu().numpy().reshape(y.shape), f["q_idx"])


@pytest.mark.parametrize("layout", ["nchw", "nhwc"])
def test_eb_quantize_index(orc, layout):
    from compressai_environment_b200 import kernels

    rng = np.random.default_rng(9)
    shape = (3, 40, 5, 7)
    x = (rng.standard_normal(shape) * 5).astype(np.float32)
    med = rng.standard_normal(40).astype(np.float32)
    dev = torch.device("cuda")
    mf = torch.channels_last if layout == "nhwc" else torch.contiguous_format
    tx = torch.from_numpy(x).to(dev).contiguous(memory_format=mf)
    sym, idx = kernels.eb_quantize_index(tx, torch.from_numpy(med).to(dev))
    assert np.array_equal(sym.cpu().numpy().reshape(shape), orc.quantize_symbols(x, med[None, :, None, None]))
    assert np.array_equal(idx.cpu().numpy().reshape(shape),
                          np.broadcast_to(np.arange(40, dtype=np.int32)[None, :, None, None], shape))
    out = kernels.dequantize(sym, None, torch.from_numpy(med).to(dev), shape, mf)
    assert np.array_equal(out.cpu().numpy(),
                          orc.dequantize(orc.quantize_symbols(x, med[None, :, None, None]), med[None, :, None, None]))


@pytest.mark.parametrize("B", [32, 40, 97])
def test_lane_per_string_ragged_vs_oracle(golden, orc, gc_table, B):
    """The lane-per-string kernels (32 strings per warp, used from 32 strings per launch up) through the raw C ABI
    with RAGGED strings (``str_begin``): lengths 0 .. ~2000 incl. empty strings, chunk-boundary lengths (31, 32, 33,
    64), every table row, escapes on both sides (symbols far outside the rows), all compared with the oracle byte
    for byte and decoded back exactly."""
    from compressai_environment_b200 import _lib
    from compressai_environment_b200._lib import check, current_stream, lib, ptr

    c = golden("cdf")
    rng = np.random.default_rng(B)
    lens = rng.integers(0, 2000, B)
    lens[:6] = [0, 31, 32, 33, 64, 1]
    begin = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lens, out=begin[1:])
    total = int(begin[-1])
    idx = rng.integers(0, 64, total).astype(np.int32)
    sym = np.rint(rng.standard_normal(total) * c["gc_scale_table"][idx] * rng.choice([0.3, 1.0, 6.0], total)).astype(np.int32)
    sym[rng.integers(0, total, 50)] = rng.integers(-2_000_000, 2_000_000, 50)  # long escapes (up to 6 payload nibbles)
    dev = torch.device("cuda")
    d_sym, d_idx, d_beg = (torch.from_numpy(a).to(dev) for a in (sym, idx, begin))
    sw = int(lib().cai_rans_slot_words(int(lens.max())))
    slots = torch.zeros((B, sw), dtype=torch.int32, device=dev)
    n_words = torch.zeros(B, dtype=torch.int32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    check(lib().cai_rans_encode_batch(gc_table.handle, ptr(d_sym), ptr(d_idx), ptr(d_beg), 0, B, ptr(slots), sw,
                                      ptr(n_words), ptr(status), current_stream()), "cai_rans_encode_batch")
    torch.cuda.synchronize()
    assert int(status.abs().max()) == 0
    nw = n_words.cpu().numpy()
    sl = slots.cpu().numpy().view(np.uint32)
    for b in range(B):
        ours = sl[b, sw - nw[b]:].tobytes()
        ref = orc.rans_encode(sym[begin[b]:begin[b + 1]], idx[begin[b]:begin[b + 1]], c["gc_cdf"], c["gc_len"], c["gc_off"])
        assert ours == ref, (b, int(lens[b]))
    # decode in place from the slots (word_begin = end of slot - n_words, word_count = n_words)
    wb = (torch.arange(1, B + 1, device=dev, dtype=torch.int64) * sw) - n_words.to(torch.int64)
    out = torch.full((max(total, 1),), -7, dtype=torch.int32, device=dev)
    dstat = torch.zeros(B, dtype=torch.int32, device=dev)
    check(lib().cai_rans_decode_batch(gc_table.handle, ptr(slots), ptr(wb), ptr(n_words), ptr(d_idx), ptr(d_beg), 0, B,
                                      ptr(out), None, 0, ptr(dstat), current_stream()), "cai_rans_decode_batch")
    torch.cuda.synchronize()
    assert int(dstat.abs().max()) == 0
    assert np.array_equal(out.cpu().numpy()[:total], sym)
    # a truncated string is reported (status 6), a bad index too (status 2), without disturbing the other strings
    nw_cut = n_words.clone()
    victim = int(np.argmax(lens))
    nw_cut[victim] = max(1, int(nw[victim]) // 2)
    wb_cut = wb.clone()
    check(lib().cai_rans_decode_batch(gc_table.handle, ptr(slots), ptr(wb_cut), ptr(nw_cut), ptr(d_idx), ptr(d_beg), 0, B,
                                      ptr(out), None, 0, ptr(dstat), current_stream()), "cai_rans_decode_batch")
    torch.cuda.synchronize()
    ds = dstat.cpu().numpy()
    assert ds[victim] == 6 and (np.delete(ds, victim) == 0).all()
    bad = d_idx.clone()
    bad[int(begin[victim])] = 64
    check(lib().cai_rans_encode_batch(gc_table.handle, ptr(d_sym), ptr(bad), ptr(d_beg), 0, B, ptr(slots), sw,
                                      ptr(n_words), ptr(status), current_stream()), "cai_rans_encode_batch")
    torch.cuda.synchronize()
    es = status.cpu().numpy()
    assert es[victim] == 2 and (np.delete(es, victim) == 0).all()


def test_lane_per_string_kernels_enabled():
    """The same ragged / oracle test with the lane-per-string kernels switched on (CAI_CODER_LANES is read once per
    process, hence the subprocess).  They are off by default (slower, see rans.cu) but must stay bit-exact."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, CAI_CODER_LANES="32")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-p",
                          "no:cacheprovider", "-k", "test_lane_per_string_ragged_vs_oracle"], capture_output=True,
                         text=True, env=env, timeout=600)
    assert out.returncode == 0 and "3 passed" in out.stdout, out.stdout[-2000:] + out.stderr[-1000:]
