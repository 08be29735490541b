"""GPU tests of the entropy models: the reference's own test contracts (tests/test_entropy_models.py) run
against the drop-in classes, plus golden / oracle parity for tables, likelihoods and byte strings."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
LIK_TOL = 2e-6      # max-abs tolerance on likelihoods (fp32, erfc / sigmoid ulp differences CPU vs GPU)
GRAD_RTOL = 2e-4    # relative (to max |ref|) tolerance on gradients


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _close(got, ref, rtol):
    got, ref = got.detach().cpu().numpy().astype(np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    scale = max(np.abs(ref).max(), 1e-12)
    assert np.abs(got - ref).max() <= rtol * scale, (np.abs(got - ref).max(), scale)


@pytest.fixture
def em():
    from compressai_environment_b200.entropy_models import EntropyModel

    return EntropyModel().to(DEV)


class TestEntropyModel:
    def test_quantize_invalid(self, em):
        x = torch.rand(1, 3, 4, 4, device=DEV)
        with pytest.raises(ValueError):
            em.quantize(x, mode="toto")

    def test_quantize_noise(self, em):
        x = torch.rand(1, 3, 4, 4, device=DEV)
        y = em.quantize(x, "noise")
        assert y.shape == x.shape
        assert ((y - x) <= 0.5).all() and ((y - x) >= -0.5).all() and (y != torch.round(x)).any()

    def test_quantize(self, em):
        x = torch.rand(1, 3, 4, 4, device=DEV) * 9 - 4
        means = torch.rand(1, 3, 4, 4, device=DEV)
        assert (em.quantize(x, "dequantize") == torch.round(x)).all()
        assert (em.quantize(x, "dequantize", means) == torch.round(x - means) + means).all()
        s = em.quantize(x, "symbols")
        assert s.dtype == torch.int32 and (s == torch.round(x).int()).all()
        assert (em.quantize(x, "symbols", means) == torch.round(x - means).int()).all()

    def test_forward(self, em):
        with pytest.raises(NotImplementedError):
            em()

    def test_invalid_coder(self):
        from compressai_environment_b200.entropy_models import EntropyModel

        with pytest.raises(ValueError):
            EntropyModel(entropy_coder="huffman")
        with pytest.raises(ValueError):
            EntropyModel(entropy_coder=0xFF)

    def test_invalid_inputs(self, em):
        with pytest.raises(TypeError):
            em.compress(torch.rand(1, 3))
        with pytest.raises(ValueError):
            em.compress(torch.rand(1, 3), torch.rand(2, 3))
        with pytest.raises(ValueError):
            em.compress(torch.rand(1, 3, 1, 1), torch.rand(2, 3))

    def test_invalid_cdf(self, em):
        x = torch.rand(1, 32, 16, 16, device=DEV)
        with pytest.raises(ValueError):
            em.compress(x, torch.rand(1, 32, 16, 16, device=DEV))

    def test_invalid_cdf_length_offsets(self, em):
        x = torch.rand(1, 32, 16, 16, device=DEV)
        idx = torch.rand(1, 32, 16, 16, device=DEV)
        em._quantized_cdf.resize_(32, 1)
        with pytest.raises(ValueError):
            em.compress(x, idx)
        em._cdf_length.resize_(32, 1)
        with pytest.raises(ValueError):
            em.compress(x, idx)
        em._cdf_length.resize_(32)
        with pytest.raises(ValueError):
            em.compress(x, idx)

    def test_invalid_decompress(self, em):
        with pytest.raises(TypeError):
            em.decompress(["ssss"])
        with pytest.raises(ValueError):
            em.decompress("sss", torch.rand(1, 3, 4, 4))
        with pytest.raises(ValueError):
            em.decompress(["sss"], torch.rand(1, 4, 4))
        with pytest.raises(ValueError):
            em.decompress(["sss"], torch.rand(2, 4, 4))
        with pytest.raises(ValueError):
            em.decompress(["sss"], torch.rand(1, 4, 4), torch.rand(2, 4, 4))


class TestEntropyBottleneck:
    def _eb(self, C=128):
        from compressai_environment_b200.entropy_models import EntropyBottleneck

        return EntropyBottleneck(C).to(DEV)

    def test_forward_training(self):
        eb = self._eb()
        x = torch.rand(1, 128, 32, 32, device=DEV)
        y, lik = eb(x)
        assert y.shape == x.shape and lik.shape == x.shape
        assert ((y - x) <= 0.5).all() and ((y - x) >= -0.5).all() and (y != torch.round(x)).any()

    def test_forward_inference_nd(self):
        eb = self._eb().eval()
        x = torch.rand(1, 128, device=DEV)
        y, lik = eb(x)
        assert y.shape == x.shape and lik.shape == x.shape and (y == torch.round(x)).all()
        for i in range(1, 6):
            x = torch.rand(1, 128, *([4] * i), device=DEV)
            y, lik = eb(x)
            assert y.shape == x.shape and lik.shape == x.shape and (y == torch.round(x)).all()

    def test_loss(self):
        loss = self._eb().loss()
        assert len(loss.size()) == 0 and loss.numel() == 1

    def test_update_flags(self):
        eb = self._eb(16)
        assert eb.update()
        assert not eb.update()
        assert not eb.update(force=False)
        assert eb.update(force=True)

    def test_compression_2d_nd(self):
        eb = self._eb()
        eb.update()
        x = torch.rand(1, 128, 32, 32, device=DEV)
        s = eb.compress(x)
        assert torch.allclose(torch.round(x), eb.decompress(s, x.size()[2:]))
        x = torch.rand(1, 128, device=DEV)
        assert torch.allclose(torch.round(x), eb.decompress(eb.compress(x), []))
        for i in range(1, 6):
            x = torch.rand(2, 128, *([4] * i), device=DEV) * 30 - 15
            assert torch.allclose(torch.round(x), eb.decompress(eb.compress(x), x.size()[2:]))

    def _golden_eb(self, golden):
        from compressai_environment_b200.entropy_models import EntropyBottleneck

        c = golden("cdf")
        eb = EntropyBottleneck(6).to(DEV)
        with torch.no_grad():
            for n, p in eb.named_parameters():
                p.copy_(_t(c["eb_" + n]))
        return eb, c

    def test_update_matches_reference_table(self, golden):
        """Offsets / lengths exact; CDF entries within the reference's own +-2 tolerance
        (tests/test_entropy_models.py:393-398) because the float pmf comes from GPU sigmoid/tanh."""
        eb, c = self._golden_eb(golden)
        eb.update(force=True)
        assert np.array_equal(eb._offset.cpu().numpy(), c["eb_off"])
        assert np.array_equal(eb._cdf_length.cpu().numpy(), c["eb_len"])
        d = np.abs(eb._quantized_cdf.cpu().numpy().astype(np.int64) - c["eb_cdf"].astype(np.int64))
        assert d.max() <= 2, d.max()

    def test_likelihood_golden(self, golden):
        eb, _ = self._golden_eb(golden)
        f = golden("fp")
        x = _t(f["ebl_x"]).requires_grad_()
        v = x.permute(1, 0, 2, 3).reshape(6, 1, -1)
        lik = eb.likelihood_lower_bound(eb._likelihood(v))
        ref_lik = np.transpose(f["ebl_lik"], (1, 0, 2, 3)).reshape(6, 1, -1)
        assert np.abs(lik.detach().cpu().numpy() - ref_lik).max() <= LIK_TOL
        gout = _t(np.transpose(f["ebl_gout"], (1, 0, 2, 3)).reshape(6, 1, -1))
        lik.backward(gout)
        _close(x.grad, f["ebl_gx"], GRAD_RTOL)
        for n, p in eb.named_parameters():
            if "ebl_g" + n in f:
                _close(p.grad, f["ebl_g" + n], GRAD_RTOL)

    def test_forward_fused_bound_and_layouts(self, golden):
        eb, _ = self._golden_eb(golden)
        f = golden("fp")
        eb.eval()
        x = _t(f["ebl_x"])
        for fmt in (torch.contiguous_format, torch.channels_last):
            out, lik = eb(x.contiguous(memory_format=fmt))
            assert np.array_equal(out.detach().cpu().numpy(), f["ebf_out"])
            assert np.abs(lik.detach().cpu().numpy() - f["ebf_lik"]).max() <= LIK_TOL

    def test_loss_golden(self, golden):
        eb, _ = self._golden_eb(golden)
        f = golden("fp")
        loss = eb.loss()
        _close(loss, f["eb_loss"], 1e-5)
        loss.backward()
        _close(eb.quantiles.grad, f["eb_loss_gquantiles"], GRAD_RTOL)
        assert eb._matrix0.grad is None

    def test_training_backward_matches_torch(self):
        """Training-mode forward/backward against a plain torch fp32 restatement with the same noise."""
        eb = self._eb(5)
        with torch.no_grad():
            for n, p in eb.named_parameters():
                if n != "quantiles":
                    p.add_(0.2 * torch.randn_like(p))
        x = (torch.randn(2, 5, 3, 7, device=DEV) * 3).requires_grad_()
        torch.manual_seed(1)
        out, lik = eb(x, training=True)
        (lik.log().sum() + (out * out).sum()).backward()
        g_fused = {n: p.grad.clone() for n, p in eb.named_parameters() if p.grad is not None}
        gx = x.grad.clone()
        eb.zero_grad()
        x.grad = None
        torch.manual_seed(1)
        noise = torch.empty_like(x).uniform_(-0.5, 0.5)
        xt = x + noise

        def logits(v):
            h = v.permute(1, 0, 2, 3).reshape(5, 1, -1)
            for i in range(5):
                h = torch.matmul(torch.nn.functional.softplus(getattr(eb, f"_matrix{i}")), h) + getattr(eb, f"_bias{i}")
                if i < 4:
                    h = h + torch.tanh(getattr(eb, f"_factor{i}")) * torch.tanh(h)
            return h.reshape(5, 2, 3, 7).permute(1, 0, 2, 3)

        lo, up = logits(xt - 0.5), logits(xt + 0.5)
        sg = -torch.sign(lo + up).detach()
        L = eb.likelihood_lower_bound(torch.abs(torch.sigmoid(sg * up) - torch.sigmoid(sg * lo)))
        assert torch.equal(out, xt)
        assert (lik - L).abs().max() <= LIK_TOL
        (L.log().sum() + (xt * xt).sum()).backward()
        _close(gx, x.grad.cpu().numpy(), GRAD_RTOL)
        for n, p in eb.named_parameters():
            if p.grad is not None:
                _close(g_fused[n], p.grad.cpu().numpy(), 5e-4)


class TestGaussianConditional:
    def _gc(self, table=None):
        from compressai_environment_b200.entropy_models import GaussianConditional

        return GaussianConditional(table).to(DEV)

    def test_invalid_scale_table(self):
        from compressai_environment_b200.entropy_models import GaussianConditional

        for bad in (1, [], (), torch.rand(10)):
            with pytest.raises(ValueError):
                GaussianConditional(bad)
        for bad in ([2, 1], [0, 1, 2], [1, 1, -1], [-1, 1], [1, 2, 0]):
            with pytest.raises(ValueError):
                GaussianConditional(bad)

    def test_scale_bound(self):
        from compressai_environment_b200.entropy_models import GaussianConditional

        with pytest.raises(ValueError):
            GaussianConditional([1, 2], scale_bound=-0.1)
        with pytest.raises(Exception):
            GaussianConditional([1, 2], scale_bound=None)

    def test_forward_training_and_eval(self):
        gc = self._gc()
        x = torch.rand(1, 128, 32, 32, device=DEV)
        scales = torch.rand(1, 128, 32, 32, device=DEV)
        y, lik = gc(x, scales)
        assert y.shape == x.shape and lik.shape == x.shape
        assert ((y - x) <= 0.5).all() and ((y - x) >= -0.5).all() and (y != torch.round(x)).any()
        gc.eval()
        y, lik = gc(x, scales)
        assert (y == torch.round(x)).all()
        means = torch.rand(1, 128, 32, 32, device=DEV)
        y, lik = gc(x, scales, means)
        assert (y == torch.round(x - means) + means).all()

    def test_update_matches_reference_table(self, golden):
        c = golden("cdf")
        gc = self._gc()
        assert gc.update_scale_table(c["gc_scale_table"].tolist())
        assert not gc.update_scale_table(c["gc_scale_table"].tolist())
        assert np.array_equal(gc._offset.cpu().numpy(), c["gc_off"])
        assert np.array_equal(gc._cdf_length.cpu().numpy(), c["gc_len"])
        # float half: pmf within fp tolerance of the reference's CPU pmf
        pmf, tail, plen, _ = gc._pmf()
        assert np.array_equal(plen.cpu().numpy(), c["gc_pmf_len"])
        assert np.abs(pmf.cpu().numpy() - c["gc_pmf"]).max() <= 2e-7
        assert np.abs(tail.cpu().numpy().ravel() - c["gc_tail"]).max() <= 1e-12
        # integer half on the REFERENCE's pmf is bit-exact (see also test_coder_gpu.py::test_pmf_gc_table_rows);
        # on the GPU pmf the table must be a valid one: strictly increasing rows ending at 2^16, and narrow rows
        # (where no normalisation cliff exists) must coincide with the reference table
        cdf = gc._quantized_cdf.cpu().numpy().astype(np.int64)
        for k in range(64):
            row = cdf[k, :c["gc_len"][k]]
            assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) > 0).all()
        narrow = c["gc_len"] <= 64
        assert np.abs(cdf[narrow] - c["gc_cdf"][narrow]).max() <= 2

    def test_likelihood_golden(self, golden):
        f = golden("fp")
        gc = self._gc()
        y, s, m = (_t(f[k]).requires_grad_() for k in ("gcl_y", "gcl_s", "gcl_m"))
        lik = gc.likelihood_lower_bound(gc._likelihood(y, s, m))
        assert np.abs(lik.detach().cpu().numpy() - f["gcl_lik"]).max() <= LIK_TOL
        lik.backward(_t(f["gcl_gout"]))
        _close(y.grad, f["gcl_gy"], GRAD_RTOL)
        _close(s.grad, f["gcl_gs"], GRAD_RTOL)
        _close(m.grad, f["gcl_gm"], GRAD_RTOL)

    def test_fused_forward_bound_gate(self, golden):
        """forward() folds the 1e-9 LowerBound and its gradient gate into the kernels."""
        f = golden("fp")
        gc = self._gc()
        y, s, m = (_t(f[k]).requires_grad_() for k in ("gcl_y", "gcl_s", "gcl_m"))
        _, lik = _apply_mode2(gc, y, s, m)
        assert np.abs(lik.detach().cpu().numpy() - f["gcl_lik"]).max() <= LIK_TOL
        lik.backward(_t(f["gcl_gout"]))
        _close(y.grad, f["gcl_gy"], GRAD_RTOL)
        _close(s.grad, f["gcl_gs"], GRAD_RTOL)

    def test_compress_bytes_vs_oracle(self, golden, orc):
        c = golden("cdf")
        gc = self._gc()
        gc.scale_table = _t(c["gc_scale_table"])
        gc._quantized_cdf, gc._cdf_length, gc._offset = _t(c["gc_cdf"]), _t(c["gc_len"]), _t(c["gc_off"])
        rng = np.random.default_rng(4)
        shape = (3, 24, 6, 10)
        scales = np.exp(rng.random(shape) * 8 - 2.5).astype(np.float32)
        means = (rng.random(shape) * 4 - 2).astype(np.float32)
        y = (means + rng.standard_normal(shape) * scales * 1.5).astype(np.float32)
        for fmt in (torch.contiguous_format, torch.channels_last):
            ty, ts, tm = (_t(a).contiguous(memory_format=fmt) for a in (y, scales, means))
            idx = gc.build_indexes(ts)
            ref_idx = orc.gc_build_indexes(scales, c["gc_scale_table"])
            assert np.array_equal(idx.cpu().numpy(), ref_idx)
            strings = gc.compress(ty, idx, tm)
            sym = orc.quantize_symbols(y, means)
            ref = [orc.rans_encode(sym[b], ref_idx[b], c["gc_cdf"], c["gc_len"], c["gc_off"]) for b in range(3)]
            assert strings == ref
            enc, _ = gc.compress_from_scales(ty, ts, tm)
            assert enc.to_bytes() == ref
            y_hat = gc.decompress(strings, idx, means=tm)
            assert np.array_equal(y_hat.cpu().numpy(), orc.dequantize(sym, means))
            y_hat2 = gc.decompress_from_scales(strings, ts, tm)
            assert np.array_equal(y_hat2.cpu().numpy(), orc.dequantize(sym, means))

    def test_deepcopy_and_state_dict(self, golden):
        from compressai_environment_b200.entropy_models import EntropyBottleneck, GaussianConditional

        gc = GaussianConditional([0.5, 1.0, 2.0]).to(DEV)
        gc.update()
        gc2 = copy.deepcopy(gc)
        assert torch.equal(gc2._quantized_cdf, gc._quantized_cdf)
        eb = EntropyBottleneck(4).to(DEV)
        eb.update()
        sd = eb.state_dict()
        assert {"_offset", "_quantized_cdf", "_cdf_length", "target", "quantiles", "_matrix0", "_bias4"} <= set(sd)
        assert sd["_quantized_cdf"].dtype == torch.int32


def _apply_mode2(gc, y, s, m):
    from compressai_environment_b200.entropy_models.entropy_models import _GaussianLikelihood

    return _GaussianLikelihood.apply(y, s, m, None, 2, gc._bound_scale(), gc._lik_bound())


def test_pixel_conversions_exact():
    """cai_pixels_u8_to_f32 / cai_pixels_f32_to_u8 against the torch definitions (ToTensor's x / 255; round half to
    even of clamp(x, 0, 1) * 255), including ragged tails, NaN and out-of-range values."""
    from compressai_environment_b200 import kernels

    g = torch.Generator().manual_seed(5)
    for n in (0, 1, 15, 16, 17, 4099, 3 * 37 * 53):
        u_cpu = torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g)
        u = u_cpu.to(DEV)
        f = kernels.pixels_to_float(u)
        # the reference's callers convert on the HOST (ToTensor): a true IEEE division.  (torch's CUDA division by a
        # scalar multiplies by the reciprocal instead and differs in the last bit for 126 of the 256 values.)
        assert torch.equal(f.cpu(), u_cpu.float() / 255.0)
        assert torch.equal(kernels.pixels_to_u8(f), u)
        x = (torch.rand(n, generator=g) * 1.4 - 0.2).to(DEV)
        if n > 3:
            x[1], x[2], x[3] = float("nan"), 0.5 / 255.0, 1.5 / 255.0   # NaN -> 0; ties go to the even value
        want = (torch.nan_to_num(x.cpu(), nan=0.0).clamp(0, 1) * 255.0).round().to(torch.uint8)
        assert torch.equal(kernels.pixels_to_u8(x).cpu(), want)
