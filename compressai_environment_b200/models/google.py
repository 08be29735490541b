"""The model classes on the hot path with the reference's interface (compressai/models/google.py):
``CompressionModel`` (:56-116), ``FactorizedPrior`` (:119-191), ``ScaleHyperprior`` (:204-321),
``MeanScaleHyperprior`` (:324-392), ``JointAutoregressiveHierarchicalPriors`` (:395-661) and ``get_scale_table``
(:195-201).

Same constructor arguments, submodule names (``g_a``, ``g_s``, ``h_a``, ``h_s``, ``entropy_bottleneck``,
``gaussian_conditional``) and ``state_dict`` keys, same return dictionaries.  Differences underneath:
  * activations are channels-last between layers;
  * ``compress`` never decodes its own z strings: the decoder's ``z_hat`` is by construction
    ``dequantize(quantize(z))``, which the fused quantize kernel already produced (reference :307,
    :374 call ``decompress`` on the strings it just wrote);
  * quantize + build_indexes + the whole batch's strings are one fused kernel + one coder launch.
"""
import math
import warnings

import os

import torch
import torch.nn as nn

from .. import _cache, _lib, coder, kernels
from ..entropy_models import EntropyBottleneck, GaussianConditional
from ..layers import GDN, MaskedConv2d
from ..transforms import Conv2d, TransformStack
from .utils import conv, deconv, update_registered_buffers

__all__ = [
    "CompressionModel",
    "FactorizedPrior",
    "ScaleHyperprior",
    "MeanScaleHyperprior",
    "JointAutoregressiveHierarchicalPriors",
    "get_scale_table",
    "SCALES_MIN",
    "SCALES_MAX",
    "SCALES_LEVELS",
]

_CL = torch.channels_last


def _nhwc(x):
    return x.contiguous(memory_format=_CL) if x.dim() == 4 else x


class CompressionModel(nn.Module):
    """Base class for an auto-encoder with at least one entropy bottleneck module."""

    def __init__(self, entropy_bottleneck_channels, init_weights=None):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)
        if init_weights is not None:
            warnings.warn("init_weights was removed as it was never functional", DeprecationWarning)

    def aux_loss(self):
        """Aggregated loss over the auxiliary entropy bottleneck module(s)."""
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def forward(self, *args):
        raise NotImplementedError()

    def update(self, force=False):
        """Updates the entropy bottleneck(s) CDF values; True if one of them was updated."""
        updated = False
        for m in self.children():
            if not isinstance(m, EntropyBottleneck):
                continue
            rv = m.update(force=force)
            updated |= rv
        # update() is the point after which the coder must see the current parameters: drop every derived cache
        # (packed weights, packed tables, host scalars), including ones that `.data` edits could not invalidate
        _cache.invalidate_caches(self)
        return updated

    def load_state_dict(self, state_dict):
        update_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck",
                                  ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        super().load_state_dict(state_dict)
        _cache.invalidate_caches(self)


class FactorizedPrior(CompressionModel):
    r"""Factorized Prior model (Ballé et al., ICLR 2018).  Args: N, M as in the reference."""

    def __init__(self, N, M, **kwargs):
        super().__init__(entropy_bottleneck_channels=M, **kwargs)
        self.g_a = TransformStack(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = TransformStack(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.N = N
        self.M = M

    @property
    def downsampling_factor(self) -> int:
        return 2**4

    def forward(self, x):
        y = self.g_a(_nhwc(x))
        y_hat, y_likelihoods = self.entropy_bottleneck(y)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods}}

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M)
        net.load_state_dict(state_dict)
        return net

    @torch.no_grad()
    def compress(self, x):
        if x.dtype == torch.uint8:
            x = kernels.pixels_to_float(x)
        y = self.g_a(_nhwc(x))
        y_strings = self.entropy_bottleneck.compress(y)
        return {"strings": [y_strings], "shape": y.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape, memory_format=_CL)
        x_hat = self.g_s(y_hat, clamp=(0.0, 1.0), nchw_out=True)
        return {"x_hat": x_hat}


# From Balle's tensorflow compression examples
SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class ScaleHyperprior(CompressionModel):
    r"""Scale Hyperprior model (Ballé et al., ICLR 2018)."""

    def __init__(self, N, M, **kwargs):
        super().__init__(entropy_bottleneck_channels=N, **kwargs)
        self.g_a = TransformStack(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))
        self.g_s = TransformStack(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
                                 deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                 nn.ReLU(inplace=True), conv(N, N))
        self.h_s = TransformStack(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                 conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    _abs_hyper_input = True

    def _hyper_in(self, y):
        return torch.abs(y)

    def _params(self, z_hat):
        """(scales_hat, means_hat or None) from the hyper-synthesis transform."""
        return self.h_s(z_hat), None

    def forward(self, x):
        y = self.g_a(_nhwc(x))
        z = self.h_a(self._hyper_in(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat, means_hat = self._params(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat, means=means_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}

    def load_state_dict(self, state_dict):
        update_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                  ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        super().load_state_dict(state_dict)

    @classmethod
    def from_state_dict(cls, state_dict):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M)
        net.load_state_dict(state_dict)
        return net

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    # Transforms run in micro-batches (activation memory), the coder runs ONCE per tensor for the whole batch: a
    # rANS string is a serial chain, so the launch latency is that of one string however many strings it codes.
    micro_batch = 32

    def _analysis_chunk(self, x):
        """g_a + h_a + hyper-latent quantisation + h_s for one micro-batch (inference)."""
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        if self._abs_hyper_input:
            y, y_abs = self.g_a(x, want_abs=True)
            z = self.h_a(y_abs)
        else:
            y = self.g_a(x)
            z = self.h_a(y)
        z_sym, z_idx = kernels.eb_quantize_index(z, eb._get_medians())
        z_hat = kernels.dequantize(z_sym, None, eb._get_medians(), tuple(z.shape), _CL)
        scales_hat, means_hat = self._params(z_hat)
        y_sym, y_idx = kernels.gc_quantize_index(y, scales_hat, means_hat, gc.scale_table, gc._bound_scale())
        return y_sym, y_idx, z_sym, z_idx, z.size()[-2:]

    # ---- stream plan ------------------------------------------------------------------------------------------
    # "ana": analysis transforms of all chunks, in chunk order.  "syn": synthesis transforms, in chunk order.
    # "coder[k]": the serial-chain coder kernels of chunk k (high priority: they need one SM each, for long).
    # A chunk's coder latency (tens of ms, a few SMs) then overlaps the tensor-core work of the other chunks, and
    # because nothing here synchronises the host, the analysis of the NEXT call overlaps this call's decode wait.
    max_streams = 8

    coder_stream_pool = int(os.environ.get("CAI_CODER_STREAMS", "16"))

    def _streams(self, device, n):
        """Shared, in-order transform streams -- "ana" (analysis), "hyp" (hyper-synthesis + index kernels: short work
        that only ever waits for a z decode) and "syn" (g_s: waits for the long y decodes) -- so that the tensor-core
        work of all in-flight requests executes in a deterministic FIFO order and a request waiting for its decode
        never blocks another request's hyper-synthesis.  Coder streams come round-robin from a pool larger than the
        number of chunks in flight, so two requests rarely queue on the same coder stream."""
        pool = self.__dict__.setdefault("_stream_pool", {})
        st = pool.get(str(device))
        if st is None:
            # "hyp" is high priority: its kernels are tiny and gate the start of the long y decodes, so they must
            # not queue behind whole analysis / synthesis grids of other requests in the block scheduler
            st = {"ana": torch.cuda.Stream(device=device), "hyp": torch.cuda.Stream(device=device, priority=int(os.environ.get("CAI_HYP_PRIO", "-1"))),
                  "syn": torch.cuda.Stream(device=device),
                  "h2d": torch.cuda.Stream(device=device), "d2h": torch.cuda.Stream(device=device),
                  "pool": [torch.cuda.Stream(device=device, priority=-1) for _ in range(self.coder_stream_pool)],
                  "next": 0}
            pool[str(device)] = st
        return st

    def _take_coder_streams(self, st, n):
        out = []
        for _ in range(n):
            out.append(st["pool"][st["next"] % len(st["pool"])])
            st["next"] += 1
        return out

    @staticmethod
    def _handoff(stream, *tensors):
        """Tensors produced on the current stream are about to be consumed on ``stream``."""
        ev = torch.cuda.current_stream().record_event()
        stream.wait_event(ev)
        for t in tensors:
            if t is not None:
                t.record_stream(stream)

    def _check_tables(self):
        for m in (self.entropy_bottleneck, self.gaussian_conditional):
            m._check_cdf_size()
            m._check_cdf_length()
            m._check_offsets_size()

    @torch.no_grad()
    def compress_to_device(self, x):
        """compress() with the strings left in HBM: per micro-batch ``coder.EncodedBatch`` lists.  ``x`` float32 in
        [0, 1] as in the reference, or uint8 pixels (host or device), converted on the device as x / 255."""
        self._check_tables()
        host_in = not x.is_cuda
        dev = self.gaussian_conditional._quantized_cdf.device if host_in else x.device
        if host_in:
            # host images (serving path): each micro-batch is copied in on a side stream while the previous one is
            # being analysed; pinned memory makes the copies asynchronous.  The compute stays on the GPU.
            _lib.require_cuda(self.gaussian_conditional._quantized_cdf, "model buffers")
        else:
            _lib.require_cuda(x, "inputs")
        eb_t, gc_t = self.entropy_bottleneck._table(), self.gaussian_conditional._table()
        starts = list(range(0, x.size(0), self.micro_batch))
        S = self._streams(dev, len(starts))
        coder_streams = self._take_coder_streams(S, len(starts))
        ana = S["ana"]
        ana.wait_event(torch.cuda.current_stream(dev).record_event())
        if not host_in:
            x.record_stream(ana)
        y_encs, z_encs, shape = [], [], None
        for k, i in enumerate(starts):
            ck = coder_streams[k]
            if host_in:
                with torch.cuda.stream(S["h2d"]):
                    xc = x[i:i + self.micro_batch].to(dev, non_blocking=True)
                    ana.wait_event(S["h2d"].record_event())
                    xc.record_stream(ana)
            else:
                xc = x[i:i + self.micro_batch]
            with torch.cuda.stream(ana):
                if xc.dtype == torch.uint8:  # 8-bit pixels: x = u8 / 255 on the device (ToTensor convention)
                    xc = kernels.pixels_to_float(xc)
                y_sym, y_idx, z_sym, z_idx, shape = self._analysis_chunk(xc)
                self._handoff(ck, y_sym, y_idx, z_sym, z_idx)
            with torch.cuda.stream(ck):
                z_encs.append(coder.encode(eb_t, z_sym, z_idx))
                y_encs.append(coder.encode(gc_t, y_sym, y_idx))
                z_encs[-1].stream = y_encs[-1].stream = ck
        return {"strings": [y_encs, z_encs], "shape": shape}

    def compress(self, x):
        """Reference interface (models/google.py:302-312): ``{"strings": [y_strings, z_strings], "shape": ...}``.
        The two string lists are ``coder.PackedStrings``: lists of read-only zero-copy views of one pinned host
        buffer (use ``bytes(s)`` for an owning copy); ``decompress`` uploads them without repacking."""
        out = self.compress_to_device(x)
        y_encs, z_encs = out["strings"]
        host, begins = coder.batches_to_host(list(y_encs) + list(z_encs))  # two host syncs for the whole request
        ny = len(y_encs)

        def joined(bgs):  # consecutive batches occupy consecutive ranges of the packed buffer
            if not bgs:
                return coder.PackedStrings(host, [0])
            return coder.PackedStrings(host, [int(bgs[0][0])] + [int(v) for bg in bgs for v in bg[1:]])

        return {"strings": [joined(begins[:ny]), joined(begins[ny:])], "shape": out["shape"]}

    @torch.no_grad()
    def _decompress_chunks(self, chunks, shape, device, statuses=None, out=None):
        """chunks: list of (coder_stream, y_words, z_words, n) with *_words = (strings | None, device_words | None)."""
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        eb_t, gc_t = eb._table(), gc._table()
        if statuses is None:
            statuses = []  # never check inside the pipeline: that would synchronise the host per chunk
        S = self._streams(device, 1)
        syn, hyp = S["syn"], S["hyp"]
        main = torch.cuda.current_stream(device)
        ev_main = main.record_event()
        syn.wait_event(ev_main)
        hyp.wait_event(ev_main)
        C = eb._quantized_cdf.size(0)
        h, w = int(shape[0]), int(shape[1])
        # pass 1 (all chunks): z decode -> h_s -> indexes -> launch the y decode.  pass 2: dequantize + g_s.
        # "syn" is in order, so issuing every h_s before any g_s keeps one chunk's y-decode latency from
        # blocking the next chunk's hyper-synthesis.
        ready = main.record_event()
        pending = []
        for ck, y_words, z_words, n in chunks:
            with torch.cuda.stream(ck):
                ck.wait_event(ready)
                z_idx = kernels.channel_indexes(n, C, h * w, device)
                z_sym = coder.decode(eb_t, z_words[0], z_idx, device_words=z_words[1], status_out=statuses)
                self._handoff(hyp, z_sym)
            with torch.cuda.stream(hyp):
                z_hat = kernels.dequantize(z_sym, None, eb._get_medians(), (n, C, h, w), _CL)
                scales_hat, means_hat = self._params(z_hat)
                _, y_idx = kernels.gc_quantize_index(None, scales_hat, None, gc.scale_table, gc._bound_scale())
                self._handoff(ck, y_idx)
                if means_hat is not None:
                    means_hat.record_stream(syn)
            with torch.cuda.stream(ck):
                y_sym = coder.decode(gc_t, y_words[0], y_idx, device_words=y_words[1], status_out=statuses)
                done = ck.record_event()  # "syn" must NOT wait here: that would serialise the chunks' decodes
                y_sym.record_stream(syn)
            pending.append((y_sym, means_hat, tuple(scales_hat.shape), done))
        x_hat = None
        with torch.cuda.stream(syn):
            syn.wait_event(hyp.record_event())  # means_hat / shapes produced on "hyp"
            row = 0
            if out is None and len(pending) > 1:
                # device result: every micro-batch's last layer writes straight into its rows of ONE batch tensor
                # (a torch.cat of the per-chunk outputs re-copied 2.4 GB per 256-image step)
                total = sum(p[2][0] for p in pending)
                hy, wy = pending[0][2][2], pending[0][2][3]  # g_s upsamples the latent grid by 16
                x_hat = torch.empty((total, 3, 16 * hy, 16 * wy), dtype=torch.float32, device=device)
            for y_sym, means_hat, shp, done in pending:
                syn.wait_event(done)
                y_hat = kernels.dequantize(y_sym, means_hat, None, shp, _CL)
                dst = x_hat[row:row + shp[0]] if x_hat is not None else None
                xc = self.g_s(y_hat, clamp=(0.0, 1.0), nchw_out=True, out=dst)
                if dst is not None and xc.data_ptr() != dst.data_ptr():
                    dst.copy_(xc)  # last layer without the direct-destination path (more than 4 output channels)
                if out is None:
                    if x_hat is None:
                        x_hat = xc
                    row += shp[0]
                else:
                    # host output buffer (serving path): this micro-batch goes home while the next one is synthesised
                    d2h = S["d2h"]
                    if out.dtype == torch.uint8:  # 8-bit result: round(x_hat * 255) on the device, 1/4 of the bytes
                        xc = kernels.pixels_to_u8(xc)
                    d2h.wait_event(syn.record_event())
                    with torch.cuda.stream(d2h):
                        out[row:row + xc.size(0)].copy_(xc, non_blocking=True)
                    xc.record_stream(d2h)
                    row += xc.size(0)
        if out is not None:
            main.wait_event(S["d2h"].record_event())
            return {"x_hat": out, "status": statuses}
        main.wait_event(syn.record_event())
        x_hat.record_stream(main)
        return {"x_hat": x_hat, "status": statuses}

    @torch.no_grad()
    def decompress(self, strings, shape, out=None):
        """``out``: optional preallocated CPU tensor [B, 3, H, W] (pinned for asynchronous copies) that receives the
        reconstruction micro-batch by micro-batch; the returned ``x_hat`` is then ``out`` itself.  A uint8 ``out``
        receives round(x_hat * 255), converted on the device (a quarter of the PCIe bytes)."""
        assert isinstance(strings, list) and len(strings) == 2
        self._check_tables()
        if len(strings[0]) != len(strings[1]):
            raise ValueError("Invalid strings or indexes parameters")
        dev = self.gaussian_conditional._quantized_cdf.device
        _lib.require_cuda(self.gaussian_conditional._quantized_cdf, "model buffers")
        n = len(strings[0])
        starts = list(range(0, n, self.micro_batch))
        S = self._streams(dev, len(starts))
        coder_streams = self._take_coder_streams(S, len(starts))
        chunks = []
        for k, i in enumerate(starts):
            ys, zs = strings[0][i:i + self.micro_batch], strings[1][i:i + self.micro_batch]
            ys, zs = (v if isinstance(v, coder.PackedStrings) else list(v) for v in (ys, zs))
            chunks.append((coder_streams[k], (ys, None), (zs, None), len(ys)))
        statuses = []
        if out is not None and (out.is_cuda or out.dim() != 4 or out.size(0) != n
                                or out.dtype not in (torch.float32, torch.uint8)):
            raise ValueError("out must be a float32 or uint8 CPU tensor [len(strings[0]), 3, H, W]")
        res = self._decompress_chunks(chunks, shape, dev, statuses, out)
        coder.wait_stream(torch.cuda.current_stream(dev))  # one (non-spinning) sync per call: surface decoder errors
        coder.check_status(statuses)
        return {"x_hat": res["x_hat"]}

    @torch.no_grad()
    def decompress_from_device(self, enc, shape):
        """decompress() of strings still in HBM (``compress_to_device``).  Fully asynchronous; the per-string decoder
        statuses come back under ``"status"`` (device tensors) for ``coder.check_status`` once the caller has
        synchronised -- nothing is checked here."""
        y_encs, z_encs = enc
        dev = y_encs[0].slots.device
        chunks, statuses = [], []
        for ye, ze in zip(y_encs, z_encs):
            with torch.cuda.stream(ye.stream):  # the word offsets are derived from n_words on the coder stream
                chunks.append((ye.stream, (None, ye.device_words()), (None, ze.device_words()), int(ye.n_words.numel())))
        for ye, ze in zip(y_encs, z_encs):  # encoder statuses (slot overflow, bad index) travel with the result too
            statuses += [ye.status, ze.status]
        return self._decompress_chunks(chunks, shape, dev, statuses)


class MeanScaleHyperprior(ScaleHyperprior):
    r"""Scale Hyperprior with non zero-mean Gaussian conditionals (Minnen et al., NeurIPS 2018)."""

    def __init__(self, N, M, **kwargs):
        super().__init__(N, M, **kwargs)
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True), conv(N, N),
                                 nn.LeakyReLU(inplace=True), conv(N, N))
        self.h_s = TransformStack(deconv(N, M), nn.LeakyReLU(inplace=True), deconv(M, M * 3 // 2),
                                 nn.LeakyReLU(inplace=True), conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))

    _abs_hyper_input = False

    def _hyper_in(self, y):
        return y

    def _params(self, z_hat):
        gaussian_params = self.h_s(z_hat)
        scales_hat, means_hat = gaussian_params.chunk(2, 1)
        return scales_hat, means_hat


class JointAutoregressiveHierarchicalPriors(MeanScaleHyperprior):
    r"""Joint autoregressive + hierarchical priors model (Minnen et al., NeurIPS 2018); reference
    models/google.py:395-661.  Same submodules and ``state_dict`` keys (``context_prediction`` is a ``MaskedConv2d``,
    ``entropy_parameters`` three 1x1 convolutions).

    ``compress`` / ``decompress``: the reference walks the latent grid pixel by pixel in Python on the CPU
    (``_compress_ar`` :535-577, ``_decompress_ar`` :620-661).  Here the whole scan of a batch is ONE persistent kernel
    launch (csrc/ar.cu: a thread-block cluster per image group, weights streamed from L2, the image's rANS chain decoded
    in the loop); the encoder's symbols / indexes are then coded by the batched rANS kernel, one string per image, in
    the reference's order (pixel-major, channel-minor), so the strings have the reference's format."""

    def __init__(self, N=192, M=192, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.g_a = TransformStack(conv(3, N, kernel_size=5, stride=2), GDN(N), conv(N, N, kernel_size=5, stride=2), GDN(N),
                                 conv(N, N, kernel_size=5, stride=2), GDN(N), conv(N, M, kernel_size=5, stride=2))
        self.g_s = TransformStack(deconv(M, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                                 deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                                 deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                                 deconv(N, 3, kernel_size=5, stride=2))
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True),
                                 conv(N, N, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 conv(N, N, stride=2, kernel_size=5))
        self.h_s = TransformStack(deconv(N, M, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 deconv(M, M * 3 // 2, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                 conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))
        self.entropy_parameters = TransformStack(
            Conv2d(M * 12 // 3, M * 10 // 3, kernel_size=1, stride=1, padding=0), nn.LeakyReLU(inplace=True),
            Conv2d(M * 10 // 3, M * 8 // 3, kernel_size=1, stride=1, padding=0), nn.LeakyReLU(inplace=True),
            Conv2d(M * 8 // 3, M * 6 // 3, kernel_size=1, stride=1, padding=0))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def forward(self, x):
        y = self.g_a(_nhwc(x))
        z = self.h_a(y)
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(y, "noise" if self.training else "dequantize")
        ctx_params = self.context_prediction(y_hat)
        gaussian_params = self.entropy_parameters(torch.cat((params, ctx_params), dim=1))
        scales_hat, means_hat = gaussian_params.chunk(2, 1)
        _, y_likelihoods = self.gaussian_conditional(y, scales_hat, means=means_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}

    # launch shape of the scan (0 = automatic): CTAs per cluster, images per cluster
    ar_cluster = 0
    ar_group = 0

    def _ar_weights(self) -> kernels.ArWeights:
        cp, ep = self.context_prediction, self.entropy_parameters
        convs = [ep[0], ep[2], ep[4]]
        tensors = [cp.weight, cp.bias, cp.mask] + [t for c in convs for t in (c.weight, c.bias)]

        def build():
            return kernels.ArWeights(cp.weight.detach() * cp.mask, cp.bias, [(c.weight, c.bias) for c in convs],
                                     slope=float(ep[1].negative_slope))

        return _cache.cached(self, "ar_weights", _cache.tensor_key(*tensors), build, cp.weight.device)

    @staticmethod
    def _to_nhwc(t):
        return t.permute(0, 2, 3, 1).contiguous()

    @torch.no_grad()
    def compress(self, x):
        self._check_tables()
        if not x.is_cuda:
            x = x.to(self.gaussian_conditional._quantized_cdf.device)
        if x.dtype == torch.uint8:
            x = kernels.pixels_to_float(x)
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        y = self.g_a(_nhwc(x))
        z = self.h_a(y)
        z_sym, z_idx = kernels.eb_quantize_index(z, eb._get_medians())
        z_enc = coder.encode(eb._table(), z_sym, z_idx)
        z_hat = kernels.dequantize(z_sym, None, eb._get_medians(), tuple(z.shape), _CL)
        params = self.h_s(z_hat)
        if params.shape[-2:] != y.shape[-2:]:
            raise ValueError("the image size must be a multiple of 64 (latent and hyper-synthesis grids differ)")
        sym, idx, _ = kernels.ar_encode(self._ar_weights(), self._to_nhwc(y), self._to_nhwc(params), gc.scale_table,
                                        gc._bound_scale(), self.ar_cluster, self.ar_group)
        y_enc = coder.encode(gc._table(), sym, idx)
        y_strings, z_strings = coder.batches_to_bytes([y_enc, z_enc])
        coder.check_status([y_enc.status, z_enc.status], "rANS encode")
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        self._check_tables()
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        dev = gc._quantized_cdf.device
        _lib.require_cuda(gc._quantized_cdf, "model buffers")
        z_hat = eb.decompress(strings[1], shape, memory_format=_CL)
        params = self.h_s(z_hat)
        words, wb, keep = coder.strings_to_device(coder.as_strings(strings[0]), dev)
        y_hat, status, _ = kernels.ar_decode(self._ar_weights(), gc._table(), words, wb, self._to_nhwc(params),
                                             gc.scale_table, gc._bound_scale(), self.ar_cluster, self.ar_group)
        coder.check_status([status], "rANS decode")
        p = self.context_prediction.kernel_size // 2
        y_hat = y_hat[:, p:-p, p:-p, :].permute(0, 3, 1, 2)
        x_hat = self.g_s(y_hat, clamp=(0.0, 1.0), nchw_out=True)
        return {"x_hat": x_hat}
