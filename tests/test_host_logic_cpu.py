"""CPU suite: host-side pieces around the kernels that need no GPU -- the packed string container the codec API
returns, the per-rank CPU placement plan of bench.py, and the gated-gradient / reparametrisation modules whose
behaviour the reference's checkpoints rely on (compressai/ops/bound_ops.py:36-80, ops/parametrizers.py:38-64)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_packed_strings_views_and_slices():
    from compressai_environment_b200 import coder

    host = torch.arange(20, dtype=torch.int32)
    raw = host.numpy().tobytes()
    ps = coder.PackedStrings(host, [0, 3, 3, 10, 20])
    assert len(ps) == 4 and [len(s) for s in ps] == [12, 0, 28, 40]
    assert ps[0] == raw[:12] and ps[2] == raw[12:40] and bytes(ps[3]) == raw[40:]
    assert b"".join(ps) == raw
    assert ps.intact()
    sub = ps[1:3]
    assert isinstance(sub, coder.PackedStrings) and sub.intact() and list(sub.begin) == [3, 3, 10]
    assert sub[1] == raw[12:40]
    assert isinstance(ps[::2], list) and not isinstance(ps[::2], coder.PackedStrings)
    assert ps.to_bytes() == [raw[:12], b"", raw[12:40], raw[40:]]
    try:
        ps[0][0:1] = b"x"
        raise AssertionError("views must be read-only")
    except TypeError:
        pass
    ps.append(b"zz")  # an edited list no longer mirrors the buffer: decode must fall back to the generic path
    assert not ps.intact()
    empty = coder.PackedStrings(host, [5])
    assert len(empty) == 0 and empty.intact()


def test_strings_to_device_generic_path_on_cpu():
    """The generic packer (arbitrary bytes objects, lengths not multiples of 4) zero-pads to whole words."""
    from compressai_environment_b200 import coder

    words, wb, _ = coder.strings_to_device([b"\x01\x02\x03\x04\x05", b"", b"\xff" * 8], torch.device("cpu"))
    assert wb.tolist() == [0, 2, 2, 4]
    assert words.numpy().view(np.uint8).tolist() == [1, 2, 3, 4, 5, 0, 0, 0] + [255] * 8


def test_plan_rank_cpus():
    import bench

    allowed = set(range(32))
    numa0, numa1 = set(range(0, 16)), set(range(16, 32))
    sets = [numa0] * 4 + [numa1] * 4
    plans = [bench.plan_rank_cpus(allowed, sets, r) for r in range(8)]
    assert all(len(p) == 4 for p in plans)
    assert set().union(*plans) == allowed and sum(len(p) for p in plans) == 32  # disjoint cover
    assert all(p <= numa0 for p in plans[:4]) and all(p <= numa1 for p in plans[4:])
    # no topology information: even split of the allowed CPUs
    plans = [bench.plan_rank_cpus(allowed, [None] * 8, r) for r in range(8)]
    assert sum(len(p) for p in plans) == 32 and len(set().union(*plans)) == 32
    # cgroup smaller than the node lists: only allowed CPUs are handed out
    plans = [bench.plan_rank_cpus(set(range(8)), [numa0, numa0], r) for r in range(2)]
    assert plans[0] | plans[1] == set(range(8)) and not (plans[0] & plans[1])
    assert bench.plan_rank_cpus(allowed, [numa0], 0) == allowed  # one rank keeps everything
    assert bench._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}


def test_lower_bound_gradient_gate():
    from compressai_environment_b200.ops import LowerBound

    lb = LowerBound(0.5)
    x = torch.tensor([0.25, 0.75, 0.375, 0.5], requires_grad=True)
    y = lb(x)
    assert y.tolist() == [0.5, 0.75, 0.5, 0.5]
    y.backward(torch.tensor([1.0, 1.0, -1.0, 2.0]))
    # below the bound: only gradients that push the value up (negative) pass; at / above the bound everything passes
    assert x.grad.tolist() == [0.0, 1.0, -1.0, 2.0]
    assert "bound" in dict(lb.named_buffers())


def test_non_negative_parametrizer_roundtrip():
    from compressai_environment_b200.ops import NonNegativeParametrizer

    p = NonNegativeParametrizer(minimum=1e-6)
    v = torch.tensor([1.0, 0.0, 2.5, 1e-7])
    out = p(p.init(v))
    assert torch.allclose(out, torch.tensor([1.0, 1e-6, 2.5, 1e-6]), rtol=1e-5, atol=1e-9)
    assert set(dict(p.named_buffers())) == {"pedestal", "lower_bound.bound"}
    ped = 2.0 ** -36
    assert abs(float(p.pedestal) - ped) < 1e-18 and abs(float(p.lower_bound.bound) - (1e-6 + ped) ** 0.5) < 1e-9


def test_cache_dropped_by_train_apply_and_pickle():
    import copy

    from compressai_environment_b200 import _cache
    from compressai_environment_b200.models import ScaleHyperprior

    net = ScaleHyperprior(16, 16)
    conv = net.g_a[0]
    calls = []
    v = _cache.cached(conv, "probe", ("k",), lambda: calls.append(1) or "built")
    assert v == "built" and _cache.cached(conv, "probe", ("k",), lambda: calls.append(1) or "again") == "built"
    assert _cache.cached(conv, "probe", ("k2",), lambda: "rebuilt") == "rebuilt"          # key change
    net.eval()
    assert _cache.cached(conv, "probe", ("k2",), lambda: "after-eval") == "after-eval"      # train()/eval() drops
    net.float()
    assert _cache.cached(conv, "probe", ("k2",), lambda: "after-apply") == "after-apply"    # _apply drops
    net.load_state_dict(net.state_dict())
    assert _cache.cached(conv, "probe", ("k2",), lambda: "after-load") == "after-load"      # load_state_dict drops
    clone = copy.deepcopy(net)
    assert "_cai_cache" not in clone.g_a[0].__dict__ or not clone.g_a[0].__dict__["_cai_cache"]
