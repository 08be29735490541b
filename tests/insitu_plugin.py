"""pytest plugin for the in-situ drop-in test (tests/test_insitu_gpu.py).

Loaded with ``-p insitu_plugin`` in a pytest process whose ``sys.path`` holds the UNMODIFIED reference package
(oracle/_ref).  Before the reference is imported it seeds ``sys.modules`` so that the reference's own
``from compressai import ans`` (compressai/entropy_models/entropy_models.py:60) and
``from compressai._CXX import pmf_to_quantized_cdf`` (:41) resolve to THIS repo's GPU-backed modules -- the binding a
maintainer of the reference would ship.  Everything else (entropy models, models, zoo, the tests
themselves) stays the reference's code.  At session end it reports how often the replaced entry points were called.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.append(ROOT)  # after oracle/_ref: `compressai` must stay the reference package

from compressai_environment_b200 import _CXX as our_cxx  # noqa: E402
from compressai_environment_b200 import ans as our_ans  # noqa: E402

CALLS = {"encode_with_indexes": 0, "decode_with_indexes": 0, "pmf_to_quantized_cdf": 0}


def _counted(fn, key):
    def wrapper(*a, **kw):
        CALLS[key] += 1
        return fn(*a, **kw)

    wrapper.__name__ = getattr(fn, "__name__", key)
    return wrapper


our_ans.RansEncoder.encode_with_indexes = _counted(our_ans.RansEncoder.encode_with_indexes, "encode_with_indexes")
our_ans.RansDecoder.decode_with_indexes = _counted(our_ans.RansDecoder.decode_with_indexes, "decode_with_indexes")
our_cxx.pmf_to_quantized_cdf = _counted(our_cxx.pmf_to_quantized_cdf, "pmf_to_quantized_cdf")
sys.modules["compressai.ans"] = our_ans
sys.modules["compressai._CXX"] = our_cxx


def pytest_sessionstart(session):
    import compressai

    assert os.path.join("oracle", "_ref") in os.path.abspath(compressai.__file__), compressai.__file__
    import compressai.entropy_models.entropy_models as em

    assert sys.modules["compressai.ans"] is our_ans and sys.modules["compressai._CXX"] is our_cxx
    assert getattr(em, "_pmf_to_quantized_cdf", None) is our_cxx.pmf_to_quantized_cdf


def pytest_terminal_summary(terminalreporter):
    terminalreporter.write_line("insitu: compressai.ans -> " + os.path.abspath(our_ans.__file__))
    terminalreporter.write_line("insitu: calls " + " ".join(f"{k}={v}" for k, v in CALLS.items()))
