"""In-situ drop-in proof (plan item 1): the reference's OWN unit tests of the coder boundary --
tests/test_entropy_models.py (incl. :258-283 compress / decompress round trips), tests/test_ops.py (incl. :104-118
pmf_to_quantized_cdf KAT and error contract) and tests/test_coder.py -- run unmodified from oracle/_ref/tests (copied
there by oracle/build_ref.sh) against the unmodified reference package, with only ``compressai.ans`` and
``compressai._CXX`` replaced by this repo's GPU-backed modules (tests/insitu_plugin.py).

The two ``test_update`` cases need pretrained weights from the network and fail with the reference's own extensions
too; they are deselected, everything else must pass (44 tests here with the reference's extensions)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def test_reference_tests_pass_on_our_coder():
    tdir = os.path.join(REF, "tests")
    files = [os.path.join(tdir, f) for f in ("test_entropy_models.py", "test_ops.py", "test_coder.py")]
    if not all(os.path.exists(f) for f in files):
        pytest.skip("oracle/_ref/tests missing (built by oracle/build_ref.sh where /root/reference is mounted)")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, os.path.join(ROOT, "tests")])
    names = [os.path.basename(f) for f in files]  # node ids are relative to the working directory (tdir)
    cmd = [sys.executable, "-m", "pytest", *names, "-q", "-p", "insitu_plugin", "-p", "no:cacheprovider",
           "--deselect", names[0] + "::TestEntropyBottleneck::test_update",
           "--deselect", names[0] + "::TestGaussianConditional::test_update"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=tdir, timeout=900)
    tail = out.stdout[-4000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) >= 44, tail
    assert "failed" not in out.stdout.splitlines()[-1], tail
    calls = dict(kv.split("=") for kv in re.search(r"insitu: calls (.*)", out.stdout).group(1).split())
    assert int(calls["encode_with_indexes"]) > 0 and int(calls["decode_with_indexes"]) > 0
    assert int(calls["pmf_to_quantized_cdf"]) > 0, calls
    assert "compressai_environment_b200" in re.search(r"insitu: compressai.ans -> (.*)", out.stdout).group(1)
