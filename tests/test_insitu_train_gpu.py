"""In-situ training parity: the reference's OWN training test protocol
(tests/test_train.py:40-88: examples/train.py on the fake image folder, 10 epochs, batch 1, patch 48x128, seed 3.14,
bmshj2018-factorized q3) run UNMODIFIED against this repo's models on the GPU, compared with the reference's committed
expected log (tests/expected/train_log_3.14.txt, produced by the reference on CPU).

The reference's assertion is: same number of numeric tokens, integer tokens equal.  That is asserted as is.  On top of
it every loss value of all 20 log lines is compared numerically.  The run sets CAI_NOISE_RNG=cpu so that the
quantisation noise is drawn from torch's CPU generator in the reference's element order (entropy_models._training_noise):
the seeded run then consumes the same random stream as the reference (same crops, same noise) and the only difference
left is floating point.  Bar: every value within ONE unit of its last printed digit (measured on B200: 19 of 20 loss
values identical to the printed digit, one off by 0.001; every MSE / bpp / aux value identical)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
NUM = r"(?P<number>([0-9]*[.])?[0-9]+)"


def test_reference_training_example_on_our_kernels(tmp_path):
    script = os.path.join(REF, "examples", "train.py")
    data = os.path.join(REF, "tests", "assets", "fakedata", "imagefolder")
    expected_path = os.path.join(REF, "tests", "expected", "train_log_3.14.txt")
    if not (os.path.exists(script) and os.path.isdir(data) and os.path.exists(expected_path)):
        pytest.skip("oracle/_ref/examples missing (built by oracle/build_ref.sh where /root/reference is mounted)")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0],
               CAI_NOISE_RNG="cpu")  # draw the quantisation noise from the CPU generator in the reference's order
    cmd = [sys.executable, os.path.join(ROOT, "tests", "insitu_train.py"), "-d", data, "-e", "10", "--batch-size", "1",
           "--patch-size", "48", "128", "--seed", "3.14", "--cuda"]
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    log, expected = out.stdout, open(expected_path).read()
    got = [m[0] for m in re.findall(NUM, log)]
    want = [m[0] for m in re.findall(NUM, expected)]
    assert len(got) == len(want), (len(got), len(want), log[-1500:])
    for a, b in zip(got, want):          # the reference's own check (tests/test_train.py:80-88)
        try:
            assert int(a) == int(b)
        except ValueError:
            pass
    # numeric comparison line by line
    pat = re.compile(r"Loss: ([0-9.]+) \|\s*MSE loss: ([0-9.]+) \|\s*Bpp loss: ([0-9.]+) \|\s*Aux loss: ([0-9.]+)")
    g_lines, w_lines = pat.findall(log), pat.findall(expected)
    assert len(g_lines) == len(w_lines) == 20
    worst = [0.0, 0.0, 0.0, 0.0]
    for g, w in zip(g_lines, w_lines):
        g, w = [float(v) for v in g], [float(v) for v in w]
        worst = [max(worst[0], abs(g[0] - w[0]) / w[0]), max(worst[1], abs(g[1] - w[1])), max(worst[2], abs(g[2] - w[2])),
                 max(worst[3], abs(g[3] - w[3]) / w[3])]
    print("in-situ training: worst deviations (loss rel, mse abs, bpp abs, aux rel):", worst)
    assert worst[0] <= 1e-4 and worst[1] <= 0.0011 and worst[2] <= 0.011 and worst[3] <= 1e-5, (worst, log[-1500:])
