"""Runner for tests/test_insitu_train_gpu.py: executes the reference's UNMODIFIED training example
(oracle/_ref/examples/train.py, copied there by oracle/build_ref.sh) with ``compressai`` resolving to THIS repo's
package: ``compressai.zoo.image_models`` -> our models (every forward / backward GEMM on our kernels); the data loading
helper ``compressai.datasets.ImageFolder`` is host-side file handling outside the hot path and is loaded from the
reference's own pure-Python source.  Usage: python insitu_train.py <argv of examples/train.py ...>"""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, ROOT)

import compressai_environment_b200 as ours  # noqa: E402
from compressai_environment_b200 import zoo as our_zoo  # noqa: E402

sys.modules["compressai"] = ours
sys.modules["compressai.zoo"] = our_zoo
spec = importlib.util.spec_from_file_location("compressai.datasets.image", os.path.join(REF, "compressai", "datasets", "image.py"))
image_mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(image_mod)
datasets = types.ModuleType("compressai.datasets")
datasets.ImageFolder = image_mod.ImageFolder
sys.modules["compressai.datasets"] = datasets

spec = importlib.util.spec_from_file_location("examples.train", os.path.join(REF, "examples", "train.py"))
train = importlib.util.module_from_spec(spec)
spec.loader.exec_module(train)
assert train.image_models is our_zoo.image_models
train.main(sys.argv[1:])
