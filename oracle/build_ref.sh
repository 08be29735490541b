#!/usr/bin/env bash
# Build the UNMODIFIED reference (CompressAI 1.2.0.dev0) into oracle/_ref/ so that tests and
# bench.py --impl reference can run the reference's own CPU implementation of the hot path.
# Recipe: SURVEY.md §8(c). Outputs go ONLY to oracle/_ref/ (git-ignored, travels with gpurun).
# The reference's own build system is not run: two g++ lines on its source files where they lie.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CAI_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/compressai" ]; then
  echo "build_ref: $REF not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
PY="${PYTHON:-python}"
INC_PY="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
INC_PB="$($PY -c 'import pybind11;print(pybind11.get_include())')"
SUF="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
rm -rf "$OUT"
mkdir -p "$OUT"
# "install" of the pure-python package (what `pip install --target` would do)
cp -r "$REF/compressai" "$OUT/compressai"
rm -rf "$OUT/compressai/cpp_exts"
find "$OUT" -name '__pycache__' -type d -prune -exec rm -rf {} +
printf '__version__ = "1.2.0.dev0"\n' > "$OUT/compressai/version.py"
CXXFLAGS="-O3 -DNDEBUG -std=c++17 -shared -fPIC -fvisibility=hidden"
g++ $CXXFLAGS -I"$INC_PY" -I"$INC_PB" -I"$REF/third_party/ryg_rans" -I"$REF/compressai/cpp_exts/rans" \
    "$REF/compressai/cpp_exts/rans/rans_interface.cpp" -o "$OUT/compressai/ans$SUF"
g++ $CXXFLAGS -I"$INC_PY" -I"$INC_PB" \
    "$REF/compressai/cpp_exts/ops/ops.cpp" -o "$OUT/compressai/_CXX$SUF"
# The reference's own unit tests of this path, kept beside the install (build output, never tracked) so that
# tests/test_insitu_gpu.py can run them UNMODIFIED on the GPU box with compressai.ans / compressai._CXX replaced
# by this repo's modules (plan item 1).
mkdir -p "$OUT/tests"
for t in test_entropy_models.py test_ops.py test_coder.py; do
  cp "$REF/tests/$t" "$OUT/tests/$t"
done
# The reference's training example, its fake dataset and the expected log of its own training test
# (tests/test_train.py:40-88), for tests/test_insitu_train_gpu.py: the UNMODIFIED script drives this repo's models.
mkdir -p "$OUT/examples" "$OUT/tests/assets/fakedata" "$OUT/tests/expected"
cp "$REF/examples/train.py" "$OUT/examples/train.py"
cp -r "$REF/tests/assets/fakedata/imagefolder" "$OUT/tests/assets/fakedata/imagefolder"
cp "$REF/tests/expected/train_log_3.14.txt" "$OUT/tests/expected/train_log_3.14.txt"
echo "build_ref: reference installed in $OUT"
