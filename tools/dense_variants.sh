#!/bin/bash
# Builds variants of the fused dense kernel (pipeline depth) into build/variants/ (run HERE, no GPU needed), then
# `tools/dense_variants_run.sh` times them on the GPU box.  Experiment tooling.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for st in 4 5 6; do
  SMT_NVCC_EXTRA="-DSMT_DENSE_STAGES=$st" python - <<PY
import os, shutil
from sparse_matrix_tuning_b200 import build
p = build.build(force=True)
shutil.copy(p, "build/variants/libsmt_stages$st.so")
PY
done
python -c "from sparse_matrix_tuning_b200 import build; build.build(force=True)"
