#!/bin/bash
# Rebuilds the library with different Adam register/vector settings and times the HBM kernels (run on the GPU box).
set -e
cd "$(dirname "$0")/.."
for cfg in "8 4" "4 5" "4 6" "4 8" "8 3"; do
  set -- $cfg
  SMT_NVCC_EXTRA="-DSMT_ADAM_VEC=$1 -DSMT_ADAM_MINB=$2" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== SMT_ADAM_VEC=$1 SMT_ADAM_MINB=$2"
  python tools/profile_kernels.py hbm 2>&1 | grep compact_adam
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
