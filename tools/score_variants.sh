#!/bin/bash
# Rebuilds the library with different score_accumulate unroll / grid settings and times the HBM kernels (GPU box).
set -e
cd "$(dirname "$0")/.."
for cfg in "1 8" "2 8" "4 8" "2 4" "2 16" "4 4"; do
  set -- $cfg
  SMT_NVCC_EXTRA="-DSMT_SCORE_UNROLL=$1 -DSMT_SCORE_CTAS=$2" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== SMT_SCORE_UNROLL=$1 SMT_SCORE_CTAS=$2"
  python tools/profile_kernels.py hbm 2>&1 | grep -E "score_accumulate|block_sum|block_score"
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
