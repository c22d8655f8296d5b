#!/bin/bash
# Rebuilds with different K-tile sizes / stage caps and times the GEMM cases (GPU box).
set -e
cd "$(dirname "$0")/.."
for cfg in "64 8" "32 8" "32 12" "64 12" "16 12"; do
  set -- $cfg
  SMT_NVCC_EXTRA="-DSMT_GEMM_KTILE=$1 -DSMT_GEMM_MAX_STAGES=$2" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== SMT_GEMM_KTILE=$1 SMT_GEMM_MAX_STAGES=$2"
  python tools/profile_kernels.py gemm 2>&1
  python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -1
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
