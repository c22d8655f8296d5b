#!/bin/bash
# Rebuilds with 8 / 12 / 16 epilogue warps and times the GEMM cases (GPU box).
set -e
cd "$(dirname "$0")/.."
for w in 8 12 16; do
  SMT_NVCC_EXTRA="-DSMT_GEMM_EPI_WARPS=$w" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== SMT_GEMM_EPI_WARPS=$w"
  python tools/profile_kernels.py gemm 2>&1
  python tools/trace_gemm.py 256,8192,9 256,8192,148 2>&1 | grep -E "##|acc_done|epi_done|reduced"
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
