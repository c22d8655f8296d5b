#!/bin/bash
# Perf-only experiment: how does the main loop speed react when fewer operand bytes are fetched per stage?
# (results are numerically wrong with SKIP > 0; this only locates the bottleneck: operand feed vs UMMA issue)
set -e
cd "$(dirname "$0")/.."
for skip in 0 2 3; do
  if [ "$skip" = "0" ]; then extra=""; else extra="-DSMT_GEMM_EXPERIMENT_SKIP_CHUNKS=$skip"; fi
  SMT_NVCC_EXTRA="$extra" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== A chunks skipped per stage: $skip (of 4 for whole-block tiles, 2 for half-block)"
  python tools/profile_kernels.py gemm 2>&1 | grep -E "n=148|n=869|n=31 "
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
