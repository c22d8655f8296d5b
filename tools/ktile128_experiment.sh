#!/bin/bash
# Is the ~560 clk floor of the small tiles a per-STAGE cost (barrier round trip / commit) or a per-box cost?
# 128-token stages halve the number of stages (and of TMA boxes: one box is then 128 rows x 128 B) per tile.
# Only the b = 64 / 128 lines are meaningful: b = 256 tiles get 1-2 stages at this size.
set -e
cd "$(dirname "$0")/.."
for kt in 64 128; do
  SMT_NVCC_EXTRA="-DSMT_GEMM_KTILE=$kt" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== SMT_GEMM_KTILE=$kt"
  python tools/l2_resident_probe.py 2>&1
  python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -1
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
