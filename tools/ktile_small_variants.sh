#!/bin/bash
# Stage height (tokens) for the small tiles, with / without a shared-memory slot for the unused half of the b = 64
# M=128 operand (GPU box).  Prints the L2-hot / L2-cold probe and runs the GEMM parity tests for every variant.
set -e
cd "$(dirname "$0")/.."
for cfg in "128 128 0" "128 128 1" "128 256 1" "128 256 0" "128 64 0" "64 128 0"; do
  set -- $cfg
  SMT_NVCC_EXTRA="-DSMT_GEMM_KTILE_128=$1 -DSMT_GEMM_KTILE_64=$2 -DSMT_GEMM_B64_ALIAS=$3" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== tokens per stage: b=128 $1, b=64 $2; b=64 operand alias $3"
  python tools/l2_resident_probe.py 2>&1 | grep -v "b=256"
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -q -x -k "gemm" 2>&1 | tail -1
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
