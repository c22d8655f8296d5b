#!/bin/bash
# Rebuilds with different numbers of independent TMEM accumulators per tile (round-robin over the K steps of a stage)
# and times small-block / split-K launches L2-hot and L2-cold (GPU box).  "1 1 1 0" is the single-accumulator kernel.
set -e
cd "$(dirname "$0")/.."
for cfg in "1 1 1 0" "2 4 4 1" "2 2 2 1" "2 4 4 0" "2 2 4 1"; do
  set -- $cfg
  SMT_NVCC_EXTRA="-DSMT_GEMM_NACC_256=$1 -DSMT_GEMM_NACC_128=$2 -DSMT_GEMM_NACC_64=$3 -DSMT_GEMM_B64_DEEP=$4" python sparse_matrix_tuning_b200/build.py --force > /dev/null
  echo "== accumulators: half-block b=256 tile $1, b=128 $2, b=64 $3; b=64 deep pipeline $4"
  python tools/l2_resident_probe.py 2>&1
  python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -1
done
python sparse_matrix_tuning_b200/build.py --force > /dev/null
