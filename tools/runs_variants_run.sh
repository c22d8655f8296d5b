#!/bin/bash
# On the GPU box: b = 64 / 128 rows of the config-2 sweep with the run-tile kernel built for different stage heights
# (libraries under build/variants, built by hand with SMT_NVCC_EXTRA="-DSMT_RUNS_KT_64=<tokens> -DSMT_RUNS_KT_128=<tokens>").
cd "$(dirname "$0")/.."
for v in "" build/variants/libsmt_runs_96_80.so; do
  echo "== ${v:-in-tree build}"
  SMT_B200_LIB=${v:+$PWD/$v} timeout 400 python tools/kernel_sweep.py --quick 2>&1 | grep "| 64 |\|| 128 |" | grep "5%\|2%"
done
