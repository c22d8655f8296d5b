#!/bin/bash
# On the GPU box: quick config-2 sweep with the run-tile kernel built for 64- and 128-token stages (tools: build/variants).
cd "$(dirname "$0")/.."
for kt in 64 128; do
  echo "== run tiles with $kt-token stages (b = 64 / 128)"
  SMT_B200_LIB=$PWD/build/variants/libsmt_runs_kt$kt.so timeout 400 python tools/kernel_sweep.py --quick 2>&1 | grep -v "| 256 |"
done
