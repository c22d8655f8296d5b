#!/bin/bash
# On the GPU box: time the fused dense kernel variants built by tools/dense_variants.sh and a sweep of raster groups.
cd "$(dirname "$0")/.."
for st in 4 5 6; do
  echo "== stages $st"
  SMT_B200_LIB=$PWD/build/variants/libsmt_stages$st.so timeout 200 python tools/fused_qkv_bench.py 2>&1 | sed -n 3,4p
done
for g in 12 16 24; do
  echo "== group_m $g"
  SMT_DENSE_GROUP_M=$g timeout 200 python tools/fused_qkv_bench.py 2>&1 | sed -n 3,4p
done
