#!/bin/bash
# On the GPU box: the in-step grouped block-gradient launch with the per-tile cta_group::2 kernel (SMT_GEMM_2SM=1) vs the
# persistent form (SMT_GEMM_2SM=2; only with profiles/r02_2sm_persistent_pair_tile.patch applied), alternating, CUDA-event
# times from bench.py.  Experiment tooling (results: profiles/r02_ncu_summary.md).
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 8 --warmup 3 --no-extra --no-cpu-baseline --capture-steps 1"
for i in 1 2; do
  for m in 1 2; do
    echo -n "SMT_GEMM_2SM=$m: "
    SMT_GEMM_2SM=$m $CMD 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('avg_launch_us', round(r['avg_launch_us'],1), 'TFLOP/s', round(r['achieved'],1), 'ms/step', round(d['ms_per_step'],2), 'adam frac', round(r['also']['compact_adam']['frac'],3), r['also']['compact_adam']['clip_norm_source'])"
  done
done
