#!/bin/bash
# On the GPU box: in-step block-gradient GEMM with L2 eviction hints on its TMA loads (SMT_GEMM_L2_HINT = 0..3; the
# switch only exists with profiles/r02_l2_hint_experiment.patch applied): CUDA-event time from bench.py, then DRAM traffic
# of one in-step launch from ncu.  Experiment tooling.
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 6 --warmup 2 --no-extra --no-cpu-baseline --capture-steps 1"
if [ "$1" != "ncu-only" ]; then
for h in 0 1 2 3; do
  echo "== SMT_GEMM_L2_HINT=$h"
  SMT_GEMM_L2_HINT=$h $CMD 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('avg_launch_us', round(r['avg_launch_us'],1), 'TFLOP/s', round(r['achieved'],1), 'ms/step', round(d['ms_per_step'],2))"
done
fi
for h in 0 1 2 3; do
  SMT_GEMM_L2_HINT=$h timeout 280 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:block_grad_umma_2sm -s 1 -c 1 --csv --log-file gpurun_out/l2hint_$h.csv $CMD > /dev/null 2>&1
  echo "hint $h:"; grep "block_grad_umma_2sm" gpurun_out/l2hint_$h.csv | awk -F'","' '{print "   " $(NF-2) " [" $(NF-1) "] = " $NF}'
done
