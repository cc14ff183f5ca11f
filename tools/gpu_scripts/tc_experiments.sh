#!/bin/bash
# where does the tensor-core matcher's time go?  (MP_TC_EXP bits give wrong results; timing only)
mkdir -p gpurun_out
for e in 0 1 2 8 9 11; do
  echo -n "MP_TC_EXP=$e "
  MP_TC_EXP=$e MP_BENCH_HOT_KERNELS=1 timeout 200 python bench.py --only-hot --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['kernels_us_per_step']['match_top2_tc_kernel']/2)"
done | tee gpurun_out/tc_experiments.log
