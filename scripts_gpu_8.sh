#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
MP_TC_A_SMEM=1 timeout 300 python bench.py --only-hot --steps 20 --warmup 3 > gpurun_out/hot_a_smem.json 2>&1
MP_TC_A_SMEM=0 timeout 300 python bench.py --only-hot --steps 20 --warmup 3 > gpurun_out/hot_a_tmem.json 2>&1
cat gpurun_out/hot_a_smem.json gpurun_out/hot_a_tmem.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"match_top2_tc" -s 6 -c 2 -o gpurun_out/prof_tc \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout -k 10 900 python tools/bench_kernels.py --quick --out gpurun_out/kernels_quick.json > gpurun_out/kernels_quick.log 2>&1
grep -E "match bf|64 pairs" gpurun_out/kernels_quick.log | grep -v SIMT
