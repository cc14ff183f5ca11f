#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?" >> gpurun_out/bench_n1.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
echo "== ncu hot launches exit $?" >> gpurun_out/ncu_launch_hot.log
timeout -k 10 600 ncu --set full --clock-control none --import-source on \
    -k regex:"nms_tile|nms_select|nms_fixup|match_top2_tc|match_recheck|match_prep" \
    -s 21 -c 9 -o gpurun_out/prof_hot python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $?" >> gpurun_out/ncu_full.log
for f in gpurun_out/tests_gpu.log gpurun_out/bench_n1.err gpurun_out/ncu_launch_hot.log gpurun_out/ncu_full.log; do echo "--- $f"; tail -n 4 $f; done
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['hot_path']['ms_per_step']); print({k:round(v['ms'],4) for k,v in d['hot_path']['stages'].items()})"
