#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
MP_NMS_TILE=1 timeout -k 10 600 python -m pytest tests -q -m gpu -k "box_nms or pipeline" 2>&1 | tail -3
timeout 300 python bench.py --only-hot --steps 20 --warmup 3
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"match_top2_tc|nms_tile_fast" -s 7 -c 3 -o gpurun_out/prof_tc \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['hot_path']['ms_per_step']); print({k:round(v['ms'],4) for k,v in d['hot_path']['stages'].items()})"
