#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/tests_gpu.log
echo "== tests exit ${PIPESTATUS[0]}" >> gpurun_out/tests_gpu.log
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?" >> gpurun_out/bench_n1.err
timeout -k 10 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
echo "== kernels exit $?" >> gpurun_out/kernels.log
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hot.csv \
    python bench.py --only-hot --steps 1 --warmup 3 > gpurun_out/ncu_launch_hot.log 2>&1
echo "== ncu hot launches exit $?" >> gpurun_out/ncu_launch_hot.log
for f in gpurun_out/tests_gpu.log gpurun_out/bench_n1.err gpurun_out/ncu_launch_hot.log; do echo "--- $f"; tail -n 4 $f; done
cat gpurun_out/kernels.log
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['hot_path']['ms_per_step']); print({k:round(v['ms'],4) for k,v in d['hot_path']['stages'].items()})"
