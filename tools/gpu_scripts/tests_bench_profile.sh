#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "== bench exit $?"; tail -3 gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'hot', d['hot_path']['ms_per_step'], 'launches', d['gpu_launches'])"
timeout 300 python tools/backbone_profile.py 2>&1 | tail -22 | tee gpurun_out/backbone_profile.log
