#!/bin/bash
# run under: gpurun --gpus N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_smi.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1
echo "== check exit $?" >> gpurun_out/multi_check_n$N.log
tail -3 gpurun_out/multi_check_n$N.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "== bench exit $?" >> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['hot_path']['value'])"
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
tail -c 300 gpurun_out/bench_ref_n$N.json
