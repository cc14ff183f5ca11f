#!/bin/bash
# first GPU bring-up: parity tests in two stages so a hang in the tcgen05 kernel cannot hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not tensor and not matching_bench and not pipeline and not batched_counts" 2>&1 | tail -150 > gpurun_out/t1.log
echo "== stage1 exit ${PIPESTATUS[0]}" >> gpurun_out/t1.log
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensor or matching_bench or pipeline or batched_counts" 2>&1 | tail -150 > gpurun_out/t2.log
echo "== stage2 exit ${PIPESTATUS[0]}" >> gpurun_out/t2.log
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "== smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/t1.log gpurun_out/t2.log gpurun_out/smoke.log
