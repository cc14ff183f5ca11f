#!/bin/bash
# parity tests + the hot path with per-kernel event times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/tests_gpu.log
MP_BENCH_HOT_KERNELS=1 timeout 300 python bench.py --only-hot --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/hot_profile.log
