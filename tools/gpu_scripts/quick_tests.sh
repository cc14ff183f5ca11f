#!/bin/bash
# GPU parity tests + the per-kernel sweep (short)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/tests_gpu.log
timeout 600 python tools/bench_kernels.py --quick --out gpurun_out/kernels_quick.json > gpurun_out/kernels_quick.log 2>&1
tail -4 gpurun_out/kernels_quick.log
