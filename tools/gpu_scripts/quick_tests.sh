#!/bin/bash
# GPU parity tests + NMS density split (short)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/tests_gpu.log
timeout 300 python tools/nms_density_profile.py 2>&1 | tee gpurun_out/nms_density_profile.log
