#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/tests_gpu.log
timeout 300 python tools/backbone_profile.py 2>&1 | tail -30 | tee gpurun_out/backbone_profile.log
