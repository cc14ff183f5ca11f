#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "nms or pipeline" 2>&1 | tail -3
for t in 0 1; do echo "MP_NMS_TILE=$t"; MP_NMS_TILE=$t timeout 300 python tools/nms_density_profile.py 2>&1; done | tee gpurun_out/nms_variants.log
