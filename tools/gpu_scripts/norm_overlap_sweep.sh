#!/bin/bash
# sweep the CTA cap of the side-stream normalise kernel (hot path only, 64 pairs)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pipeline or normalize" 2>&1 | tail -3
for k in 0 1 2 3 4 6 8; do
  echo "MP_NORM_CTAS_PER_SM=$k" 
  MP_NORM_CTAS_PER_SM=$k timeout 300 python bench.py --only-hot --steps 20 --warmup 5 2>&1 | tail -1
done | tee gpurun_out/norm_overlap_sweep.log
