#!/bin/bash
mkdir -p gpurun_out
cat > gpurun_out/sp.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from multipoint_b200 import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
lg = torch.randn((128, 65, 64, 80), generator=g, device=dev) * 2.0
lg[:, 64] += 5.0
prob = ops.detector_head(lg).reshape(128, 512, 640)
for _ in range(4):
    ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms_sparse|nms_candidates|nms_select" -s 6 -c 3 -o gpurun_out/prof_sparse python gpurun_out/sp.py > gpurun_out/ncu_sparse.log 2>&1
tail -3 gpurun_out/ncu_sparse.log
