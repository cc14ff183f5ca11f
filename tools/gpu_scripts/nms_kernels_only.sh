#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "nms" 2>&1 | tail -3
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/nms_sparse_kernels.log
import json, torch, sys
sys.path.insert(0, '.')
from multipoint_b200 import _lib, ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
lg = torch.randn((128, 65, 64, 80), generator=g, device=dev) * 2.0
lg[:, 64] += 5.0
prob = ops.detector_head(lg).reshape(128, 512, 640)
for _ in range(3):
    ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048)
torch.cuda.synchronize()
_lib.profile_begin()
for _ in range(20):
    ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048)
torch.cuda.synchronize()
print(json.dumps({k: round(v['total_ms'] * 50, 1) for k, v in _lib.profile_end().items()}))
PY
