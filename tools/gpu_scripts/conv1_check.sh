#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "conv1 or fused_inference or relu_bn" 2>&1 | tail -3
timeout 100 python - <<'PY' 2>&1 | tail -2
import sys, torch
sys.path.insert(0, '.')
from multipoint_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
img = torch.rand((64, 1, 512, 640), generator=g, device='cuda')
w = torch.randn((64, 1, 3, 3), generator=g, device='cuda')
sc = torch.rand(64, device='cuda') + 0.5
sh = torch.randn(64, device='cuda')
for _ in range(3):
    ops.conv1_relu_bn_pad(img, w, sh, sc, sh)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv1_relu_bn_pad(img, w, sh, sc, sh)
e1.record(); torch.cuda.synchronize()
print("conv1_relu_bn_pad ms per launch", e0.elapsed_time(e1) / 20)
PY
