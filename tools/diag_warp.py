#!/usr/bin/env python
"""Image-warp timings: planes per launch, gather array on/off (MP_WARP_NO_GATHER=1)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multipoint_b200 import _lib, ops, utils
H, W = 512, 640
dev = torch.device("cuda", 0)
cfg = utils._check_ha_config({'num': 100})
np.random.seed(0)
Hs, _ = utils.sample_adaptation_homographies((H, W), cfg, with_masks=False)
A = utils.normalized_warp_matrix(torch.from_numpy(Hs.astype(np.float32)), (H, W), (H, W)).to(dev)
tables = ops.linspace_tables(H, W, dev)
for planes, groups in ((2, 1), (4, 2), (4, 1), (8, 2)):
    img = torch.rand((planes, H, W), device=dev)
    for _ in range(3):
        ops.warp(img, A, 'bilinear', 'reflection', tables, groups=groups)
    torch.cuda.synchronize()
    _lib.profile_begin()
    for _ in range(5):
        ops.warp(img, A, 'bilinear', 'reflection', tables, groups=groups)
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    print(planes, groups, {k: round(v["total_ms"] * 1e3 / 5, 1) for k, v in prof.items()}, flush=True)
