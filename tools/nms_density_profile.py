#!/usr/bin/env python
"""Per-kernel split of box_nms (dense map out) at several candidate densities, B=128 images of 512x640,
measured with the library's own event profiling (mp_profile_begin / mp_profile_end)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multipoint_b200 import _lib, ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    B, H, W = 128, 512, 640
    rows = []
    for sigma, bias in ((2.0, 5.0), (3.0, 9.0), (4.0, 14.0), (4.0, 20.0)):
        lg = torch.randn((B, 65, 64, 80), generator=g, device=dev) * sigma
        lg[:, 64] += bias
        prob = ops.detector_head(lg).reshape(B, H, W)
        frac = float((prob > 0.015).float().mean())
        for _ in range(3):
            ops.box_nms(prob, 4, 0.015)
        torch.cuda.synchronize()
        _lib.profile_begin()
        for _ in range(10):
            ops.box_nms(prob, 4, 0.015)
        torch.cuda.synchronize()
        prof = _lib.profile_end()
        row = {"candidates_pct": round(100 * frac, 2),
               "kernels_us": {k: round(v["total_ms"] * 100, 1) for k, v in prof.items()}}
        rows.append(row)
        print(json.dumps(row))
    return rows


if __name__ == "__main__":
    main()
