#!/usr/bin/env python
"""box_nms timings at 128 images of 512x640 (supporting evidence for profiles/): the dense path (keep_top_k=0, what
every shipped config uses) at three candidate densities and the top-k 2048 path, per kernel.

    python tools/bench_nms.py [--out profiles/r2_nms.json] [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multipoint_b200 import _lib, ops  # noqa: E402

H, W, B = 512, 640, 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default=None, help="run only the case whose name contains this")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for sigma, bias in ((2.0, 5.0), (3.0, 9.0), (4.0, 14.0)):
        lg = torch.randn((B, 65, 64, 80), generator=g, device=dev) * sigma
        lg[:, 64] += bias
        prob = ops.detector_head(lg).reshape(B, H, W)
        del lg
        dens = float((prob > 0.015).float().mean())
        cases = [("dense %.1f%% candidates" % (100 * dens), lambda: ops.box_nms(prob, 4, 0.015), B * 2 * H * W * 4)]
        if sigma == 2.0:
            cases.append(("top-k 2048 + keypoints %.1f%%" % (100 * dens),
                          lambda: ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048), B * (2 * H * W * 4 + 20 * 2048)))
            cases.append(("top-k 2048 keypoints only (no dense map) %.1f%%" % (100 * dens),
                          lambda: ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048, want_dense=False),
                          B * (H * W * 4 + 20 * 2048)))
        for name, fn, nbytes in cases:
            if args.only and args.only not in name:
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / args.iters
            _lib.profile_begin()
            for _ in range(args.iters):
                fn()
            torch.cuda.synchronize()
            prof = _lib.profile_end()
            kern = {k: round(v["total_ms"] * 1e3 / args.iters, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])}
            r = {"case": name, "us": round(us, 1), "algorithmic_bytes": nbytes, "GBps": round(nbytes / us / 1e3, 1),
                 "frac_of_hbm": round(nbytes / us / 1e3 / hbm, 4), "kernels_us": kern}
            rows.append(r)
            print(json.dumps(r), flush=True)
        del prob
    if args.out:
        json.dump({"rows": rows, "peak_hbm_gbs": hbm}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
