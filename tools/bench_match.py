#!/usr/bin/env python
"""Matcher timings with the library's own per-kernel CUDA-event profile (supporting evidence for profiles/).

    python tools/bench_match.py [--out profiles/r2_match.json] [--sweep]

64 pairs x 2048 x 2048 x 256 (the bench.py step) and, with --sweep, BASELINE config 3 (1k-16k, D = 256 and 64).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multipoint_b200 import _lib, ops  # noqa: E402


def sets(P, N, D, dev, seed=0, noise=0.05):
    g = torch.Generator(device=dev).manual_seed(seed)
    a = torch.nn.functional.normalize(torch.randn((P, N, D), generator=g, device=dev), dim=2)
    perm = torch.randperm(N, generator=g, device=dev)
    b = torch.nn.functional.normalize(a[:, perm] + noise * torch.randn((P, N, D), generator=g, device=dev), dim=2)
    return a, b


def run(P, N, D, metric, iters, dev, **kw):
    a, b = sets(P, N, D, dev)
    fn = lambda: ops.match(a, b, metric=metric, algo='tensor', kind='mutual', cross_check=True, **kw)
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    _lib.profile_begin()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    kern = {k: round(v["total_ms"] * 1e3 / iters, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])}
    fl = 2.0 * P * N * N * D
    tc = kern.get("match_top2_tc_kernel", 0.0)
    row = {"P": P, "N": N, "D": D, "metric": metric, "ms_total": round(ms, 4), "kernels_us": kern,
           "matches": int(out[3].sum()),
           "tc_algorithmic_TFLOPs": round(fl / tc / 1e6, 1) if tc else None, "tc_executed_TFLOPs": round(3 * fl / tc / 1e6, 1) if tc else None,
           "chain_algorithmic_TFLOPs": round(fl / ms / 1e9, 1)}
    print(json.dumps(row), flush=True)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = [run(64, 2048, 256, 'l2', 20, dev), run(64, 2048, 256, 'nn', 20, dev, threshold=0.7), run(64, 2048, 64, 'l2', 20, dev)]
    if args.sweep:
        for D in (256, 64):
            for N in (1024, 2048, 4096, 8192, 16384):
                rows.append(run(1, N, D, 'l2', 20 if N <= 4096 else 8, dev))
    if args.out:
        json.dump({"rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
