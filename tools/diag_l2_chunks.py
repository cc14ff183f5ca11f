#!/usr/bin/env python
"""Experiment: descriptor normalise + sampling in groups of images through one reused channels-last scratch, so that the
map stays in L2 between the two kernels, against the whole-batch launches.  Same kernels, same results."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multipoint_b200 import _lib, ops
from multipoint_b200.ops import _ptr, _stream

dev = torch.device("cuda", 0)
B, D, Hc, Wc, K, H, W = 128, 256, 64, 80, 2048, 512, 640
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn((B, D, Hc, Wc), generator=g, device=dev)
kp = torch.stack([torch.randint(0, H, (B, K), generator=g, device=dev), torch.randint(0, W, (B, K), generator=g, device=dev)], dim=2)
counts = torch.full((B,), K, dtype=torch.int32, device=dev)
lib = _lib.load()


def whole():
    _, nhwc = ops.normalize_descriptors(raw, nchw=False, nhwc=True)
    return ops.sample_descriptors(kp, nhwc, H, W, counts=counts, channels_last=True, split=True)


out = torch.empty((B, K, D), device=dev)
hi = torch.empty((B, K, D), dtype=torch.bfloat16, device=dev)
mid = torch.empty((B, K, D), dtype=torch.bfloat16, device=dev)
sq = torch.empty((B, K), device=dev)


def chunked(C, scratch):
    s = _stream(raw)
    for c0 in range(0, B, C):
        n = min(C, B - c0)
        _lib.check(lib.mp_normalize_descriptors_f32(_ptr(raw[c0:c0 + n]), n, D, Hc * Wc, None, _ptr(scratch), s), "normalize")
        _lib.check(lib.mp_sample_descriptors_split_f32(_ptr(kp[c0:c0 + n]), _ptr(counts[c0:c0 + n]), n, K, _ptr(scratch), D, Hc, Wc, 1, H, W,
                                                       _ptr(out[c0:c0 + n]), _ptr(hi[c0:c0 + n]), _ptr(mid[c0:c0 + n]), _ptr(sq[c0:c0 + n]), s), "sample")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


ref, refsp = whole()
print("whole batch: %.1f us" % timeit(whole))
for C in (2, 4, 8, 12, 16, 24, 32, 64):
    scratch = torch.empty((C, Hc, Wc, D), device=dev)
    us = timeit(lambda: chunked(C, scratch))
    same = torch.equal(out, ref) and torch.equal(hi, refsp['hi']) and torch.equal(sq, refsp['sq_norms'])
    print("groups of %3d images (%5.1f MB scratch): %.1f us  identical=%s" % (C, C * Hc * Wc * D * 4 / 1e6, us, same), flush=True)
