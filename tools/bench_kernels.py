#!/usr/bin/env python
"""Per-kernel timings at the BASELINE sizes (not the bench.py contract line: supporting evidence
for profiles/).  CUDA events on the current stream, >= 3 warm-ups, inputs larger than L2 where the
kernel is HBM-bound, roofline fractions against MEASURED_PEAKS.json.

    python tools/bench_kernels.py [--out profiles/r1_kernels.json]

Covers: detector head / normalise / NMS / sampling at 128 images (config 2), the matching sweep
1k-16k x 256-d and 64-d (config 3), the homographic-adaptation warp + aggregate kernels at
100 homographies per image (config 4).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multipoint_b200 import ops  # noqa: E402
from multipoint_b200 import synthetic as syn  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops", 1590.0))
    res = {"peaks": {"hbm_gbs": hbm, "bf16_tflops": tf, "source": "MEASURED_PEAKS.json" if peaks else "fallback"}, "kernels": []}

    def add(name, ms, nbytes=None, flops=None, **extra):
        row = {"kernel": name, "ms": round(ms, 5)}
        if nbytes:
            row["algorithmic_GBps"] = round(nbytes / ms / 1e6, 1)
            row["frac_hbm"] = round(nbytes / ms / 1e6 / hbm, 4)
        if flops:
            row["algorithmic_TFLOPs"] = round(flops / ms / 1e9, 2)
            row["frac_bf16_peak_algorithmic"] = round(flops / ms / 1e9 / tf, 4)
        row.update(extra)
        res["kernels"].append(row)
        print(json.dumps(row))

    H, W, B = 512, 640, 128
    g = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn((B, 65, 64, 80), generator=g, device=dev) * 2.0
    logits[:, 64] += 5.0
    add("detector_head B=128", timed(lambda: ops.detector_head(logits)), B * (65 * 5120 * 4 + H * W * 4))
    prob = ops.detector_head(logits).reshape(B, H, W)
    add("box_nms dense (tile + fixup) B=128", timed(lambda: ops.box_nms(prob, 4, 0.015)), B * 2 * H * W * 4)
    add("box_nms + top-k 2048 + keypoints B=128",
        timed(lambda: ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048)), B * (2 * H * W * 4 + 20 * 2048))
    add("extract_keypoints B=128", timed(lambda: ops.extract_keypoints(prob, 0.2, kp_cap=4096)), B * H * W * 4)
    # the same NMS on a sparser heatmap (a trained detector: ~2 % of the pixels above the threshold instead of the
    # 12.8 % of the synthetic sigma=2 logits): the kernel is instruction-bound on candidates, not on pixels
    for sigma, bias in ((3.0, 9.0), (4.0, 14.0)):
        lg2 = torch.randn((B, 65, 64, 80), generator=g, device=dev) * sigma
        lg2[:, 64] += bias
        pr2 = ops.detector_head(lg2).reshape(B, H, W)
        frac = float((pr2 > 0.015).float().mean())
        add("box_nms dense B=128, %.1f%% candidates" % (100 * frac), timed(lambda: ops.box_nms(pr2, 4, 0.015)), B * 2 * H * W * 4)
        del lg2, pr2
    for D in (256, 64):
        raw = torch.randn((B, D, 64, 80), generator=g, device=dev)
        add("normalize_descriptors NCHW D=%d" % D, timed(lambda: ops.normalize_descriptors(raw, True, False)), B * 2 * 4 * D * 5120)
        add("normalize_descriptors NHWC D=%d" % D, timed(lambda: ops.normalize_descriptors(raw, False, True)), B * 2 * 4 * D * 5120)
        nchw, nhwc = ops.normalize_descriptors(raw, True, True)
        _, kp, _, cnt = ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048)
        nb = B * (min(16 * 2048 * D, 4 * D * 5120) + 4 * 2048 * D + 16 * 2048)
        add("sample_descriptors NHWC D=%d K=2048" % D, timed(lambda: ops.sample_descriptors(kp, nhwc, H, W, counts=cnt, channels_last=True)), nb)
        add("sample_descriptors NCHW D=%d K=2048" % D, timed(lambda: ops.sample_descriptors(kp, nchw, H, W, counts=cnt)), nb)
        del raw, nchw, nhwc

    # the fused glue between the backbone convolutions at the two largest layer shapes (one encoder = 64 images)
    for (Bg, Cg, Hg, Wg, pool) in ((64, 64, 512, 640, False), (64, 64, 512, 640, True), (64, 64, 256, 320, False), (64, 128, 128, 160, True)):
        xg = torch.randn((Bg, Cg, Hg, Wg), generator=g, device=dev)
        scg = torch.rand((Cg,), generator=g, device=dev) + 0.5
        shg = torch.randn((Cg,), generator=g, device=dev)
        cbg = torch.randn((Cg,), generator=g, device=dev)
        og = ops.relu_bn_pad(xg, scg, shg, pool=pool, pad=1, reflect=True, conv_bias=cbg)
        add("relu_bn_pad %dx%dx%dx%d%s + reflection pad" % (Bg, Cg, Hg, Wg, " + maxpool" if pool else ""),
            timed(lambda: ops.relu_bn_pad(xg, scg, shg, pool=pool, pad=1, reflect=True, conv_bias=cbg), iters=10), 4 * (xg.numel() + og.numel()))
        del xg, og
    imgc = torch.rand((64, 1, H, W), generator=g, device=dev)
    wc = torch.randn((64, 1, 3, 3), generator=g, device=dev) * 0.3
    sc64, sh64 = torch.rand((64,), generator=g, device=dev) + 0.5, torch.randn((64,), generator=g, device=dev)
    add("conv1_relu_bn_pad 64 images 1->64 channels 512x640 + reflection pads",
        timed(lambda: ops.conv1_relu_bn_pad(imgc, wc, sh64, sc64, sh64), iters=10), 64 * 4 * (H * W + 64 * (H + 2) * (W + 2)))
    # config 3: matching sweep
    sizes = (1024, 2048, 4096) if args.quick else (1024, 2048, 4096, 8192, 16384)
    for D in (256, 64):
        for N in sizes:
            a, b = syn.descriptor_sets(N, N, N, D, 0.05)
            A, Bm = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
            fl = 2.0 * N * N * D
            it = 20 if N <= 4096 else 5
            ms = timed(lambda: ops.match(A, Bm, metric='l2', algo='tensor', kind='mutual', cross_check=True), iters=it)
            add("match bfmatcher crossCheck N=%d D=%d (prep+GEMM+recheck+select)" % (N, D), ms, flops=fl, executed_TFLOPs=round(3 * fl / ms / 1e9, 1))
            ms = timed(lambda: ops.match(A, Bm, metric='nn', algo='tensor', kind='mutual', cross_check=True, threshold=0.7), iters=it)
            add("match nnmatcher N=%d D=%d" % (N, D), ms, flops=fl, executed_TFLOPs=round(3 * fl / ms / 1e9, 1))
            if N <= 4096:
                ms = timed(lambda: ops.match(A, Bm, metric='l2', algo='simt', kind='mutual', cross_check=True), iters=3)
                add("match bfmatcher SIMT reference N=%d D=%d" % (N, D), ms, flops=fl)
    # batched: 64 pairs of 2048 (the bench.py step)
    a = torch.nn.functional.normalize(torch.randn((64, 2048, 256), generator=g, device=dev), dim=2)
    b = torch.nn.functional.normalize(a[:, torch.randperm(2048, device=dev)] + 0.05 * torch.randn((64, 2048, 256), generator=g, device=dev) / 16, dim=2)
    fl = 2.0 * 64 * 2048 * 2048 * 256
    ms = timed(lambda: ops.match(a, b, metric='l2', algo='tensor', kind='mutual', cross_check=True))
    add("match bfmatcher crossCheck 64 pairs x 2048 x 256", ms, flops=fl, executed_TFLOPs=round(3 * fl / ms / 1e9, 1))

    # config 4: homographic adaptation kernels, 100 homographies per image, one pair
    n, Bp = 99, 1
    img = torch.rand((2 * Bp, H, W), generator=g, device=dev)
    th = torch.rand((n,), generator=g, device=dev) * 0.6 - 0.3
    A = torch.zeros((n, 3, 3), device=dev)
    A[:, 0, 0] = torch.cos(th) * 1.1; A[:, 0, 1] = -torch.sin(th); A[:, 1, 0] = torch.sin(th); A[:, 1, 1] = torch.cos(th) * 1.1
    A[:, 0, 2] = 0.05; A[:, 1, 2] = -0.03; A[:, 2, 2] = 1.0; A[:, 2, 0] = 0.02
    add("warp images bilinear/reflection n=99 x 2 planes", timed(lambda: ops.warp(img, A, 'bilinear', 'reflection')), n * 2 * Bp * 2 * H * W * 4)
    pa = torch.rand((n, Bp, H, W), generator=g, device=dev) * 0.3
    pb = torch.rand((n, Bp, H, W), generator=g, device=dev) * 0.3
    masks = (torch.rand((n, H, W), generator=g, device=dev) > 0.1).to(torch.uint8)
    p0 = torch.rand((Bp, H, W), generator=g, device=dev) * 0.1
    add("ha_aggregate prod n=99 (one pair)", timed(lambda: ops.ha_aggregate(p0, pa, pb, masks, A, 'prod', 2)),
        n * (2 * Bp * H * W * 4 + H * W) + 2 * Bp * H * W * 4)
    add("ha_aggregate single n=99", timed(lambda: ops.ha_aggregate(p0, pa, None, masks, A, 'none', 2)),
        n * (Bp * H * W * 4 + H * W) + 2 * Bp * H * W * 4)
    # the 99 valid masks of one adaptation batch, device raster vs the host cv2 loop it replaces
    import time
    import numpy as np
    from multipoint_b200 import utils
    np.random.seed(0)
    cfg = utils._check_ha_config({})
    Hs, _ = utils.sample_adaptation_homographies((H, W), cfg, with_masks=False)
    Minv = torch.from_numpy(utils.invert_homographies(Hs)).to(dev)
    ms = timed(lambda: ops.valid_masks(Minv, H, W, cfg['erosion_radius'], cfg['mask_border']))
    t0 = time.perf_counter()
    for Hm in Hs[:20]:
        utils.compute_valid_mask((H, W), Hm, cfg['erosion_radius'], cfg['mask_border'])
    host_ms = (time.perf_counter() - t0) * 1e3 * len(Hs) / 20
    add("valid_mask n=99 512x640 erosion 5 + border", ms, len(Hs) * H * W, host_cv2_loop_ms=round(host_ms, 1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
