#!/usr/bin/env python
"""Adaptation kernel timings on the homographies export_keypoints samples (supporting evidence for profiles/).

    python tools/bench_adapt.py [--out profiles/r2_adapt.json] [--pairs 2] [--num 100]

warp_kernel / ha_aggregate_kernel / valid_mask_kernel with the library's own per-kernel CUDA-event profile, on the
reference's homography distribution (np.random.seed(0), configs/config_export_keypoints.yaml parameters).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multipoint_b200 import _lib, ops, utils  # noqa: E402

H, W = 512, 640


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--pairs", type=int, default=2)
    ap.add_argument("--num", type=int, default=100)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    cfg = utils._check_ha_config({'num': args.num, 'erosion_radius': 5, 'mask_border': True, 'min_count': 5})
    np.random.seed(0)
    Hs, _ = utils.sample_adaptation_homographies((H, W), cfg, with_masks=False)
    n, B = Hs.shape[0], args.pairs
    Hm = torch.from_numpy(Hs.astype(np.float32))
    A_warp = utils.normalized_warp_matrix(Hm, (H, W), (H, W)).to(dev)
    A_unwarp = utils.normalized_warp_matrix(torch.inverse(Hm), (H, W), (H, W)).to(dev)
    Minv = torch.from_numpy(utils.invert_homographies(Hs)).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand((B, H, W), generator=g, device=dev)
    pa = torch.rand((n, B, H, W), generator=g, device=dev) * 0.3
    pb = torch.rand((n, B, H, W), generator=g, device=dev) * 0.3
    p0 = torch.rand((B, H, W), generator=g, device=dev) * 0.1
    tables = ops.linspace_tables(H, W, dev)
    masks = ops.valid_masks(Minv, H, W, cfg['erosion_radius'], cfg['mask_border'])

    def step():
        ops.valid_masks(Minv, H, W, cfg['erosion_radius'], cfg['mask_border'])
        ops.warp(img, A_warp, 'bilinear', 'reflection', tables)
        ops.ha_aggregate(p0, pa, pb, masks, A_unwarp, 'prod', cfg['min_count'], tables=tables)
        ops.ha_aggregate(p0, pa, None, masks, A_unwarp, 'none', cfg['min_count'], tables=tables)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.profile_begin()
    for _ in range(args.iters):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    HW = H * W
    rows = []
    warp_us = prof["warp_kernel"]["total_ms"] * 1e3 / args.iters
    rows.append({"kernel": "warp_kernel (n=%d, %d planes)" % (n, B), "us": round(warp_us, 1), "bytes": n * B * 2 * HW * 4})
    agg_us = prof["ha_aggregate_kernel"]["total_ms"] * 1e3 / args.iters   # the two launches together
    rows.append({"kernel": "ha_aggregate_kernel prod + single (n=%d, B=%d)" % (n, B), "us": round(agg_us, 1),
                 "bytes": n * (3 * B * HW * 4 + 2 * HW) + 4 * B * HW * 4})
    vm_us = prof["valid_mask_kernel"]["total_ms"] * 1e3 / args.iters
    rows.append({"kernel": "valid_mask_kernel (n=%d)" % n, "us": round(vm_us, 1), "bytes": n * HW})
    for r in rows:
        r["GBps"] = round(r["bytes"] / r["us"] / 1e3, 1)
        r["frac_of_hbm"] = round(r["bytes"] / r["us"] / 1e3 / hbm, 4)
        print(json.dumps(r), flush=True)
    # the two aggregate flavours separately (events around single calls)
    img2 = torch.cat([img, torch.rand((B, H, W), generator=g, device=dev)])
    for name, fn in (("warp, both spectra of the pairs in one launch (n=%d, %d planes in 2 groups)" % (n, 2 * B),
                      lambda: ops.warp(img2, A_warp, 'bilinear', 'reflection', tables, groups=2)),
                     ("ha_aggregate prod (pairs)", lambda: ops.ha_aggregate(p0, pa, pb, masks, A_unwarp, 'prod', cfg['min_count'], tables=tables)),
                     ("ha_aggregate single", lambda: ops.ha_aggregate(p0, pa, None, masks, A_unwarp, 'none', cfg['min_count'], tables=tables)),
                     ("ha_aggregate prod (pairs), TMA-staged variant", lambda: ops.ha_aggregate(p0, pa, pb, masks, A_unwarp, 'prod', cfg['min_count'], tables=tables, staged=True)),
                     ("ha_aggregate single, TMA-staged variant", lambda: ops.ha_aggregate(p0, pa, None, masks, A_unwarp, 'none', cfg['min_count'], tables=tables, staged=True))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        _lib.profile_begin()
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        prof1 = _lib.profile_end()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        if name.startswith("warp"):   # kernel time (the events also see the 1 GB output allocation and the copy into the gather array)
            us = prof1["warp_kernel"]["total_ms"] * 1e3 / args.iters
        nb = n * ((2 if 'prod' in name else 1) * B * HW * 4 + HW) + 2 * B * HW * 4
        if name.startswith("warp"):
            nb = n * 2 * B * 2 * HW * 4
        r = {"kernel": name, "us": round(us, 1), "bytes": nb, "GBps": round(nb / us / 1e3, 1), "frac_of_hbm": round(nb / us / 1e3 / hbm, 4)}
        rows.append(r)
        print(json.dumps(r), flush=True)
    if args.out:
        json.dump({"rows": rows, "peak_hbm_gbs": hbm}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
