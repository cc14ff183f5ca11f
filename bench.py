#!/usr/bin/env python
"""Benchmark of the keypoint extract-and-match path (BASELINE.json metric: image pairs/s at
512x640, extract+match).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (library port)

Own arm.  One process per GPU (torchrun for N>1), weak scaling: every rank runs the workload of
BASELINE config 2 -- 64 synthetic 512x640 optical/thermal pairs, box NMS size 4, threshold 0.015,
top-k 2048 -- followed by mutual-NN matching of each pair, on its own shard of pairs; there is no
data-path collective (pairs are independent units).  A step is one pass over the 64
pairs: MultiPoint forward for both spectra in one batch (cuDNN fp32 backbone, random-init weights
with calibrated final BatchNorms so heatmaps are not degenerate) -> detector-head kernel -> NMS +
top-k + ordered keypoints -> descriptor normalise (channels-last) -> descriptor sampling ->
tcgen05 matcher.
  value     pairs/s with the images already resident in HBM (CUDA events, max over ranks)
  e2e       the same step from pinned host images: H2D of the images + the step + D2H of keypoints
            and matches, through the public pipeline API
  hot_path  the same without the cuDNN backbone (backbone outputs resident in HBM): the part this
            repo implements; per-stage times and the roofline of its dominant kernel
  roofline  the dominant kernel of the path this repo implements: the tcgen05 matcher, algorithmic
            flop = 2*N1*N2*D once per pair (the three bf16 passes it executes are reported beside it)
  config3_matching_sweep   BASELINE config 3: 1k-16k keypoints x 256-d (and 64-d) per pair, whole matcher chain
  config4_adaptation       BASELINE config 4: export_keypoints' homographic adaptation, 100 homographies per image;
            at N > 1 also one batch with the homography samples sharded over the ranks and the two accumulators
            summed by one NCCL all-reduce (the only data-path collective of the whole path)
  cpu_baseline  (N=1) the reference's CPU chain for ONE pair on the host cores: torch CPU backbone +
            oracle/reference_port.py (torch softmax / torchvision nms / grid_sample / cv2.BFMatcher)
Reference arm: that same CPU chain, one pair per step (value is per pair, so it compares with the 64-pair step's
pairs/s), all host threads, rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 512, 640
METRIC = "image_pairs_per_sec_512x640_extract_match"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="image pairs per rank per step")
    ap.add_argument("--desc", type=int, default=256, help="descriptor size (256 = class default, 64 = shipped params.yaml)")
    ap.add_argument("--topk", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config 3 / config 4 sections")
    ap.add_argument("--only-value", action="store_true", help="time only the resident step (used under ncu)")
    ap.add_argument("--only-hot", action="store_true",
                    help="run only the post-backbone hot path on synthetic backbone outputs (used under ncu)")
    return ap.parse_args()


def model_config(desc):
    # MultiPoint class defaults (multipoint/models/MultiPoint.py:9-23): two encoders, 256-d descriptors
    return {'multispectral': True, 'descriptor_size': desc}


def workload_config(args, extra=None):
    cfg = {"workload": "config2+matching: %d synthetic 512x640 optical/thermal pairs per GPU per step, "
                       "MultiPoint(multispectral, D=%d) fp32 cuDNN backbone, box NMS size 4 thr 0.015 top-k %d, "
                       "bfmatcher crossCheck" % (args.pairs, args.desc, args.topk),
           "pairs_per_gpu": args.pairs, "descriptor_size": args.desc, "topk": args.topk, "nms": 4,
           "detection_threshold": 0.015, "weights": "random init (seed 0), final BatchNorms calibrated",
           "l2_policy": "inputs larger than L2 (168 MB of images, 1.5 GB of backbone outputs per step)",
           "precision": "fp32 throughout: cudnn.allow_tf32=False and matmul.allow_tf32=False (the reference's CPU arithmetic). "
                        "Stock PyTorch leaves cudnn.allow_tf32=True, i.e. the reference on any Ampere-or-later GPU would run its "
                        "convolutions in TF32: that configuration is reported as backbone_tf32_context, not as the headline"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.gpu_id, self.proc, self.path = gpu_id, None, "/tmp/mp_bench_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_id)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val == "Active":
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU chain
def cpu_reference_chain(net_cpu, pair_np, topk, threads):
    """The reference's CPU path for one pair: two forward passes (torch CPU) + the library port of
    box_nms / nonzero / interpolate_descriptors / get_matches.  Returns seconds and the match count."""
    import cv2
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_port as rp
    torch.set_num_threads(threads)
    cv2.setNumThreads(threads)
    t0 = time.perf_counter()
    with torch.no_grad():
        lo, ro = net_cpu.backbone_outputs({'image': torch.from_numpy(pair_np['optical']['image']),
                                           'is_optical': torch.from_numpy(pair_np['optical']['is_optical'])})
        lt, rt = net_cpu.backbone_outputs({'image': torch.from_numpy(pair_np['thermal']['image']),
                                           'is_optical': torch.from_numpy(pair_np['thermal']['is_optical'])})
        t1 = time.perf_counter()
        kps, ds, matches = rp.pair_chain(torch.cat([lo, lt]), torch.cat([ro, rt]), H, W, 4, 0.015, topk)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, len(matches), [len(k) for k in kps]


def build_net(desc, device):
    import torch
    from multipoint_b200 import synthetic as syn
    from multipoint_b200.models import MultiPoint
    from multipoint_b200.pipeline import calibrate_random_init
    torch.manual_seed(0)
    net = MultiPoint(model_config(desc)).eval()
    calib = syn.image_pair_batch(999, 2, H, W)
    imgs = torch.from_numpy(__import__("numpy").concatenate([calib['optical']['image'], calib['thermal']['image']]))
    opt = torch.tensor([[True], [True], [False], [False]])
    net = net.to(device)
    calibrate_random_init(net, imgs.to(device), is_optical=opt.to(device))
    return net


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    import torch
    from multipoint_b200 import synthetic as syn
    threads = os.cpu_count() or 1
    net_cpu = build_net(args.desc, "cpu")
    times = []
    for i in range(args.warmup + args.steps):
        pair = syn.image_pair_batch(5000 + i, 1, H, W)
        sec, fwd, nm, nk = cpu_reference_chain(net_cpu, pair, args.topk, threads)
        if i >= args.warmup:
            times.append(sec)
    ms = 1000.0 * sum(times) / len(times)
    v = 1000.0 / ms
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, {"sample": "1 pair per step (bounded sample of the 64-pair workload; value is per pair, so it compares "
                                                       "with the 64-pair step's pairs/s)",
                                             "backbone": "multipoint_b200.models.MultiPoint on the CPU through torch's own modules (same module list and "
                                                         "state dict as the reference class, which cannot be imported on the GPU box); everything after it "
                                                         "is the reference's own library calls (oracle/reference_port.py)"}),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": "1 pair per step: torch CPU backbone x2 + oracle/reference_port.py "
                                       "(torch softmax, torchvision batched_nms, grid_sample, cv2.BFMatcher crossCheck)"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------- BASELINE configs 3 and 4
def matching_sweep(dev, hbm_peak, tf_peak):
    """BASELINE config 3: dense mutual-NN matching, 1k-16k keypoints x 256-d (and 64-d, the shipped size) for one
    pair: the whole chain behind get_matches('bfmatcher', crossCheck=True) on device-resident descriptors
    (prep, one tcgen05 GEMM with both directions, flag, fp64 recheck, mutual test, compaction)."""
    import torch
    from multipoint_b200 import _lib, ops
    rows = []
    for D in (256, 64):
        for N in (1024, 2048, 4096, 8192, 16384):
            g = torch.Generator(device=dev).manual_seed(N + D)
            a = torch.nn.functional.normalize(torch.randn((1, N, D), generator=g, device=dev), dim=2)
            perm = torch.randperm(N, generator=g, device=dev)
            b = torch.nn.functional.normalize(a[:, perm] + 0.05 * torch.randn((1, N, D), generator=g, device=dev), dim=2)
            fn = lambda: ops.match(a, b, metric='l2', algo='tensor', kind='mutual', cross_check=True)  # noqa: E731
            iters = 10 if N <= 4096 else 5
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
            tc = prof.get("match_top2_tc_kernel", {"total_ms": 0.0})["total_ms"] / iters
            fl = 2.0 * N * N * D
            rows.append({"N": N, "D": D, "ms": round(ms, 4), "matches": int(out[3].sum()), "gemm_launches": int(prof.get("match_top2_tc_kernel", {"launches": 0})["launches"] / iters),
                         "gemm_us": round(tc * 1e3, 1), "algorithmic_TFLOPs": round(fl / ms / 1e9, 1),
                         "gemm_algorithmic_TFLOPs": round(fl / tc / 1e9, 1) if tc else None,
                         "gemm_frac_of_bf16_peak": round(fl / tc / 1e9 / tf_peak, 4) if tc else None,
                         "gemm_executed_frac": round(3 * fl / tc / 1e9 / tf_peak, 4) if tc else None})
            del a, b
    return {"workload": "config 3: get_matches('bfmatcher', crossCheck=True) for one pair, N x N x D, descriptors resident in HBM",
            "flop": "2*N*N*D once per pair (algorithmic); the GEMM executes 3 bf16 passes of it", "rows": rows}


def adaptation_bench(net, dev, rank, world, hbm_peak, n_pairs=2, num=100, steps=2):
    """BASELINE config 4: what export_keypoints does per batch -- homographic_adaptation_multispectral with ``num``
    homographies (99 sampled + identity) on ``n_pairs`` 512x640 pairs, then box NMS and keypoints.  Returns
    images/s for the whole call (backbone included) and the library's kernels with their HBM fractions."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from multipoint_b200 import _lib, parallel, utils
    from multipoint_b200 import synthetic as syn
    cfg = {'num': num, 'aggregation': 'prod', 'erosion_radius': 5, 'mask_border': True, 'min_count': 5, 'filter_size': 0}
    batch = syn.image_pair_batch(7000 + rank, n_pairs, H, W)
    data = {s: {k: torch.from_numpy(v).to(dev) for k, v in batch[s].items() if k != 'valid_mask'} for s in ('optical', 'thermal')}

    def call(shard=None, Hs=None):
        with torch.no_grad():
            prob = utils.homographic_adaptation_multispectral(data, net, cfg, homographies=Hs, shard=shard)
            nms = utils.box_nms(prob, 4, 0.015, keep_top_k=0)
            return prob, nms

    np.random.seed(1234)
    call()
    torch.cuda.synchronize()
    _lib.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        prob, nms = call()
    e1.record()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    ms = e0.elapsed_time(e1) / steps
    n = num - 1
    HW = H * W
    algorithmic = {   # bytes per BATCH (the samples go through in chunks, so a batch is several launches)
        "warp_kernel": 2 * n * n_pairs * 2 * HW * 4,              # both spectra: every warped plane read + written
        "ha_aggregate_kernel": n * (2 * n_pairs * HW * 4 + HW) + 2 * n_pairs * HW * 4,   # heatmaps + mask read, result written
        "valid_mask_kernel": n * HW,
    }
    rows = []
    for name in ("warp_kernel", "ha_aggregate_kernel", "valid_mask_kernel", "detector_head_kernel", "nms_tile_fast_kernel"):
        if name in prof:
            rec = prof[name]
            ms_batch = rec["total_ms"] / steps
            row = {"kernel": name, "launches_per_batch": rec["launches"] / steps, "us_per_batch": round(ms_batch * 1e3, 1)}
            if name in algorithmic:
                row.update(algorithmic_bytes_per_batch=algorithmic[name], achieved_GBps=round(algorithmic[name] / ms_batch / 1e6, 1),
                           frac_of_hbm=round(algorithmic[name] / ms_batch / 1e6 / hbm_peak, 4))
            rows.append(row)
    lib_ms = sum(v["total_ms"] for k, v in prof.items() if k in ("warp_kernel", "ha_aggregate_kernel", "valid_mask_kernel")) / steps
    res = {"workload": "config 4: homographic_adaptation_multispectral(num=%d, prod) + box_nms(topk=0) on %d synthetic 512x640 pairs "
                       "per GPU (the export_keypoints batch)" % (num, n_pairs),
           "ms_per_batch": round(ms, 2), "pairs_per_s": round(n_pairs * world * 1000.0 / ms, 3),
           "images_through_backbone_per_s": round(2 * n_pairs * num * world * 1000.0 / ms, 1),
           "adaptation_kernels_ms_per_batch": round(lib_ms, 3), "kernels": rows,
           "keypoints_per_pair": int((nms > 0.015).sum()) // n_pairs}
    if world > 1:
        # the one data-path collective: homography samples sharded over the ranks (same images on every rank),
        # partial (prob, count) accumulators summed by one NCCL all-reduce each, then the fused finish
        data0 = [{k: v.clone() for k, v in data[s].items()} for s in ('optical', 'thermal')]
        for d in data0:
            for v in d.values():
                dist.broadcast(v, 0)
        shared = {'optical': data0[0], 'thermal': data0[1]}
        ar = {"ms": 0.0, "bytes": 0, "calls": 0}

        def timed_all_reduce(t):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            a1.record()
            ar.setdefault("events", []).append((a0, a1))
            ar["bytes"] += t.numel() * t.element_size()
            ar["calls"] += 1
            return t

        def call_sharded():
            hw = (H, W)
            Hs, _ = parallel.broadcast_homographies(lambda: utils.sample_adaptation_homographies(hw, utils._check_ha_config(cfg), with_masks=False), device=dev)
            with torch.no_grad():
                return utils.homographic_adaptation_multispectral(shared, net, cfg, homographies=Hs, shard=(rank, world, timed_all_reduce))

        np.random.seed(4321)
        call_sharded()
        ar.update(ms=0.0, bytes=0, calls=0, events=[])
        np.random.seed(4321)           # rank 0 draws the samples: the same stream as the unsharded check below
        dist.barrier(); torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        out_sh = call_sharded()
        s1.record()
        torch.cuda.synchronize(); dist.barrier()
        ms_sh = torch.tensor([s0.elapsed_time(s1), sum(a.elapsed_time(b) for a, b in ar["events"])], dtype=torch.float64, device=dev)
        dist.all_reduce(ms_sh, op=dist.ReduceOp.MAX)
        # the same batch unsharded on this rank: the sharded result must agree to fp32 summation order
        np.random.seed(4321)
        Hs_ref, _ = utils.sample_adaptation_homographies((H, W), utils._check_ha_config(cfg), with_masks=False)
        with torch.no_grad():
            ref = utils.homographic_adaptation_multispectral(shared, net, cfg, homographies=Hs_ref)
        diff = (out_sh - ref).abs().max().reshape(1).double()
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        # the collective on its own, ranks aligned by a barrier: two all-reduces of (B,H,W) fp32
        buf = torch.zeros_like(out_sh[:, 0])
        dist.all_reduce(buf)
        dist.barrier(); torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(10):
            dist.all_reduce(buf)
        c1.record()
        torch.cuda.synchronize()
        iso = torch.tensor([c0.elapsed_time(c1) / 10 * 2], dtype=torch.float64, device=dev)
        dist.all_reduce(iso, op=dist.ReduceOp.MAX)
        res["sharded"] = {"what": "one batch, identity pass + %d homography samples split round-robin over %d ranks (the identity pass counts as rank 0's first unit); NCCL all-reduce(SUM) of prob and count" % (n, world),
                          "all_reduce_ms_isolated": round(float(iso[0]), 4),
                          "all_reduce_note": "all_reduce_ms is measured inside the batch and includes waiting for the slowest rank; "
                                             "all_reduce_ms_isolated is the same two collectives with the ranks aligned",
                          "ms_per_batch": round(float(ms_sh[0]), 2), "pairs_per_s": round(n_pairs * 1000.0 / float(ms_sh[0]), 3),
                          "all_reduce_ms": round(float(ms_sh[1]), 3), "all_reduce_calls": ar["calls"], "all_reduce_bytes": ar["bytes"],
                          "max_abs_diff_vs_unsharded": float(diff[0])}
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from multipoint_b200 import _lib, ops, parallel
    from multipoint_b200 import synthetic as syn
    from multipoint_b200.pipeline import KeypointPipeline

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    rank, local_rank, world = parallel.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = False           # fp32 like the reference: no reduced precision
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = os.environ.get("MP_BENCH_NO_AUTOTUNE", "0") != "1"   # autotuning floods ncu
    K, Wm, P = args.steps, max(args.warmup, 3), args.pairs

    if args.only_hot:
        # profiling aid: the hot path alone at the full batch size, on synthetic backbone outputs
        # (logits ~ 2*N(0,1) with +5 on the dustbin, random descriptor maps), no cuDNN in the process
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        logits = torch.randn((2 * P, 65, H // 8, W // 8), generator=g, device=dev) * 2.0
        logits[:, 64] += 5.0
        raw = torch.randn((2 * P, args.desc, H // 8, W // 8), generator=g, device=dev)
        pipe = KeypointPipeline(None, nms=4, detection_threshold=0.015, topk=args.topk, metric='l2', cross_check=True, dense_nms_map=False)

        def hot_only():
            ext = pipe.extract_from_backbone(logits, raw, H, W)
            pipe.match({k: v[:P] for k, v in ext.items()}, {k: v[P:] for k, v in ext.items()})

        for _ in range(Wm):
            hot_only()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            hot_only()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        launches = (_lib.launch_count() - l0) // K
        kern = None
        if os.environ.get("MP_BENCH_HOT_KERNELS"):   # per-kernel event times of the same steps (adds an event per launch)
            _lib.profile_begin()
            for _ in range(K):
                hot_only()
            torch.cuda.synchronize()
            kern = {k: round(v["total_ms"] * 1e3 / K, 1) for k, v in sorted(_lib.profile_end().items(), key=lambda kv: -kv[1]["total_ms"])}
        print(json.dumps({"note": "--only-hot run (profiling aid, not a bench line)", "ms_per_step": ms,
                          "pairs_per_s": P * 1000.0 / ms, "gpu_launches_per_step": launches, "kernels_us_per_step": kern}))
        return 0

    net = build_net(args.desc, dev)
    pipe = KeypointPipeline(net, nms=4, detection_threshold=0.015, topk=args.topk, metric='l2', cross_check=True, dense_nms_map=False)
    batch = syn.image_pair_batch(1000 + rank, P, H, W)
    host = {s: {k: torch.from_numpy(v).pin_memory() for k, v in batch[s].items() if k != 'valid_mask'} for s in ('optical', 'thermal')}
    resident = {s: {k: v.to(dev) for k, v in host[s].items()} for s in host}

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(); barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)

    # ---- value: whole step, images resident in HBM
    out_holder = {}

    def step_resident():
        out_holder['r'] = pipe(resident)

    gpu_uuid = None
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        pass
    sampler = ClockSampler(gpu_uuid if gpu_uuid else local_rank)
    for _ in range(Wm):
        step_resident()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = _lib.launch_count()
    ms_value = timed(step_resident, K, 0)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()

    if args.only_value:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": P * world * 1000.0 / ms_value, "unit": "pairs/s", "ms_per_step": ms_value,
                              "gpu_launches": int(launches), "note": "--only-value run (profiling aid, not a bench line)"}))
        return 0

    # ---- e2e: pinned host images -> H2D -> step -> D2H of keypoints and matches
    res = out_holder['r']
    d2h_src = lambda r: [r['optical']['keypoints'], r['optical']['counts'], r['thermal']['keypoints'], r['thermal']['counts'],  # noqa: E731
                         r['matches']['query'], r['matches']['train'], r['matches']['distance'], r['matches']['counts']]
    pinned_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in d2h_src(res)]
    h2d_bytes = sum(v.numel() * v.element_size() for s in host for v in host[s].values())
    d2h_bytes = sum(t.numel() * t.element_size() for t in pinned_out)

    def run_e2e(n):
        # the public streaming API: every step uploads its own batch from pinned host memory (the upload of step
        # i+1 overlaps step i on a copy stream) and brings keypoints + matches back to pinned host memory
        for r in pipe.stream((host for _ in range(n)), dev):
            for dst, src in zip(pinned_out, d2h_src(r)):
                dst.copy_(src, non_blocking=True)

    run_e2e(Wm)
    barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(K)
    e1.record()
    torch.cuda.synchronize(); barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / K)
    n_matches = int(pinned_out[-1].sum())
    n_kp = int(pinned_out[1].sum()) + int(pinned_out[3].sum())

    # ---- hot path only: backbone outputs resident in HBM
    both = {'image': torch.cat([resident['optical']['image'], resident['thermal']['image']]),
            'is_optical': torch.cat([resident['optical']['is_optical'], resident['thermal']['is_optical']])}
    with torch.no_grad():
        logits, raw = net.backbone_outputs(both)
    torch.cuda.synchronize()

    def hot():
        ext = pipe.extract_from_backbone(logits, raw, H, W)
        ea = {k: v[:P] for k, v in ext.items()}
        eb = {k: v[P:] for k, v in ext.items()}
        out_holder['h'] = (ext, pipe.match(ea, eb))

    ms_hot = timed(hot, K, Wm)
    ext, _ = out_holder['h']
    B2 = 2 * P
    prob = ext['prob'].reshape(B2, H, W)
    kp, cnt, desc_s = ext['keypoints'], ext['counts'], ext['desc']
    stages = {
        "detector_head": (lambda: ops.detector_head(logits), B2 * (65 * 5120 * 4 + H * W * 4)),
        "nms_tile(+fixup launch)": (lambda: ops.box_nms(prob, 4, 0.015), B2 * 2 * H * W * 4),
        "nms_full(top-k+keypoints, no dense map)": (lambda: ops.box_nms(prob, 4, 0.015, keep_top_k=args.topk, want_keypoints=True, kp_cap=args.topk,
                                                                        want_dense=False),
                                                    B2 * (H * W * 4 + 20 * args.topk)),
        "normalize_desc_nhwc": (lambda: ops.normalize_descriptors(raw, nchw=False, nhwc=True), B2 * 2 * 4 * args.desc * 5120),
        "sample_descriptors": (lambda: ops.sample_descriptors(kp, ops.normalize_descriptors(raw, nchw=False, nhwc=True)[1], H, W, counts=cnt, channels_last=True,
                                                              split=args.desc in (64, 128, 256)), None),
        "match(both directions+select)": (lambda: pipe.match({k: v[:P] for k, v in ext.items()}, {k: v[P:] for k, v in ext.items()}), None),
    }
    kernels = {}
    for name, (fn, nbytes) in stages.items():
        ms = timed(fn, max(K, 10), 3)
        kernels[name] = {"ms": ms}
        if nbytes:
            kernels[name]["algorithmic_GBps"] = nbytes / ms / 1e6
    # sample_descriptors timed above includes the normalise it depends on; subtract
    kernels["sample_descriptors"]["ms"] = max(0.0, kernels["sample_descriptors"]["ms"] - kernels["normalize_desc_nhwc"]["ms"])
    # gathered corner rows + the fp32 rows written + (split) the two bf16 planes and the squared norms written for the matcher
    kernels["sample_descriptors"]["algorithmic_GBps"] = B2 * (16 * args.topk * args.desc + 8 * args.topk * args.desc + 20 * args.topk) / max(kernels["sample_descriptors"]["ms"], 1e-6) / 1e6
    flops = 2.0 * P * args.topk * args.topk * args.desc
    kernels["match(both directions+select)"]["algorithmic_TFLOPs"] = flops / kernels["match(both directions+select)"]["ms"] / 1e9
    kernels["match(both directions+select)"]["executed_TFLOPs"] = 3 * flops / kernels["match(both directions+select)"]["ms"] / 1e9   # one GEMM, three bf16 passes

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md: 6650 GB/s, 1590 TFLOP/s)"

    # ---- per-kernel durations: CUDA events recorded by the library on the launching stream, after
    #      every one of its kernels, over K more WHOLE steps (mp_profile_begin / mp_profile_end): the
    #      post-backbone kernels and the fused glue kernel that runs between the cuDNN convolutions
    glue = {"bytes": 0, "calls": 0}
    _orig_glue = ops.relu_bn_pad

    def _counting_glue(x, *a, **kw):           # algorithmic bytes of the glue kernel: input read once + output written once
        out = _orig_glue(x, *a, **kw)
        glue["bytes"] += 4 * (x.numel() + out.numel())
        glue["calls"] += 1
        return out

    ops.relu_bn_pad = _counting_glue
    step_resident()
    ops.relu_bn_pad = _orig_glue
    torch.cuda.synchronize()
    _lib.profile_begin()
    for _ in range(K):
        step_resident()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    Kp, Dd = args.topk, args.desc
    algorithmic = {  # per launch: ("hbm", bytes) or ("tensor", flops); SURVEY.md 8d / DESIGN.md section 4
        "detector_head_kernel": ("hbm", B2 * (65 * 5120 * 4 + H * W * 4)),
        "nms_tile_fast_kernel": ("hbm", B2 * 2 * H * W * 4),
        "nms_candidates_kernel": ("hbm", B2 * H * W * 4),       # heatmap read (keypoints only: no dense map; the candidate list is extra)
        "normalize_desc_kernel": ("hbm", B2 * 2 * 4 * Dd * 5120),
        # descriptor map read once + fp32 rows written + the matcher's bf16 planes (2 x 2 B) and squared norms written
        "sample_descriptors_kernel": ("hbm", B2 * (min(16 * Kp * Dd, 4 * Dd * 5120) + 8 * Kp * Dd + 20 * Kp)),
        "match_prep_vec_kernel": ("hbm", P * Kp * Dd * (4 + 2 + 2)),
        "match_top2_tc_kernel": ("tensor", 2.0 * P * Kp * Kp * Dd),
    }
    if glue["calls"]:   # launches of different sizes (one per layer): average algorithmic bytes per launch
        algorithmic["relu_bn_pad_kernel"] = ("hbm", glue["bytes"] / glue["calls"])
    # first encoder layer, one launch per encoder (P images each): image read + padded 64-channel activation written
    algorithmic["conv1_relu_bn_pad_kernel"] = ("hbm", P * 4 * (H * W + 64 * (H + 2) * (W + 2)))
    # DRAM bytes per launch from the committed ncu --set full capture of this same workload (bench.py --only-hot);
    # only meaningful at the default sizes the capture was taken at
    traffic = {}
    if (P, Kp, Dd, H, W) == (64, 2048, 256, 512, 640):
        try:
            tpath = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
            tj = json.load(open(tpath if os.path.exists(tpath) else os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
            traffic = {k: v["dram_bytes"] for k, v in tj["per_launch"].items()}
            traffic["sample_descriptors_kernel"] = traffic.get("sample_descriptors_nhwc_vec_kernel")
            traffic["nms_sparse_kernel"] = traffic.get("nms_sparse2_kernel", traffic.get("nms_sparse_kernel"))
            # the glue kernel's launches differ per layer: DRAM bytes / algorithmic bytes of the captured full-resolution
            # launches, applied to the average algorithmic bytes per launch
            if "relu_bn_pad_kernel_dram_over_algorithmic" in tj and glue["calls"]:
                traffic["relu_bn_pad_kernel"] = int(tj["relu_bn_pad_kernel_dram_over_algorithmic"] * glue["bytes"] / glue["calls"])
            traffic_src = tj["source"]
        except Exception:
            traffic = {}
    kernel_rows = []
    for name, rec in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]):
        avg_ms = rec["total_ms"] / max(rec["launches"], 1)
        row = {"kernel": name, "launches_per_step": rec["launches"] / K, "avg_us": round(avg_ms * 1e3, 2),
               "ms_per_step": round(rec["total_ms"] / K, 4)}
        if name in algorithmic:
            kind, work = algorithmic[name]
            row["bound"] = kind
            if kind == "hbm":
                row.update(algorithmic_bytes_per_launch=work, achieved_GBps=round(work / avg_ms / 1e6, 1),
                           frac=round(work / avg_ms / 1e6 / hbm_peak, 4))
            else:
                row.update(algorithmic_flop_per_launch=work, achieved_TFLOPs=round(work / avg_ms / 1e9, 1),
                           frac=round(work / avg_ms / 1e9 / tf_peak, 4),
                           executed_TFLOPs=round(3 * work / avg_ms / 1e9, 1), executed_frac=round(3 * work / avg_ms / 1e9 / tf_peak, 4))
        if traffic.get(name):
            row["traffic"] = traffic[name]
        kernel_rows.append(row)
    # the roofline object is the dominant kernel of the path SURVEY section 8 scopes (the post-backbone chain); the backbone
    # glue kernels (relu_bn_pad / conv1_relu_bn_pad, outside section 8) keep their rows in hot_path.kernels
    section8 = ("match_top2_tc_kernel", "detector_head_kernel", "nms_tile_fast_kernel", "nms_candidates_kernel", "normalize_desc_kernel",
                "sample_descriptors_kernel", "match_prep_vec_kernel")
    dom = next((r for r in kernel_rows if "bound" in r and r["kernel"] in section8), None)
    if dom is not None and dom["bound"] == "tensor":
        roofline = {"kernel": dom["kernel"], "bound": "tensor", "achieved": dom["achieved_TFLOPs"], "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": dom["frac"], "traffic": dom.get("traffic"), "avg_launch_us": dom["avg_us"], "launches_per_step": dom["launches_per_step"],
                    "algorithmic_flop_per_launch": dom["algorithmic_flop_per_launch"],
                    "algorithmic_flop": "2*N1*N2*D per pair, counted once (both directions come out of the one GEMM) x %d pairs per launch" % P,
                    "executed": {"TFLOP/s": dom["executed_TFLOPs"], "frac": dom["executed_frac"],
                                 "note": "3 bf16 MMA passes (hi*hi, hi*mid, mid*hi) per algorithmic fp32 product"},
                    "peak_source": peak_src + " bf16_tflops (burst)"}
        tf_sus = float(peaks.get("bf16_tflops_sustained", 0.0)) if peaks else 0.0
        if tf_sus > 0:
            # the same file's sustained cuBLAS rate: this kernel runs power-capped (ncu: SM clock 1.54 GHz under it), which
            # is the regime that figure describes; `peak` / `frac` above stay on the stricter burst number
            roofline["sustained"] = {"peak": tf_sus, "frac": round(dom["achieved_TFLOPs"] / tf_sus, 4),
                                     "executed_frac": round(dom["executed_TFLOPs"] / tf_sus, 4),
                                     "note": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS back to back for 4 s)"}
    elif dom is not None:
        roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": dom["frac"], "traffic": dom.get("traffic"), "avg_launch_us": dom["avg_us"], "launches_per_step": dom["launches_per_step"],
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "peak_source": peak_src + " hbm_gbs"}
    else:
        roofline = None
    if roofline is not None and roofline["traffic"] is not None:
        roofline["traffic_unit"] = "B per launch (dram__bytes_read.sum + dram__bytes_write.sum)"
        roofline["traffic_source"] = traffic_src

    line = {"metric": METRIC, "value": P * world * 1000.0 / ms_value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks,
            "e2e": {"value": P * world * 1000.0 / ms_e2e, "unit": "pairs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches),
            "hot_path": {"value": P * world * 1000.0 / ms_hot, "unit": "pairs/s", "ms_per_step": ms_hot,
                         "note": "backbone outputs resident in HBM; everything after cuDNN", "stages": kernels,
                         "kernels": kernel_rows},
            "roofline": roofline,
            "result_check": {"keypoints_per_step": n_kp, "matches_per_step": n_matches}}

    # context only (not the headline): the same resident step with cuDNN allowed to use TF32 tensor cores
    # for the backbone convolutions -- PyTorch's own default, and what the reference would run on a GPU
    torch.backends.cudnn.allow_tf32 = True
    ms_tf32 = timed(step_resident, max(3, K // 2), 3)
    torch.backends.cudnn.allow_tf32 = False
    line["backbone_tf32_context"] = {"value": P * world * 1000.0 / ms_tf32, "unit": "pairs/s", "ms_per_step": ms_tf32,
                                     "note": "cudnn.allow_tf32=True for the backbone only; reported for context, headline stays fp32"}

    if not args.no_extra_configs:
        if rank == 0:
            line["config3_matching_sweep"] = matching_sweep(dev, hbm_peak, tf_peak)
        barrier()
        line["config4_adaptation"] = adaptation_bench(net, dev, rank, world, hbm_peak)

    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        net_cpu = build_net(args.desc, "cpu")
        pair = syn.image_pair_batch(5000, 1, H, W)
        cpu_reference_chain(net_cpu, pair, args.topk, threads)          # warm-up (allocator, thread pools)
        secs = [cpu_reference_chain(net_cpu, syn.image_pair_batch(5001 + i, 1, H, W), args.topk, threads) for i in range(2)]
        sec = sum(s[0] for s in secs) / len(secs)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "pairs/s", "cores": threads, "kind": "port",
                                "sample": "1 pair (of the 64-pair step), mean of 2 runs after 1 warm-up: torch CPU backbone x2 + "
                                          "oracle/reference_port.py (torch softmax, torchvision batched_nms, grid_sample, cv2.BFMatcher)",
                                "seconds_per_pair": sec, "backbone_seconds": sum(s[1] for s in secs) / len(secs),
                                "matches": secs[0][2], "keypoints": secs[0][3]}
        # box_nms is effectively serial on the CPU (torchvision's greedy scan): SURVEY 8d asks for it at 1 thread too
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import reference_port as rp
        with torch.no_grad():
            lo, _ = net_cpu.backbone_outputs({'image': torch.from_numpy(pair['optical']['image']),
                                              'is_optical': torch.from_numpy(pair['optical']['is_optical'])})
            prob_cpu = rp.detector_head(lo)
        nms_sec = {}
        for nthr in (1, threads):
            torch.set_num_threads(nthr)
            t0 = time.perf_counter()
            kept = rp.box_nms(prob_cpu, 4, 0.015, keep_top_k=args.topk)
            nms_sec[nthr] = time.perf_counter() - t0
        torch.set_num_threads(threads)
        line["cpu_baseline"]["box_nms_threads_1"] = {"seconds_per_image": round(nms_sec[1], 4), "candidates": int((prob_cpu > 0.015).sum()),
                                                     "kept": int((kept > 0).sum())}
        line["cpu_baseline"]["box_nms_threads_all"] = {"seconds_per_image": round(nms_sec[threads], 4), "threads": threads}
        line["cpu_baseline"]["seconds_per_pair_with_1_thread_nms"] = round(sec - 2 * (nms_sec[threads] - nms_sec[1]), 4)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
