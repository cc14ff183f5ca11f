#!/usr/bin/env python
"""Multi-GPU functional check, one process per GPU (NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

1. image pairs sharded round-robin over the ranks, no data-path collective, match counters reduced;
2. homographic adaptation with the sampled homographies split over the ranks: rank 0 samples
   (numpy RNG), broadcast, partial accumulators all-reduced (NCCL), fused finish -- compared with
   the single-GPU fused result on every rank.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from multipoint_b200 import ops, parallel, utils  # noqa: E402
from multipoint_b200 import synthetic as syn  # noqa: E402
from multipoint_b200.pipeline import KeypointPipeline  # noqa: E402


def main():
    rank, local_rank, world = parallel.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.backends.cudnn.allow_tf32 = False
    out = {"world": world}

    # 1. pairs are independent units
    n_pairs, H, W, D = 6, 128, 160, 64
    mine = parallel.shard_indices(n_pairs, rank, world)
    pipe = KeypointPipeline(None, nms=4, detection_threshold=0.015, topk=256)
    matches = kps = 0
    for i in mine:
        lg = torch.from_numpy(syn.logits(100 + i, 2, H // 8, W // 8)).to(dev)
        raw = torch.from_numpy(syn.descriptor_map(200 + i, 2, D, H // 8, W // 8)).to(dev)
        ext = pipe.extract_from_backbone(lg, raw, H, W)
        m = pipe.match({k: v[:1] for k, v in ext.items()}, {k: v[1:] for k, v in ext.items()})
        matches += int(m['counts'].sum())
        kps += int(ext['counts'].sum())
    tot = parallel.reduce_counters({'matches': matches, 'keypoints': kps, 'pairs': len(mine)}, device=dev)
    out["pairs"] = tot
    assert tot['pairs'] == n_pairs

    # 2. sharded homographic adaptation vs the fused single-GPU result
    conv = torch.nn.Conv2d(1, 65, 8, stride=8).to(dev)
    torch.manual_seed(7)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(65, 1, 8, 8) * 1.5)
        conv.bias.copy_(torch.randn(65) * 0.5)
        conv.bias[64] += 2.0

    def net(data):
        with torch.no_grad():
            return {'prob': ops.detector_head(conv(data['image']).float())}

    batch = syn.image_pair_batch(71, 2, 64, 80)
    data = {s: {'image': torch.from_numpy(batch[s]['image']).to(dev), 'is_optical': torch.from_numpy(batch[s]['is_optical']).to(dev)}
            for s in ('optical', 'thermal')}
    cfg = dict(num=12, aggregation='prod', erosion_radius=3, min_count=2)
    full_cfg = utils._check_ha_config(cfg)

    def sample():
        np.random.seed(5)
        return utils.sample_adaptation_homographies((64, 80), full_cfg, with_masks=False)

    Hs, masks = parallel.broadcast_homographies(sample, device=dev)
    sharded = utils.homographic_adaptation_multispectral(data, net, cfg, homographies=Hs, masks=masks, shard=parallel.adaptation_shard())
    fused = utils.homographic_adaptation_multispectral(data, net, cfg, homographies=Hs, masks=masks)
    diff = float((sharded - fused).abs().max())
    out["adaptation_max_abs_diff"] = diff
    out["adaptation_max"] = float(fused.max())
    assert diff < 1e-4 * max(1.0, float(fused.max())), diff
    # every rank ends with the same aggregated heatmap
    chk = sharded.double().sum().reshape(1)
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    assert all(abs(float(g) - float(gathered[0])) < 1e-9 for g in gathered)
    if rank == 0:
        print("MULTI_GPU_CHECK PASS " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
