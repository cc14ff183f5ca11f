"""Library-level PORT of the reference's CPU path (test / benchmark infrastructure, NOT product).

The reference is pure Python over torch, torchvision and OpenCV and cannot travel to the GPU box
(/root/reference does not exist there), so this module restates its post-backbone chain with the
same library calls the reference makes, each citing the line it follows.  bench.py times it as
the CPU baseline (``cpu_baseline.kind = "port"``) and as ``--impl reference``; tests pin it to the
golden fixtures frozen from the real reference.  Nothing under multipoint_b200/ imports it.
"""
import numpy as np
import torch
from torchvision.ops import nms
from torchvision.ops.boxes import batched_nms


def detector_head(logits):
    """multipoint/models/MultiPoint.py:156-157: Softmax2d, drop the dustbin, PixelShuffle(8)."""
    prob = torch.softmax(logits, dim=1)
    return torch.nn.functional.pixel_shuffle(prob[:, :-1], 8)


def descriptor_head(raw):
    """multipoint/models/MultiPoint.py:163-164."""
    return torch.nn.functional.normalize(raw, p=2, dim=1)


def box_nms(prob, size, min_prob, iou=0.1, keep_top_k=0):
    """multipoint/utils/utils.py:78-122 on a CPU tensor (the shipped configs set cpu_nms: true)."""
    if prob.dim() not in (2, 4):
        raise ValueError('The probability must be either 2D (H,W), or 4D (B, 1, H, W)')
    pts = (prob > min_prob).nonzero()                                   # :97
    scores = prob[tuple(pts.t())]                                       # :98
    half = size * 0.5
    if prob.dim() == 4:
        yx = pts[:, 2:]
        boxes = torch.cat([yx - half, yx + half], dim=1)                # :101
        keep = batched_nms(boxes, scores, pts[:, 0], iou)               # :102-103
        if keep_top_k > 0:                                              # :109-114
            img = pts[keep, 0]
            keep = torch.cat([keep[img == b][:keep_top_k] for b in range(prob.shape[0])])
    else:
        boxes = torch.cat([pts - half, pts + half], dim=1)              # :105
        keep = nms(boxes, scores, iou)                                  # :106
        if keep_top_k > 0:
            keep = keep[:keep_top_k]                                    # :116
    out = torch.zeros_like(prob)                                        # :119
    out[tuple(pts[keep].t())] = scores[keep]                            # :120
    return out


def interpolate_descriptors(keypoints, desc, H, W):
    """multipoint/utils/utils.py:159-167."""
    kp = keypoints.float().clone()
    kp[:, 0] = kp[:, 0] / (float(H) * 0.5) - 1.0
    kp[:, 1] = kp[:, 1] / (float(W) * 0.5) - 1.0
    grid = torch.flip(kp.view(1, 1, -1, 2), [3])
    d = torch.nn.functional.grid_sample(desc.unsqueeze(0), grid, align_corners=True)[0, :, 0, :].transpose(0, 1)
    return torch.nn.functional.normalize(d, p=2, dim=1)


def get_matches_bf_crosscheck(desc_1, desc_2):
    """multipoint/utils/matching.py:7,31 with the shipped method_kwargs {crossCheck: True}."""
    import cv2
    return cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(desc_1, desc_2)


def pair_chain(logits, raw_desc, H, W, nms_size=4, threshold=0.015, topk=2048):
    """What predict_align_image_pair.py:126-190 does after the two forward passes, for one pair:
    logits (2,65,Hc,Wc) and raw descriptor maps (2,D,Hc,Wc) for (optical, thermal) -> matches."""
    prob = detector_head(logits)
    desc = descriptor_head(raw_desc)
    kps, ds = [], []
    for s in range(2):
        p = box_nms(prob[s:s + 1], nms_size, threshold, keep_top_k=topk)          # :127-137 (4-D call)
        kp = torch.nonzero((p.squeeze() > threshold).float())                      # :170-171
        kps.append(kp)
        ds.append(interpolate_descriptors(kp, desc[s], H, W))                      # :182-183
    matches = get_matches_bf_crosscheck(ds[0].numpy(), ds[1].numpy())              # :186-190
    return kps, ds, matches
