"""Seeded synthetic inputs shaped like the reference's data.

numpy only, and only IEEE-exact elementwise arithmetic after the RNG draw, so that the same seed
reproduces the same bits wherever it runs: golden fixtures store the seed and a checksum instead
of megabytes of input.  Layouts follow the reference: images ``(B,1,H,W)`` fp32 in [0,1],
``is_optical`` ``(B,1)`` bool, logits ``(B,65,H/8,W/8)``, coarse descriptors ``(B,D,H/8,W/8)``
(multipoint/datasets/ImagePairDataset.py:173-241, multipoint/models/MultiPoint.py:126-135).
"""
import hashlib

import numpy as np


def checksum(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def images(seed, B, H=512, W=640):
    rng = np.random.default_rng(seed)
    return rng.random((B, 1, H, W), dtype=np.float32)


def image_pair_batch(seed, B, H=512, W=640):
    """The nested dict ImagePairDataset yields, as numpy arrays (caller converts to torch)."""
    rng = np.random.default_rng(seed)
    opt = rng.random((B, 1, H, W), dtype=np.float32)
    # the thermal image is a smoothed, inverted view of the optical one plus noise
    th = 1.0 - 0.5 * (opt + np.roll(opt, 1, axis=3)) + 0.05 * rng.random((B, 1, H, W), dtype=np.float32)
    th = np.clip(th, 0.0, 1.0).astype(np.float32)
    ones = np.ones((B, 1, H, W), dtype=bool)
    return {
        'optical': {'image': opt, 'valid_mask': ones, 'is_optical': np.ones((B, 1), dtype=bool)},
        'thermal': {'image': th, 'valid_mask': ones.copy(), 'is_optical': np.zeros((B, 1), dtype=bool)},
    }


def logits(seed, B, Hc=64, Wc=80, sigma=2.0, bias=5.0, quant=None):
    """Detector-head logits: randn*sigma with +bias on the dustbin channel.  (sigma,bias)=(2,5)
    gives ~42 k candidates / 11.8 k NMS survivors per 512x640 image at threshold 0.015."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, 65, Hc, Wc), dtype=np.float32) * np.float32(sigma)
    x[:, 64] += np.float32(bias)
    if quant:
        x = np.round(x * np.float32(quant)) / np.float32(quant)  # tie stress
    return x.astype(np.float32)


def heatmap(seed, B, H=512, W=640, squarings=4, scale=0.25, quant=None, batched=True):
    """A peaky heatmap made with exact fp32 multiplies only: scale * u^(2^squarings).
    squarings=4, scale=0.25 -> ~16 % of pixels above 0.015.  quant=q floors to multiples of
    1/q (q a power of two) to plant exact ties."""
    rng = np.random.default_rng(seed)
    h = rng.random((B, 1, H, W), dtype=np.float32)
    for _ in range(squarings):
        h = h * h
    h = h * np.float32(scale)
    if quant:
        h = np.floor(h * np.float32(quant)) / np.float32(quant)
    h = h.astype(np.float32)
    return h if batched else h[0, 0]


def descriptor_map(seed, B, D, Hc=64, Wc=80):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, D, Hc, Wc), dtype=np.float32)


def keypoints(seed, K, H=512, W=640, corners=True):
    """(K,2) int64 (y,x), unique, row-major sorted like torch.nonzero output."""
    rng = np.random.default_rng(seed)
    flat = rng.choice(H * W, size=K, replace=False)
    if corners and K >= 4:
        flat[:4] = [0, W - 1, (H - 1) * W, H * W - 1]
        flat = np.unique(flat)
    flat = np.sort(flat)
    return np.stack([flat // W, flat % W], axis=1).astype(np.int64)


def descriptor_sets(seed, N1, N2=None, D=256, noise=0.05, duplicates=0):
    """Two unit-norm descriptor sets with planted correspondences: set 2 = a row permutation of
    set 1 (truncated / padded with fresh rows to N2) + noise*randn, renormalised.
    duplicates > 0 copies that many rows inside set 2 to plant exact ties."""
    N2 = N1 if N2 is None else N2
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((N1, D), dtype=np.float32)
    a /= np.sqrt((a.astype(np.float64) ** 2).sum(1, keepdims=True)).astype(np.float32)
    perm = rng.permutation(max(N1, N2))
    base = np.concatenate([a, rng.standard_normal((max(0, N2 - N1), D), dtype=np.float32)], 0) \
        if N2 > N1 else a
    b = base[perm[perm < base.shape[0]][:N2]].copy()
    if b.shape[0] < N2:
        b = np.concatenate([b, rng.standard_normal((N2 - b.shape[0], D), dtype=np.float32)], 0)
    b = b + np.float32(noise) * rng.standard_normal(b.shape, dtype=np.float32)
    b /= np.sqrt((b.astype(np.float64) ** 2).sum(1, keepdims=True)).astype(np.float32)
    b = b.astype(np.float32)
    for d in range(duplicates):
        src, dst = rng.integers(0, N2, size=2)
        b[dst] = b[src]
    return np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
