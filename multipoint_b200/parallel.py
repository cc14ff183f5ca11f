"""One process per GPU (torch.distributed: NCCL on the box, gloo in CPU tests).

The hot path shards by independent units, so the data path has no collective:
  - image pairs: pair i -> rank i mod world; weights replicated;
  - homographic adaptation: the num-1 sampled homographies are drawn once (rank 0's numpy stream,
    broadcast) and split round-robin; the two accumulators are summed with one all-reduce.
Only metric scalars / counters and those two accumulators ever cross NVLink.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend=None):
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_indices(n_items, rank, world):
    """Round-robin assignment of independent units: item i -> rank i mod world."""
    return list(range(rank, n_items, world))


def all_reduce_sum_(t):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_max_float(x, device=None):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or ("cuda" if torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl" else "cpu"))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def reduce_counters(counters, device=None):
    """Sum a dict of python numbers across ranks (match counts, keypoint counts, repeatability
    numerators / denominators).  Returns plain floats on every rank."""
    keys = sorted(counters)
    dev = device or ("cuda" if torch.cuda.is_available() and dist.is_initialized() and dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=dev)
    all_reduce_sum_(t)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def broadcast_homographies(sample_fn, device="cpu"):
    """Rank 0 draws (homographies (n,3,3) f64, masks (n,H,W) u8 or None) with ``sample_fn`` -- keeping
    the reference's numpy RNG stream on one process -- and every rank receives them.  With masks=None
    only the 72 B matrices travel; each rank rasters the masks of its own share on its GPU."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return sample_fn()
    if rank == 0:
        Hs, masks = sample_fn()
        shape = torch.tensor([Hs.shape[0]] + (list(masks.shape[1:]) if masks is not None else [0, 0]), dtype=torch.int64, device=device)
    else:
        shape = torch.zeros(3, dtype=torch.int64, device=device)
    dist.broadcast(shape, 0)
    n, H, W = (int(v) for v in shape.tolist())
    h_t = torch.from_numpy(Hs).to(device) if rank == 0 else torch.zeros((n, 3, 3), dtype=torch.float64, device=device)
    dist.broadcast(h_t, 0)
    if H == 0:
        return h_t.cpu().numpy(), None
    m_t = torch.from_numpy(np.asarray(masks)).to(device) if rank == 0 else torch.zeros((n, H, W), dtype=torch.uint8, device=device)
    dist.broadcast(m_t, 0)
    return h_t.cpu().numpy(), m_t.cpu().numpy()


def adaptation_shard():
    """The ``shard`` argument of utils.homographic_adaptation* for the current process group."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    return dist.get_rank(), dist.get_world_size(), all_reduce_sum_
