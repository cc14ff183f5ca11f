"""World-size-2 CPU (gloo) tests of the multi-GPU host logic: round-robin sharding of independent
units, rank-0 homography sampling + broadcast, counter / accumulator reductions.  The data path
itself has no collective; the kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from multipoint_b200 import parallel, utils
    r, lr, w = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    # 1. pairs shard round-robin, cover everything exactly once
    mine = parallel.shard_indices(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    assert sorted(sum(gathered, [])) == list(range(11))
    # 2. rank 0 samples the homographies (numpy global RNG), every rank receives the same bits
    cfg = utils._check_ha_config(dict(num=5, erosion_radius=2))

    def sample():
        np.random.seed(7)
        return utils.sample_adaptation_homographies((32, 40), cfg)

    Hs, masks = parallel.broadcast_homographies(sample)
    np.random.seed(7)
    want_H, want_m = utils.sample_adaptation_homographies((32, 40), cfg)
    np.testing.assert_array_equal(Hs, want_H)
    np.testing.assert_array_equal(masks, want_m)
    # 2b. without masks only the 3x3 matrices travel (each rank rasters its own share of the masks on its GPU)
    def sample_no_masks():
        np.random.seed(7)
        return utils.sample_adaptation_homographies((32, 40), cfg, with_masks=False)

    Hs2, none = parallel.broadcast_homographies(sample_no_masks)
    assert none is None
    np.testing.assert_array_equal(Hs2, want_H)
    # 3. the two accumulators of sharded adaptation: per-rank partial sums all-reduce to the total
    rng = np.random.default_rng(3)
    contrib = rng.random((4, 2, 8, 8)).astype(np.float32)       # one term per sampled homography
    part = torch.from_numpy(contrib[parallel.shard_indices(4, rank, world)].sum(0))
    parallel.all_reduce_sum_(part)
    np.testing.assert_allclose(part.numpy(), contrib.sum(0), rtol=1e-6)
    shard = parallel.adaptation_shard()
    assert shard[0] == rank and shard[1] == world and shard[2] is parallel.all_reduce_sum_
    # 4. metric counters and max-over-ranks timing
    tot = parallel.reduce_counters({'matches': 10 + rank, 'keypoints': 100 * (rank + 1)}, device="cpu")
    assert tot == {'keypoints': 300.0, 'matches': 21.0}
    assert parallel.all_reduce_max_float(1.5 + rank, device="cpu") == 2.5
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5)[0] for _ in range(2)) == [0, 1]


def test_single_process_defaults():
    from multipoint_b200 import parallel
    assert parallel.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.shard_indices(5, 1, 2) == [1, 3]
    assert parallel.adaptation_shard() is None
    t = torch.ones(3)
    assert parallel.all_reduce_sum_(t) is t
    assert parallel.reduce_counters({'a': 2}, device="cpu") == {'a': 2.0}
