"""CPU, world_size 2 over gloo: the batch-sharding host logic of the N>1 path (no collective inside the
loop, one all_gather at the end)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import fidm_b200  # noqa: F401
    from fidm_b200.parallel import gather_batch, shard, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 5                                     # uneven on purpose
        full = torch.arange(B * 3 * 4, dtype=torch.float32).reshape(B, 3, 2, 2)
        mine = shard(full)
        a, b = shard_bounds(B, rank, world)
        assert mine.shape[0] == b - a and torch.equal(mine, full[a:b])
        out = gather_batch(mine * 2, B)           # stand-in for the per-rank sampler output
        assert torch.equal(out, full * 2)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2():
    from fidm_b200.parallel import shard_bounds
    assert [shard_bounds(8, r, 8) for r in range(8)] == [(i, i + 1) for i in range(8)]
    assert [shard_bounds(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(1, r, 2) for r in range(2)] == [(0, 1), (1, 1)]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
