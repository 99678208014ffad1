"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the shard / all-gather layout
bench.py and a multi-GPU caller use (NCCL on the GPUs). The 'model' here is a deterministic per-clip
function, so the gathered result must equal the single-process result row for row."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from syntalker_b200 import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def per_clip(x):                         # stands in for sample -> decode -> 330-d on one rank's clips
    return torch.stack([x.sum(dim=(1, 2)), (x ** 2).sum(dim=(1, 2))], dim=1).unsqueeze(1).expand(-1, 4, -1).contiguous()


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(B, 3, 5, generator=g)
    local = per_clip(sharding.shard(full, rank, world))
    out = sharding.gather_rows(local, B)
    ok = torch.equal(out, per_clip(full))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(float(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 8), (2, 7), (3, 8)])
def test_shard_and_gather(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0


def test_shard_ranges_cover_the_batch():
    for B in (1, 7, 32, 256):
        for W in (1, 2, 3, 8):
            rows = []
            for r in range(W):
                lo, hi = sharding.shard_range(B, r, W)
                assert 0 <= lo <= hi <= B and (hi - lo) in (B // W, B // W + 1)
                rows += list(range(lo, hi))
            assert rows == list(range(B))
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)
