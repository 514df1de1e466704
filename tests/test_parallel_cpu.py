"""world_size-2 gloo tests (CPU) of the batch-sharding plumbing used for multi-GPU inference."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wave_mamba_b200 import parallel


def test_shard_range_is_a_partition():
    for n in (0, 1, 2, 7, 8, 9):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                b, e = parallel.shard_range(n, r, world)
                assert 0 <= b <= e <= n
                covered += list(range(b, e))
            assert covered == list(range(n))
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch_size, result_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Conv2d(3, 3, 1)
        if rank != 0:
            for p in net.parameters():
                p.data.zero_()
        parallel.broadcast_parameters(net, src=0)
        torch.manual_seed(0)
        ref = torch.nn.Conv2d(3, 3, 1)
        same = all(torch.equal(a, b) for a, b in zip(net.parameters(), ref.parameters()))
        batch = torch.arange(batch_size * 3 * 4 * 4, dtype=torch.float32).view(batch_size, 3, 4, 4) \
            if rank == 0 else None
        out = parallel.sharded_forward(lambda t: net(t) + 0, batch, torch.device("cpu"), src=0)
        ok = True
        if rank == 0:
            ok = torch.allclose(out, ref(batch))
        # uint8 batch (the configs[3] edge format), ragged shards, a shape-preserving map
        u8 = (torch.arange(batch_size * 6 * 5 * 3) % 251).to(torch.uint8).view(batch_size, 6, 5, 3) \
            if rank == 0 else None
        out8 = parallel.sharded_forward(lambda t: 255 - t, u8, torch.device("cpu"), src=0)
        if rank == 0:
            ok = ok and out8.dtype == torch.uint8 and torch.equal(out8, 255 - u8)
        else:
            ok = ok and out8 is None
        result_queue.put((rank, bool(same), bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch_size", [1, 2, 5])
def test_broadcast_and_sharded_forward_world2(batch_size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch_size, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=10) for _ in range(2))
    assert results == [(0, True, True), (1, True, True)]
