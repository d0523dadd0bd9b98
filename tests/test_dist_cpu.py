"""N>1 host logic on CPU: world_size-2 gloo processes (sharding, example all-gather, weight broadcast)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nypc_yacht_auction_b200.dist import shard_range, owner_of, allgather_examples, broadcast_weights


def test_shard_ranges_partition_the_games():
    for total in (1, 7, 8, 65536, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                a, b = shard_range(total, r, world)
                assert a == prev and b >= a
                prev = b
                if b > a:
                    assert owner_of(a, total, world) == r and owner_of(b - 1, total, world) == r
            assert prev == total


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, last = shard_range(total, rank, world)
        n = last - first
        ids = torch.arange(first, last)
        ex = {
            "features": (ids.float().view(1, n, 1) + torch.arange(3).float().view(3, 1, 1)).expand(3, n, 59).contiguous(),
            "counts": ids.int().view(1, n, 1).expand(3, n, 4).contiguous(),
            "actions": ids.to(torch.int16).view(1, n, 1).expand(3, n, 4).contiguous(),      # NCCL has no int16: sent as bytes
            "result_p1": ids.float(),
        }
        full = allgather_examples(ex)
        assert full["features"].shape == (3, total, 59)
        assert torch.equal(full["counts"][0, :, 0], torch.arange(total, dtype=torch.int32))
        assert full["actions"].dtype == torch.int16 and torch.equal(full["actions"][1, :, 3], torch.arange(total, dtype=torch.int16))
        assert torch.equal(full["result_p1"], torch.arange(total, dtype=torch.float32))
        assert torch.equal(full["features"][2, :, 5], torch.arange(total).float() + 2)
        torch.manual_seed(rank)
        net = torch.nn.Linear(4, 3)
        broadcast_weights(net, src=0)
        w = [torch.empty_like(net.weight) for _ in range(world)]
        dist.all_gather(w, net.weight.data)
        assert all(torch.equal(w[0], x) for x in w)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allgather_and_broadcast():
    mp.spawn(_worker, args=(2, _free_port(), 11), nprocs=2, join=True)
