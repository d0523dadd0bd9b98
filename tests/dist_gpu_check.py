"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs; tests/test_dist_gpu.py launches it from pytest -m gpu
and skips below 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tests/dist_gpu_check.py

Every rank self-plays its shard of the global game ids (no collective on the path), the examples are
all-gathered over NCCL, and rank 0 compares them with a single-process run of ALL games: sharding over
GPUs must not change a single visit count or outcome.  Then the evaluator weights are broadcast.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nypc_yacht_auction_b200.coach import BatchedSelfPlay          # noqa: E402
from nypc_yacht_auction_b200.dist import allgather_examples, broadcast_weights, shard_range   # noqa: E402
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator      # noqa: E402
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet       # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    total, sims, seed = 24, 8, 31
    torch.manual_seed(rank)                                    # different weights per rank until the broadcast
    net = YachtPolicyValueNet().to(dev).eval()
    broadcast_weights(net, src=0)
    w = [torch.empty_like(net.pi_head[2].weight) for _ in range(world)]
    dist.all_gather(w, net.pi_head[2].weight.data)
    assert all(torch.equal(w[0], x) for x in w), "weights differ after broadcast"

    first, last = shard_range(total, rank, world)
    # A. uniform evaluator: the whole path is ours -> bitwise shard invariance
    sp = BatchedSelfPlay(last - first, sims, evaluator=None, seed=seed, game_base=first, device=dev)
    full = allgather_examples(sp.execute_episodes())
    if rank == 0:
        ref = BatchedSelfPlay(total, sims, evaluator=None, seed=seed, game_base=0, device=dev)
        rex = ref.execute_episodes()
        for k in ("features", "actions", "counts", "value", "result_p1"):
            assert torch.equal(full[k], rex[k]), "sharded run differs from the single-process run in %s" % k
    # B. network evaluator: the whole forward is one hand-written kernel whose rows do not depend on the batch they
    # sit in (csrc/ya_forward.cu), so the sharded AlphaZero self-play is bitwise the single-process one too
    sp = BatchedSelfPlay(last - first, sims, evaluator=FusedYachtEvaluator(net, last - first), seed=seed, game_base=first,
                         device=dev)
    full_nn = allgather_examples(sp.execute_episodes())
    assert full_nn["counts"].shape[1] == total and bool((full_nn["result_p1"] != 0).all())
    visits = full_nn["counts"].sum(-1)
    assert int(visits[0].min()) == sims - 1 and int(visits[0].max()) == sims - 1      # fresh roots: numMCTSSims - 1 visits
    if rank == 0:
        ref = BatchedSelfPlay(total, sims, evaluator=FusedYachtEvaluator(net, total), seed=seed, game_base=0, device=dev)
        rex = ref.execute_episodes()
        for k in ("features", "actions", "counts", "value", "result_p1"):
            assert torch.equal(full_nn[k], rex[k]), "sharded network run differs from the single-process run in %s" % k
    # C. the configs[4] shape: every rank plays its shard as TWO waves on one tree pool with the simulation wave replayed
    # as a CUDA graph (global game ids read from device memory); still bitwise the single-process batch
    from nypc_yacht_auction_b200.coach import self_play_in_waves
    per = last - first
    assert per % 2 == 0
    parts = []
    self_play_in_waves(per, per // 2, sims, FusedYachtEvaluator(net, per // 2), first_game=first, use_graph=True, seed=seed,
                       device=dev, on_wave=lambda w, ex: parts.append({k: v.clone() for k, v in ex.items()}))
    mine = {k: torch.cat([p[k] for p in parts], dim=0 if parts[0][k].dim() == 1 else 1) for k in ("features", "actions", "counts", "value", "result_p1")}
    full_w = allgather_examples(mine)
    if rank == 0:
        for k in ("features", "actions", "counts", "value", "result_p1"):
            assert torch.equal(full_w[k], rex[k]), "sharded + waved + graphed network run differs from the single-process run in %s" % k
        print("dist ok: %d games over %d GPUs == 1 process, bitwise, for the uniform AND the network evaluator (eager, and as "
              "2 graphed waves per rank); NCCL all-gather %s, weight broadcast verified" % (total, world, tuple(full["counts"].shape)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
