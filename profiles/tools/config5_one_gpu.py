"""BASELINE.json configs[4]: 1,048,576 games, numMCTSSims=100, random-init YachtNNet, full 48-ply episodes with
example recording.  Every GPU plays its contiguous share as waves on one tree pool (coach.self_play_in_waves).
    python config5_one_gpu.py [games_per_gpu] [wave_games]                              # one GPU's share
    torchrun --nproc-per-node 8 ... config5_one_gpu.py 131072 16384                     # the whole configuration
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.coach import self_play_in_waves
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
games = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                                  # same weights on every rank
net = YachtPolicyValueNet().to(dev).eval()
ev = FusedYachtEvaluator(net, wave)
seen = []
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
p1, p2, dr = self_play_in_waves(games, wave, 100, ev, first_game=rank * games, seed=0, device=dev, record_examples=True,
                                on_wave=lambda w, ex: seen.append(int(ex["counts"].sum().item())))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
assert p1 + p2 + dr == games
if world > 1:
    t = torch.tensor([dt, p1, p2, dr], dtype=torch.float64, device=dev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dt, p1, p2, dr = float(tmax[0]), int(t[1]), int(t[2]), int(t[3])
if rank == 0:
    total = games * world
    print("configs[4]: %d games on %d GPU(s) in waves of %d, 100 sims: %.2f s wall (max over ranks) -> %.3e sims/s, %.3e game steps/s; "
          "p1 %d p2 %d draws %d" % (total, world, wave, dt, total * 48 * 100 / dt, total * 48 / dt, p1, p2, dr))
if world > 1:
    dist.destroy_process_group()
