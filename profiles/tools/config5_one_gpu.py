"""BASELINE.json configs[4] on ONE GPU's share: 131,072 games, numMCTSSims=100, random-init YachtNNet, played as
waves on one tree pool (coach.self_play_in_waves).  python config5_one_gpu.py [games] [wave_games]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from nypc_yacht_auction_b200.coach import self_play_in_waves
from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
from nypc_yacht_auction_b200.nnet import YachtPolicyValueNet
games = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = YachtPolicyValueNet().to(dev).eval()
ev = FusedYachtEvaluator(net, wave)
seen = []
torch.cuda.synchronize()
t0 = time.perf_counter()
p1, p2, dr = self_play_in_waves(games, wave, 100, ev, seed=0, device=dev, record_examples=True,
                                on_wave=lambda w, ex: seen.append(int(ex["counts"].sum().item())))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
assert p1 + p2 + dr == games
print("config5 share: %d games in waves of %d, 100 sims: %.2f s wall -> %.3e sims/s, %.3e game steps/s; p1 %d p2 %d draws %d; "
      "visit counts recorded per wave %s" % (games, wave, dt, games * 48 * 100 / dt, games * 48 / dt, p1, p2, dr, seen[:2]))
