"""MCTS self-play throughput lines for bench.py (BASELINE.json configs[2] and configs[3])."""
from __future__ import annotations

import time


def _run_selfplay(torch, dev, n, sims, evaluator, plies, warm_plies, seed, game_base, use_graph, arena_mb=None,
                  dist=None, world=1):
    from .coach import BatchedSelfPlay
    sp = BatchedSelfPlay(n, sims, cpuct=1.5, evaluator=evaluator, temp_threshold=15, seed=seed, game_base=game_base,
                         device=dev, arena_mb_per_game=arena_mb, record_examples=True)
    if use_graph:
        sp.mcts.capture_graph()
    t = 0
    for _ in range(warm_plies):
        sp.play_ply(t)
        t += 1
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(plies):
        sp.play_ply(t)
        t += 1
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
        t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms = float(t_ms.item())
    wall = time.perf_counter() - w0
    sp.mcts.check_errors()
    nodes = sp.mcts.pool.node_counts()
    return {"ms": ms, "wall_s": wall, "sims": world * n * sims * plies, "game_steps": world * n * plies,
            "sims_per_sec": world * n * sims * plies / (ms * 1e-3), "steps_per_sec": world * n * plies / (ms * 1e-3),
            "pool_gb": sp.mcts.pool.bytes() / 1e9, "max_nodes_in_use": int(nodes.max().item()),
            # per move: features, root counts, sparse root policy, pick, transition + the search itself (uniform prior:
            # ONE launch for all simulations; network: select, forward, expand per simulation)
            "launches": plies * (5 + (1 if getattr(evaluator, "uniform", False) else 3 * sims))}


def run(args, torch, dev, rank=0, world=1, dist=None):
    from .mcts import UniformEvaluator, FusedYachtEvaluator
    from .nnet import YachtPolicyValueNet
    out = {}
    # configs[2]: MCTS self-play, numMCTSSims=25, uniform prior (no NN), 4,096 concurrent games
    r = _run_selfplay(torch, dev, 4096, 25, UniformEvaluator(), plies=24, warm_plies=6, seed=args.seed + 1,
                      game_base=rank * 4096, use_graph=True, dist=dist, world=world)
    out["mcts_uniform"] = {
        "workload": "configs[2]: MCTS self-play numMCTSSims=25, uniform prior, 4096 games/GPU, cpuct 1.5, plies 6..29 of the episode",
        "sims_per_sec": r["sims_per_sec"], "game_steps_per_sec": r["steps_per_sec"], "ms": r["ms"], "gpu_launches": r["launches"],
        "pool_gb": r["pool_gb"], "max_nodes_in_use": r["max_nodes_in_use"]}
    # configs[3]: AlphaZero self-play, random-init yacht NNet (H=256, 6 blocks), numMCTSSims=100, 16,384 games
    torch.manual_seed(0)
    net = YachtPolicyValueNet().to(dev)
    ev = FusedYachtEvaluator(net, 16384)
    r = _run_selfplay(torch, dev, 16384, 100, ev, plies=4, warm_plies=4, seed=args.seed + 2, game_base=rank * 16384,
                      use_graph=True, dist=dist, world=world)
    # the leaf evaluator alone: one whole-network forward over 16,384 leaves (CUDA events, 50 launches)
    x = torch.rand((16384, 59), device=dev)
    for _ in range(5):
        ev(x)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(50):
        ev(x)
    f1.record()
    torch.cuda.synchronize(dev)
    fwd_us = f0.elapsed_time(f1) * 1000.0 / 50
    real_flops = 2.0 * YachtPolicyValueNet.num_macs() * 16384
    out["nn_forward"] = {"kernel": "ya_k_forward (tcgen05, bf16 operands, float32 accumulation)", "leaves": 16384,
                         "us_per_launch": fwd_us, "tflops": real_flops / fwd_us * 1e-6,
                         "note": "3.32 MFLOP per leaf (unpadded); 128 CTAs of 128 leaves on 148 SMs"}
    out["mcts_nn"] = {
        "workload": "configs[3]: AlphaZero self-play, random-init YachtNNet (hidden 256, 6 blocks; the whole forward is one tcgen05 kernel, csrc/ya_forward.cu, bf16 operands / float32 accumulation), "
                    "numMCTSSims=100, 16384 games/GPU, one batched forward per simulation wave (3 launches per wave: select, forward, expand), "
                    "softmax+mask fused into the expand kernel, plies 4..7",
        "sims_per_sec": r["sims_per_sec"], "game_steps_per_sec": r["steps_per_sec"], "ms": r["ms"], "gpu_launches": r["launches"],
        "pool_gb": r["pool_gb"], "max_nodes_in_use": r["max_nodes_in_use"],
        "nn_flops_per_leaf": 2 * YachtPolicyValueNet.num_macs()}
    return out
