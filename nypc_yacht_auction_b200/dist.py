"""Game-index data parallelism: one process per GPU, every rank owns a contiguous
range of global game ids, Philox streams are keyed by the global id, weights are replicated.  There
is NO collective on the stepping / MCTS path; torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests) is used only off the hot path: all-gather of the training examples after self-play
(what Coach.learn appends per episode, Coach.py:91-98) and broadcast of the network weights after
the accept/reject step (Coach.py:130-139).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_games, rank, world):
    """Contiguous, balanced [first, last) of global game ids for `rank`."""
    base, rem = divmod(int(total_games), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def owner_of(game_id, total_games, world):
    base, rem = divmod(int(total_games), int(world))
    cut = rem * (base + 1)
    return game_id // (base + 1) if game_id < cut else rem + (game_id - cut) // max(base, 1)


def allgather_examples(examples, group=None):
    """examples: dict of tensors whose dim 1 is the local game index (BatchedSelfPlay.execute_episodes).
    Returns the same dict with dim 1 = all games of all ranks, ordered by global game id.  Ranks may
    own different numbers of games (padded to the maximum for the collective, trimmed after)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return examples
    world = dist.get_world_size(group)
    out = {}
    any_t = next(iter(examples.values()))
    n_local = torch.tensor([any_t.shape[1] if any_t.dim() > 1 else any_t.shape[0]], device=any_t.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes)
    for name, t in examples.items():
        game_dim = 1 if t.dim() > 1 else 0
        pad_shape = list(t.shape)
        pad_shape[game_dim] = n_max
        buf = t.new_zeros(pad_shape)
        buf.narrow(game_dim, 0, t.shape[game_dim]).copy_(t)
        # NCCL has no int16 / bool collectives: every tensor travels as raw bytes and is viewed back
        wire = buf.contiguous().view(torch.uint8) if buf.dim() > 0 else buf
        parts = [torch.empty_like(wire) for _ in range(world)]
        dist.all_gather(parts, wire, group=group)
        parts = [p.view(t.dtype).view(pad_shape) for p in parts]
        out[name] = torch.cat([p.narrow(game_dim, 0, s) for p, s in zip(parts, sizes)], dim=game_dim)
    return out


def broadcast_weights(module, src=0, group=None):
    """Replicate the evaluator's weights from rank `src` (after Coach.learn accepts a new net)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
