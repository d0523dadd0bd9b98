"""CPU checks of the host side of the leaf evaluator: the weight-image layout of the CTA-pair forward kernel
(mcts.FusedYachtEvaluator.swizzled_image / pair_image) and the wave plan of the configs[4] block."""
import torch

from nypc_yacht_auction_b200.mcts import FusedYachtEvaluator
from nypc_yacht_auction_b200.mcts_bench import wave_plan


class _Host:                                               # the two image builders need nothing but the operand dtype
    op_dtype = torch.float16
    swizzled_image = FusedYachtEvaluator.swizzled_image
    pair_image = FusedYachtEvaluator.pair_image


def _unswizzle(img, rows, k):
    """Inverse of swizzled_image: K-blocks of [rows][64 x 16 bit], 16-byte chunk c of row r stored at chunk c ^ (r & 7)."""
    t = img.view(torch.float16).view(k // 64, rows, 8, 8)
    out = torch.empty((rows, k), dtype=torch.float16)
    for kb in range(k // 64):
        for r in range(rows):
            for c in range(8):
                out[r, kb * 64 + c * 8:kb * 64 + c * 8 + 8] = t[kb, r, c ^ (r & 7)]
    return out


def test_swizzled_image_is_the_128_byte_swizzle():
    w = torch.randn(16, 128)
    img = _Host().swizzled_image(w)
    assert img.dtype == torch.uint8 and img.numel() == 16 * 128 * 2
    assert torch.equal(_unswizzle(img, 16, 128), w.to(torch.float16))


def test_pair_image_gives_each_cta_half_of_every_n_block():
    """ya_k_forward: rank r of a CTA pair holds output columns [64 r, 64 r + 64) of each 128-column block of a trunk layer
    (its share of the layer contiguous: [rank][block][K-block][64 rows x 128 B]), rows [128 r, 128 r + 128) of the input
    layer, and of a policy tile (tile-major layout) its 64 columns."""
    h = _Host()
    w = torch.randn(256, 256)
    img = h.pair_image(w, 128)
    assert img.numel() == 256 * 256 * 2
    half = 64 * 256 * 2                                    # one (rank, block) image: 32 KB
    for rank in range(2):
        for block in range(2):
            part = img[(rank * 2 + block) * half:(rank * 2 + block + 1) * half]
            rows = w[block * 128 + rank * 64:block * 128 + rank * 64 + 64]
            assert torch.equal(_unswizzle(part, 64, 256), rows.to(torch.float16))
    w_in = torch.randn(256, 64)
    img = h.pair_image(w_in, 256)
    for rank in range(2):
        assert torch.equal(_unswizzle(img[rank * 16384:(rank + 1) * 16384], 128, 64), w_in[rank * 128:(rank + 1) * 128].to(torch.float16))
    w_pi = torch.randn(3 * 128, 256)
    img = h.pair_image(w_pi, 128, tile_major=True)
    for tile in range(3):
        for rank in range(2):
            part = img[(tile * 2 + rank) * half:(tile * 2 + rank + 1) * half]
            assert torch.equal(_unswizzle(part, 64, 256), w_pi[tile * 128 + rank * 64:tile * 128 + rank * 64 + 64].to(torch.float16))


def test_wave_plan():
    cap = 2 * 148 * 128
    assert wave_plan(1 << 20, cap) == (28, 37504)          # one GPU: 28 waves, 1.4 % padding
    assert wave_plan(1 << 17, cap) == (4, 32768)           # one of eight GPUs: four full waves
    assert wave_plan(1 << 19, cap) == (14, 37504)
    assert wave_plan(100, cap) == (1, 128)
    for games in (1, 127, 128, 129, 37888, 37889, 75776, 1000003):
        n, wave = wave_plan(games, cap)
        assert wave % 128 == 0 and wave <= cap and n * wave >= games and (n - 1) * wave < games


def test_wave_plans_of_the_shards_cover_every_game_once():
    """configs[4] strong scaling: for every GPU count the shards (dist.shard_range) and their waves (wave_plan) tile the
    global game ids 0 .. total - 1 exactly once -- the games a wave plays are first_game + w * wave + slot for the live
    slots, which is what keys the Philox streams, so the set of games does not depend on the GPU count."""
    from nypc_yacht_auction_b200.dist import shard_range
    total, cap = 1 << 20, 2 * 148 * 128
    for world in (1, 2, 4, 8):
        covered = 0
        for rank in range(world):
            first, last = shard_range(total, rank, world)
            assert first == covered
            n_waves, wave = wave_plan(last - first, cap)
            live = [min(wave, last - first - w * wave) for w in range(n_waves)]
            assert all(x > 0 for x in live) and sum(live) == last - first
            covered = last
        assert covered == total
