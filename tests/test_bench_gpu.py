"""bench.py contract on the GPU (short run) and size edge cases of the C ABI."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_json_line_has_every_contract_key():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-extras",
                        "--no-cpu-baseline"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["gpu_launches"] == 3 * 48 and d["scaling"] == "weak"
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["algorithmic_bytes_per_launch"] == 3294 * 65536
    assert d["e2e"]["h2d_bytes_per_step"] == 65536 * 64 and d["e2e"]["value"] > 0
    assert d["e2e"]["per_ply_round_trip"]["h2d_bytes_per_step"] == 65536 * 64 * 48
    assert d["value"] > 1e8


def test_bench_mcts_blocks_are_first_class():
    """The MCTS half of the metric: configs[2], [3], [4] blocks each carry sims/s over full episodes, a roofline and an
    end-to-end figure through host buffers (configs[4] shrunk to 32,768 games to keep the test short)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu-baseline",
                        "--games-1m", "32768"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["sustained"]["seconds"] >= 2.0 and d["sustained"]["clocks"]["samples"] >= 10
    for name in ("mcts_uniform", "mcts_nn", "mcts_nn_1m"):
        b = d[name]
        assert b["sims_per_sec"] > 1e6 and b["unit"] == "sims/s", name
        assert b["roofline"]["bound"] == "hbm" and 0 < b["roofline"]["frac"] < 1.2, name
        assert b["e2e"]["value"] > 0 and b["e2e"]["d2h_bytes_per_step"] > 0, name
    assert d["mcts_uniform"]["general_path"]["sims_per_sec"] > 1e6
    assert 0 < d["mcts_nn"]["nn_share_of_time"] < 1 and d["mcts_nn"]["roofline_forward"]["bound"] == "tensor"
    assert d["mcts_nn_1m"]["games"] == 32768 and sum(d["mcts_nn_1m"]["outcomes_p1_p2_draw"]) == 32768
    assert d["dropin_mcts"]["sims_per_sec"] > 0


def test_size_edge_cases():
    from nypc_yacht_auction_b200 import _lib
    from nypc_yacht_auction_b200.engine import BatchedYacht
    lib = _lib.load()
    # n = 0 is a no-op for every entry point that takes a count
    z = torch.zeros(16, dtype=torch.int32, device="cuda")
    s = _lib.current_stream()
    assert lib.ya_valid_moves(_lib.ptr(z), 0, _lib.ptr(z), _lib.ptr(z), 0, s) == 0
    assert lib.ya_play_ply(_lib.ptr(z), 0, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), None, None, 0, 0, 0, 1, s) == 0
    assert lib.ya_enumerate_scores(_lib.ptr(z), 0, _lib.ptr(z), _lib.ptr(z), 0, s) == 0
    # misaligned mask buffer is refused, not mis-written
    env = BatchedYacht(8)
    buf = torch.zeros(8 * 3226 + 16, dtype=torch.uint8, device="cuda")
    assert lib.ya_valid_moves(_lib.ptr(env.states), 8, _lib.ptr(env.players), buf.data_ptr() + 1, 8, s) != 0
    # BASELINE.json configs[4]: 1,048,576 concurrent games in one batch (32 MB of state, 3.4 GB of masks)
    n = 1 << 20
    big = BatchedYacht(n, seed=3)
    masks = torch.empty((n, 3226), dtype=torch.uint8, device="cuda")
    for ply in range(6):
        acts, _ = big.play_ply(masks=masks, auto_reset=False)
    cnt = masks[::4099].sum(dim=1, dtype=torch.int32)                # ply 5 is the second score ply of round 2
    assert int(cnt.min()) == 3024 and int(cnt.max()) == 3024
    assert int(big.err_flag.item()) == 0
    small = BatchedYacht(64, seed=3, game_base=n - 64)               # shard invariance at the far end
    for ply in range(6):
        small.play_ply(masks=None, auto_reset=False)
    assert torch.equal(small.states, big.states[:, n - 64:, :])
