"""bench.py contract on CPU: the reference arm prints one JSON line with the agreed keys; non-zero ranks
stay silent.  (The GPU arm's line is checked by tests/test_bench_gpu.py.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=300)


def test_reference_arm_json_line():
    r = _run(None, "--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "game_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    # the unmodified reference when it is on this machine (baseline/_ref or /root/reference), else the oracle port
    import bench
    assert d["cpu_baseline"]["kind"] == ("reference" if bench.reference_dir() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--steps", "1", "--warmup", "0",
             "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_unmodified_reference_timing_driver():
    """baseline/run_reference.py drives the reference's own Arena / MCTS / Coach and reports rates."""
    import bench
    ref = bench.reference_dir()
    if ref is None:
        import pytest
        pytest.skip("the reference is not on this machine")
    a = bench._ref_result(bench._ref_proc(ref, "arena", 8))
    assert a["games"] == 8 and a["steps"] == 8 * 48 and a["steps_per_s"] > 0 and a["p1"] + a["p2"] + a["draws"] == 8
    m = bench._ref_result(bench._ref_proc(ref, "mcts", 5, 1))
    assert m["plies"] == 48 and m["sims"] == 48 * 5 and m["sims_per_s"] > 0
