"""Shard invariance over real GPUs: launches tests/dist_gpu_check.py under torchrun (NCCL, one process per GPU) and
requires its "dist ok" line.  Skips on boxes with fewer than 2 GPUs.  The log is kept under gpurun_out/ when that
directory exists (evidence for profiles/)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_gpu_sharded_self_play_equals_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out = res.stdout + res.stderr
    log_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log_dir):
        with open(os.path.join(log_dir, "dist_gpu_check.log"), "w") as f:
            f.write(out)
    assert res.returncode == 0, out[-4000:]
    assert "dist ok" in out, out[-4000:]
