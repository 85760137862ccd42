"""Real multi-process / multi-GPU run of the sharded encode (NCCL): skipped on boxes with one GPU."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_two_ranks_nccl_sharded_plain_and_huffman(gpu):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(ROOT / "tests" / "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=str(ROOT))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    assert "MGPU_RESULT OK" in out, out[-2000:]
