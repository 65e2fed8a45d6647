"""N > 1 on real GPUs: the library's NCCL collectives (delta all-reduce, flatness, deltaG, window joins
through an all-gather) with walkers sharded over 2 B200s, against a single-process oracle run.  Needs two
visible GPUs (skipped on the one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_nccl_collectives_match_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multigpu worker ok" in r.stdout
