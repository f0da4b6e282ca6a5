"""N>1 on real GPUs: launches tests/multirank_worker.py under torchrun with one rank per visible GPU (2, 4 or 8).
Skipped on a single-GPU box (the two-rank host logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_multirank_sigma_and_davidson():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "multirank_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stderr[-4000:]
    assert res.stdout.count(" ok") == 3
