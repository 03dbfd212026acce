"""N > 1 on real GPUs: the row-stripe sharded driver over NCCL equals the single-GPU run.  Needs >= 2
GPUs on the box (skipped otherwise; the host logic is covered on CPU by test_sharded_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nccl_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + world), os.path.join(HERE, "_sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sharded-vs-unsharded" in out.stdout
