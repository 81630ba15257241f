"""2-GPU NCCL run of the sharded circuits (skipped unless >= 2 CUDA devices):
the gathered output of the sharded circuit must equal the single-GPU output."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("query", ["q3", "q4", "q7"])
def test_sharded_nccl_equals_single(query, tmp_path):
    script = os.path.join(ROOT, "tests", "multigpu_worker.py")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29517", script, query],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
