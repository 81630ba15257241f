"""Multi-GPU run (one process per GPU, skipped unless >= 2 CUDA devices) of the in-library exchange
(dbsp_shard / dbsp_shard2 / dbsp_gather / dbsp_allreduce_max_u64 over NVLink peer memory) and of the sharded
circuits: the gathered output of the sharded circuit must equal the single-GPU output, bit for bit."""
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
@pytest.mark.parametrize("query", ["ops", "q3", "q4", "q7"])
def test_sharded_equals_single(query, tmp_path):
    script = os.path.join(ROOT, "tests", "multigpu_worker.py")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(_ngpus(), 4)}", "--master-addr", "127.0.0.1",
         "--master-port", "29517", script, query],
        capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert ("OPS_OK" if query == "ops" else "SHARDED_OK") in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
