"""Multi-GPU test of the in-library NCCL merge (needs >= 2 GPUs; skipped on the single-GPU boxes)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_nccl_merge_matches_single_grid():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_sharded.py")],
                       capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
