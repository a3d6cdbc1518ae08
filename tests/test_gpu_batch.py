"""Batch-dimension sharding on real GPUs (needs >= 2, skipped otherwise): config C3 split over
the GPUs reproduces the single-GPU theta gradient with one all_reduce."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,n,batch,layers", [(2, 10, 24, 4), (8, 12, 100, 5)])
def test_c3_batch_split_over_gpus(world, n, batch, layers):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29800 + world),
           os.path.join(ROOT, "tests", "_batch_gpu_worker.py"), str(n), str(batch), str(layers)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("OK batch-sharded C3") == 2, out.stdout[-2000:]
