"""Multi-GPU parity (needs >= 2 GPUs on the box, skipped otherwise): the NCCL sharded path
must reproduce the single-GPU engine's state on the same circuit."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "_sharded_gpu_worker.py"), "20", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("OK dtype") == 4, out.stdout[-2000:]
    assert "OK host stream" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("world,total_qubits,layers", [(2, 26, 3), (8, 33, 2), (8, 36, 2)])
def test_sharded_invariants_inverse_and_ghz(world, total_qubits, layers):
    """BASELINE config 5 sizes (33 and 36 qubits on 8 GPUs) have no oracle: circuit . circuit^-1
    must return |0...0> and the GHZ circuit must leave two amplitudes of 1/sqrt(2), with both
    exchange formulations.  Self-skips below the GPU count (and memory) it needs."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    per_gpu = (8 << (total_qubits - (world.bit_length() - 1))) * 2.2       # state + spare + slack
    if torch.cuda.mem_get_info()[0] < per_gpu:
        pytest.skip("not enough device memory")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + total_qubits),
           os.path.join(ROOT, "tests", "_sharded_invariants_worker.py"), str(total_qubits), str(layers)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("OK inverse") == 2 and out.stdout.count("OK ghz") == 2, out.stdout[-2000:]
