"""torchrun worker: config C3 (ry/rz + CNOT ladder, loss = sum_b <Z_0>) with the batch split over
the GPUs of the box -- no communication on the data path, ONE all_reduce of theta.grad -- against
the same computation on a single GPU.  argv: qubits, batch, layers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
import unitair_b200 as ua  # noqa: E402


def loss_fn(theta, st, n, layers, assume_unitary):
    cn = ua.gates.cnot(device=st.device, dtype=torch.complex64)
    z0 = torch.where((torch.arange(2 ** n, device=st.device) >> (n - 1)) & 1 == 0, 1.0, -1.0).to(torch.float32)
    gl = []
    for l in range(layers):
        for q in range(n):
            gl.append(([q], ua.gates.exp_y(theta[l, q, 0])))
            gl.append(([q], ua.gates.exp_z(theta[l, q, 1])))
        for q in range(n - 1):
            gl.append(([q, q + 1], cn))
    psi = ua.circuit.apply_gates(gl, st, assume_unitary=assume_unitary)
    ez = ua.diag_expectation_value(z0, psi)
    return ez, ez.sum()


def main():
    n, B, layers = (int(x) for x in sys.argv[1:4])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    rng = np.random.default_rng(33)
    s = rng.standard_normal((B, 2 ** n)) + 1j * rng.standard_normal((B, 2 ** n))
    s /= np.linalg.norm(s, axis=-1, keepdims=True)
    states = torch.from_numpy(s.astype(np.complex64)).to(dev)
    theta0 = torch.from_numpy((rng.random((layers, n, 2)) * 2 * np.pi).astype(np.float32)).to(dev)
    for assume_unitary in (True, False):
        theta = theta0.clone().requires_grad_(True)
        ez, loss = loss_fn(theta, ua.shard_batch(states), n, layers, assume_unitary)
        loss.backward()
        ua.all_reduce_gradients([theta])
        ez_all = ua.batch.gather_batch(ez.detach(), B)
        if rank == 0:
            ref_theta = theta0.clone().requires_grad_(True)
            ez_ref, loss_ref = loss_fn(ref_theta, states, n, layers, assume_unitary)
            loss_ref.backward()
            err = float((theta.grad - ref_theta.grad).norm() / ref_theta.grad.norm())
            assert err < 1e-5, f"theta gradient, batch over {world} GPUs vs one: rel err {err:.2e}"
            assert torch.allclose(ez_all, ez_ref.detach(), rtol=0, atol=2e-6)
            print(f"OK batch-sharded C3 world={world} n={n} B={B} assume_unitary={assume_unitary} grad err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
