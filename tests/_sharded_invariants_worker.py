"""torchrun worker: invariants of the sharded engine at sizes no single-GPU oracle can hold
(hard part 7): a circuit followed by its inverse is the identity; the GHZ circuit
leaves exactly two amplitudes of 1/sqrt(2).  argv: total qubits, layers."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
from bench import random_circuit  # noqa: E402
import unitair_b200 as ua  # noqa: E402
from unitair_b200 import sharded  # noqa: E402


def main():
    n = int(sys.argv[1])
    layers = int(sys.argv[2])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    nl = n - (world.bit_length() - 1)
    for mode in ("p2p", "nccl"):
        # ---- circuit . circuit^-1 = identity ------------------------------------------------
        gates_np = random_circuit(n, layers, 36)
        fwd = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in gates_np]
        inv = [(qs, torch.as_tensor(np.ascontiguousarray(u.conj().T).astype(np.complex64)).to(dev))
               for qs, u in reversed(gates_np)]
        st = sharded.ShardedState.zero_state(n, torch.complex64, dev)
        sc = sharded.ShardedCircuit(fwd + inv, n, torch.complex64, world, restore=True, exchange=mode)
        half = sharded.ShardedCircuit(fwd, n, torch.complex64, world, restore=True, exchange=mode)
        half.run(st)
        spread = float(st.local[0].abs()) if rank == 0 else 0.0       # the circuit must move the state
        nrm_mid = float(st.norm_squared())
        st.release_peers()
        st.local = st.spare = None           # 36 qubits on 8 GPUs: 64 GiB + 64 GiB spare per state
        del half
        torch.cuda.empty_cache()
        dist.barrier()
        st2 = sharded.ShardedState.zero_state(n, torch.complex64, dev)
        sc.run(st2)
        nrm = float(st2.norm_squared())
        a0 = st2.local[0].clone() if rank == 0 else torch.zeros((), dtype=torch.complex64, device=dev)
        if rank == 0:
            st2.local[0] = 0
        rest = float(st2.norm_squared())
        if rank == 0:
            assert abs(nrm_mid - 1) < 1e-4 and abs(nrm - 1) < 1e-4, (nrm_mid, nrm)
            assert spread < 0.5, spread
            assert abs(float(a0.real) - 1) < 1e-4 and abs(float(a0.imag)) < 1e-4, a0
            assert rest < 1e-7, rest
            print(f"OK inverse mode={mode} world={world} n={n} swaps={sc.num_swaps} "
                  f"fused_swaps={sc.num_fused_swaps} |psi-e0|^2={rest:.2e}")
        st2.release_peers()
        st2.local = st2.spare = None
        del st, st2, sc
        torch.cuda.empty_cache()
        dist.barrier()
        # ---- GHZ: H on qubit 0, CNOT(q, q+1) chain ------------------------------------------
        h = ua.gates.hadamard(device=dev, dtype=torch.complex64)
        cn = ua.gates.cnot(device=dev, dtype=torch.complex64)
        ghz = [([0], h)] + [([q, q + 1], cn) for q in range(n - 1)]
        st = sharded.ShardedState.zero_state(n, torch.complex64, dev)
        sharded.ShardedCircuit(ghz, n, torch.complex64, world, restore=True, exchange=mode).run(st)
        first = st.local[0].clone()
        last = st.local[-1].clone()
        ends = torch.stack([first.abs() ** 2 if rank == 0 else torch.zeros_like(first.abs()),
                            last.abs() ** 2 if rank == world - 1 else torch.zeros_like(last.abs())]).double()
        dist.all_reduce(ends)
        nrm = float(st.norm_squared())
        if rank == 0:
            assert abs(float(ends[0]) - 0.5) < 1e-5 and abs(float(ends[1]) - 0.5) < 1e-5, ends
            assert abs(nrm - 1) < 1e-5, nrm
            print(f"OK ghz mode={mode} world={world} n={n}")
        st.release_peers()
        st.local = st.spare = None
        del st
        torch.cuda.empty_cache()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
