"""torchrun worker: sharded circuit on N GPUs (NCCL) vs the single-GPU engine on rank 0."""
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
from unitair_b200 import circuit, sharded  # noqa: E402

circuit.SMALL_STATE_AMPS = 0      # small test states: device gates still take the register-blocked pass kernel


def main():
    n = int(sys.argv[1])
    layers = int(sys.argv[2])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["p2p", "nccl"]
    import itertools
    for (dtype, npc, tol), mode in itertools.product(
            ((torch.complex64, np.complex64, 2e-5), (torch.complex128, np.complex128, 1e-11)), modes):
        gates = [(qs, torch.as_tensor(u.astype(npc)).to(dev)) for qs, u in random_circuit(n, layers, 11)]
        st = sharded.ShardedState.zero_state(n, dtype, dev)
        sc = sharded.ShardedCircuit(gates, n, dtype, world, exchange=mode)
        if mode == "p2p":
            assert sc.num_fused_swaps > 0, "the peer-memory exchange path was not planned"
        else:
            assert sc.num_fused_swaps == 0
        sc.run(st)
        sc.run(st)                                   # replayable plan
        nrm = float(st.norm_squared())
        # diagonal ops on the sharded state: <Z_0> and a phase layer, no communication
        nl = n - (world.bit_length() - 1)
        idx = torch.arange(1 << nl, device=dev) + st.local_index_offset()
        z0 = torch.where((idx >> (n - 1)) & 1 == 0, 1.0, -1.0).to(torch.float64 if dtype == torch.complex128 else torch.float32)
        ez = float(st.diag_expectation_value(z0))
        ang = (idx % 97).to(z0.dtype) * 0.01
        st.apply_phase(ang)
        got = st.gather_logical()
        if rank == 0:
            ref = ua.unit_vector(0, num_qubits=n, device=dev, dtype=dtype)
            for _ in range(2):
                ref = ua.circuit.apply_gates(gates, ref)
            full_idx = torch.arange(1 << n, device=dev)
            zf = torch.where((full_idx >> (n - 1)) & 1 == 0, 1.0, -1.0).to(z0.dtype)
            ez_ref = float(ua.diag_expectation_value(zf, ref))
            assert abs(ez - ez_ref) < 1e-4, (ez, ez_ref)
            ref = ua.simulation.apply_phase((full_idx % 97).to(z0.dtype) * 0.01, ref)
            err = float((got - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
            assert err < tol, f"sharded vs single-GPU rel err {err}"
            assert abs(nrm - 1.0) < 1e-4, nrm
            print(f"OK dtype={dtype} mode={mode} world={world} n={n} err={err:.2e} swaps={sc.num_swaps} "
                  f"fused_swaps={sc.num_fused_swaps} passes={sc.num_passes}")
        st.release_peers()
    # host-resident shards through the pipelined stream: 5 jobs with different inputs back to back
    # (upload / circuit / download of neighbouring jobs overlap), each against the single-GPU engine
    dtype, tol = torch.complex64, 2e-5
    g = world.bit_length() - 1
    shard = 1 << (n - g)
    h_gates = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in random_circuit(n, layers, 12)]
    gen = torch.Generator().manual_seed(5)
    fulls = []
    for j in range(5):
        f = torch.randn(1 << n, 2, generator=gen)
        fulls.append(torch.view_as_complex(f / f.norm()))
    h_ins = [f[rank * shard:(rank + 1) * shard].clone().pin_memory() for f in fulls]
    h_outs = [torch.zeros(shard, dtype=dtype).pin_memory() for _ in fulls]
    hs = ua.ShardedHostStream(n, dtype, dev, exchange=modes[0])
    plans = [hs.submit(h_gates, hi, ho) for hi, ho in zip(h_ins, h_outs)]
    hs.drain()
    for j, (f, ho, plan) in enumerate(zip(fulls, h_outs, plans)):
        view = sharded.ShardedState(ho.to(dev), n, layout=plan.end_layout)
        got = view.gather_logical()
        if rank == 0:
            ref = ua.circuit.apply_gates(h_gates, f.to(dev))
            err = float((got - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
            assert err < tol, f"host stream job {j}: rel err {err}"
    if rank == 0:
        print(f"OK host stream world={world} n={n} jobs={len(fulls)}")
    hs.state.release_peers()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
