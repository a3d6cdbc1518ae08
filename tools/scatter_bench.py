#!/usr/bin/env python
"""Time the scatter (peer-store) pass against a local pass and an NCCL exchange of the same
volume.  Launch with torchrun on 2/4/8 GPUs:

    python -m torch.distributed.run --nproc-per-node N tools/scatter_bench.py [qubits_per_gpu]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
from bench import haar_unitary  # noqa: E402
from unitair_b200 import circuit, sharded  # noqa: E402


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    g = world.bit_length() - 1
    n = nl + g
    rng = np.random.default_rng(5)
    local = torch.randn(1 << nl, dtype=torch.complex64, device=dev)
    st = sharded.ShardedState(local, n)
    peers = st.peer_pointers()
    spare = st._spare()
    out = {"world": world, "qubits_per_gpu": nl}
    for m in range(1, g + 1):
        victims = list(range(nl - m, nl))          # top local bits: contiguous blocks
        ep = sharded.Epoch(incoming=list(range(m)), victims=list(range(m)), rank_bits=list(range(m)),
                           victim_bits=victims)
        a = sharded.exchange_block_id(rank, ep)
        block_bytes = (1 << (nl - m)) * 8
        dst = [peers[spare.data_ptr()][sharded.exchange_peer(rank, ep, b)] + a * block_bytes for b in range(1 << m)]
        dst_local = [spare.data_ptr() + b * block_bytes for b in range(1 << m)]
        for ngates in (0, 2, 6):
            gl = []
            pairs = [(3, 4), (5, 6), (7, 8), (3, 5), (4, 7), (6, 8)]      # six high qubits below the victims
            for i in range(ngates):
                q0, q1 = pairs[i]
                gl.append(([q0 + m, q1 + m], torch.as_tensor(haar_unitary(rng, 4).astype(np.complex64)).to(dev)))
            cc = circuit.CompiledCircuit(gl, nl, torch.complex64, merge=False) if gl else None
            assert cc is None or cc.num_passes == 1
            tail = circuit.ScatterTail(cc, nl, torch.complex64, victims)
            assert ngates == 0 or tail.reused
            t_peer = timeit(lambda: tail.run(local, dst, visit_xor=a))
            t_peer0 = timeit(lambda: tail.run(local, dst, visit_xor=0))
            t_loc = timeit(lambda: tail.run(local, dst_local))
            t_plain = timeit(lambda: cc.run(local, in_place=True)) if cc is not None else None
            out[f"m{m}_gates{ngates}"] = {"scatter_to_peers_ms": round(t_peer, 3),
                                         "scatter_to_peers_no_rotation_ms": round(t_peer0, 3),
                                         "scatter_local_ms": round(t_loc, 3),
                                         "plain_pass_ms": None if t_plain is None else round(t_plain, 3),
                                         "remote_GBs": round((1 - 2.0 ** -m) * 8 * 2 ** nl / t_peer / 1e6, 1)}
        # NCCL exchange of the same blocks

        def nccl_exchange():
            ops = []
            for b in range(1 << m):
                if b == a:
                    continue
                peer = sharded.exchange_peer(rank, ep, b)
                blk = 1 << (nl - m)
                ops.append(dist.P2POp(dist.isend, torch.view_as_real(local[b * blk:(b + 1) * blk]), peer))
                ops.append(dist.P2POp(dist.irecv, torch.view_as_real(spare[b * blk:(b + 1) * blk]), peer))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        t_n = timeit(nccl_exchange)
        out[f"m{m}_nccl"] = {"ms": round(t_n, 3), "GBs": round((1 - 2.0 ** -m) * 8 * 2 ** nl / t_n / 1e6, 1)}
    if rank == 0:
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"scatter_bench_n{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
    st.release_peers()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
