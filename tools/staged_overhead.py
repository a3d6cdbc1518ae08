"""Single-GPU cost of running the last passes of an epoch slice by slice (pipelined exchange)
against the one-kernel scatter tail, all destinations in local memory: launch / ring fill
overhead of the slices and of the copy-engine transfers, without NVLink in the picture."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
from bench import random_circuit  # noqa: E402
from unitair_b200 import circuit  # noqa: E402


def main():
    n = int(os.environ.get("N", 30))
    victims = [int(v) for v in os.environ.get("VICTIMS", "15,20,25").split(",")]
    m = len(victims)
    gates = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in random_circuit(n, 2, 5)]
    st = torch.zeros(1 << n, dtype=torch.complex64, device="cuda")
    st[0] = 1
    out = torch.empty_like(st)
    stage = torch.empty_like(st)
    block_bytes = (8 << n) >> m
    dst = [out.data_ptr() + b * block_bytes for b in range(1 << m)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    order = list(range(1, 1 << m)) + [0]
    res = {"n": n, "victims": victims}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    cc0 = circuit.CompiledCircuit(gates, n, torch.complex64, tail_forbidden=victims)
    t0 = circuit.ScatterTail(cc0, n, torch.complex64, victims)
    res["direct"] = {"passes": t0.num_passes, "ms": round(timed(lambda: t0.run(st, dst)), 3)}
    for c in (1, 2, 3):
        for depth in (1, 2, 3):
            stay = sorted([b for b in range(n - 1, -1, -1) if b not in victims][:c])
            cc1 = circuit.CompiledCircuit(gates, n, torch.complex64, tail_forbidden=victims, tail_chunk=(stay, depth))
            t1 = circuit.ScatterTail(cc1, n, torch.complex64, victims, chunk_bits=stay, depth=depth)
            if not t1.chunk_bits:
                continue
            ms = timed(lambda: t1.run_staged(st, stage, dst, order, streams))
            res[f"staged_c{c}_d{depth}"] = {"passes": t1.num_passes, "pipe_depth": t1.pipe_depth, "ms": round(ms, 3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
