#!/usr/bin/env python
"""Time the bench circuit (C2 recipe) under different fused-pass settings (env knobs)."""
import itertools
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dt = sys.argv[2] if len(sys.argv) > 2 else "c64"
    layers = 6
    cd = torch.complex64 if dt == "c64" else torch.complex128
    npc = np.complex64 if dt == "c64" else np.complex128
    dev = torch.device("cuda")
    gates = [(qs, torch.as_tensor(u.astype(npc)).to(dev)) for qs, u in random_circuit(n, layers, 202)]
    state = torch.zeros(2 ** n, dtype=cd, device=dev)
    state[0] = 1
    rows = []
    tiles = [(13, 7), (12, 7), (14, 7), (13, 8)] if dt == "c64" else [(12, 6), (11, 6), (13, 6)]
    for (tile, low), threads, swz, stages in itertools.product(tiles, (512, 256), (0, 1), (0, 1, 2)):
        os.environ["UA_TILE_BITS"] = str(tile)
        os.environ["UA_TILE_LOW_BITS"] = str(low)
        os.environ["UA_FUSED_THREADS"] = str(threads)
        os.environ["UA_FUSED_SWZ"] = str(swz)
        os.environ["UA_FUSED_STAGES"] = str(stages)
        try:
            cc = circuit.CompiledCircuit(gates, n, cd)
            for _ in range(2):
                cc.run(state, in_place=True)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                cc.run(state, in_place=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            r = dict(tile=tile, low=low, threads=threads, swz=swz, stages=stages, passes=cc.num_passes,
                     merged_gates=cc.num_gates, ms=round(ms, 2), ms_per_pass=round(ms / cc.num_passes, 3),
                     updates_per_s=float(len(gates)) * 2 ** n / ms * 1e3)
        except Exception as e:
            r = dict(tile=tile, low=low, threads=threads, swz=swz, stages=stages, error=str(e)[:100])
        rows.append(r)
        print(json.dumps(r), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"tune_fused_{dt}_n{n}.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
