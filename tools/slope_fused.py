#!/usr/bin/env python
"""Pass time versus number of gates in the pass (intercept = tile traffic, slope = gate phase).

    python tools/slope_fused.py [qubits] [name=K1:V1,K2:V2 ...]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
from bench import haar_unitary  # noqa: E402
from unitair_b200 import circuit  # noqa: E402
from tools.ab_fused import KNOBS  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    specs = sys.argv[2:] or ["base="]
    dev = torch.device("cuda")
    rng = np.random.default_rng(3)
    state = torch.randn(2 ** n, dtype=torch.complex64, device=dev)
    state /= state.norm()
    rows = []
    for spec in specs:
        name, _, kv = spec.partition("=")
        for k in KNOBS:
            os.environ.pop(k, None)
        for item in filter(None, kv.split(",")):
            k, v = item.split(":")
            os.environ[k] = v
        geo = circuit.default_geometry(n, torch.complex64)
        hb = list(range(n - geo.max_high, n))          # the tile's high bits = top bits of the index
        lb = list(range(2, geo.low_bits))              # low bits above the 16-byte vector's neighbourhood
        for kind, pool in (("high", hb), ("mixed", hb + lb)):
            times = {}
            for cnt in (1, 2, 4, 6, 8, 12):
                gl = []
                for i in range(cnt):
                    b0 = pool[i % len(pool)]
                    b1 = pool[(i * 2 + 1) % len(pool)]
                    if b0 == b1:
                        b1 = pool[(i * 2 + 2) % len(pool)]
                    u = torch.as_tensor(haar_unitary(rng, 4).astype(np.complex64)).to(dev)
                    gl.append(([n - 1 - b0, n - 1 - b1], u))
                cc = circuit.CompiledCircuit(gl, n, torch.complex64, (), merge=False)
                assert cc.num_passes == 1, cc.num_passes
                for _ in range(2):
                    cc.run(state, in_place=True)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    cc.run(state, in_place=True)
                e1.record()
                torch.cuda.synchronize()
                times[cnt] = round(e0.elapsed_time(e1) / 5, 3)
            xs = np.array(list(times.keys()), dtype=float)
            ys = np.array(list(times.values()))
            slope, icpt = np.polyfit(xs, ys, 1)
            r = dict(name=name, env=kv, targets=kind, ms=times, slope_ms_per_gate=round(float(slope), 3),
                     intercept_ms=round(float(icpt), 3))
            rows.append(r)
            print(json.dumps(r), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"slope_n{n}.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
