#!/usr/bin/env python
"""Run ONE named kernel scenario a few times (for `ncu --set full -k regex:... -s N -c M`)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
import unitair_b200 as ua  # noqa: E402
from unitair_b200 import _engine, circuit  # noqa: E402


def haar(rng, dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenario")
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--dtype", default="c64")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tile", type=int, default=13)
    ap.add_argument("--low", type=int, default=7)
    args = ap.parse_args()
    n = args.qubits
    cd = torch.complex64 if args.dtype == "c64" else torch.complex128
    npc = np.complex64 if args.dtype == "c64" else np.complex128
    rng = np.random.default_rng(0)
    dev = torch.device("cuda")
    a = torch.randn(2 ** n, dtype=cd, device=dev)
    b = torch.empty_like(a)

    def gate(k):
        return torch.as_tensor(haar(rng, 2 ** k).astype(npc)).to(dev)

    sc = args.scenario
    if sc.startswith("gate"):              # gate<k>:<q0>,<q1>...   e.g. gate1:0  gate5:0,1,2,3,4
        k = int(sc[4])
        qs = [int(x) for x in sc.split(":")[1].split(",")]
        u = gate(k)
        fn = lambda: _engine.launch_gate(b, a, u, n, k, qs, 1, 1 << n, 0, False)   # noqa
    elif sc.startswith("fused"):           # fused:<n1q>,<n2q>
        n1, n2 = [int(x) for x in sc.split(":")[1].split(",")]
        geo = circuit.TileGeometry(n, args.tile, args.low, args.tile - args.low)
        bits = list(range(args.low)) + list(range(n - (args.tile - args.low), n))
        gl = []
        for i in range(n1):
            gl.append(([n - 1 - bits[(i * 5) % len(bits)]], gate(1)))
        for i in range(n2):
            b0 = bits[(i * 3) % len(bits)]
            b1 = bits[(i * 3 + 7) % len(bits)]
            if b0 == b1:
                b1 = bits[(i * 3 + 8) % len(bits)]
            gl.append(([n - 1 - b0, n - 1 - b1], gate(2)))
        cc = circuit.CompiledCircuit(gl, n, cd, (), geometry=geo, merge=False)
        print("passes", cc.num_passes, "gates", cc.num_gates)
        fn = lambda: cc.run(a, in_place=True)   # noqa
    elif sc.startswith("hi2q"):            # hi2q:<count>  2-qubit gates on the tile's HIGH bits only
        cnt = int(sc.split(":")[1])
        geo = circuit.TileGeometry(n, args.tile, args.low, args.tile - args.low)
        hb = list(range(n - (args.tile - args.low), n))          # bit positions
        gl = []
        for i in range(cnt):
            b0, b1 = hb[i % len(hb)], hb[(i * 2 + 1) % len(hb)]
            if b0 == b1:
                b1 = hb[(i * 2 + 2) % len(hb)]
            gl.append(([n - 1 - b0, n - 1 - b1], gate(2)))
        cc = circuit.CompiledCircuit(gl, n, cd, (), geometry=geo, merge=False)
        print("passes", cc.num_passes, "gates", cc.num_gates)
        fn = lambda: cc.run(a, in_place=True)   # noqa
    elif sc.startswith("bench"):           # bench:<layers>  -- the bench.py circuit (C2 recipe)
        sys.path.insert(0, ROOT)
        from bench import random_circuit
        layers = int(sc.split(":")[1])
        gl = [(qs, torch.as_tensor(u.astype(npc)).to(dev)) for qs, u in random_circuit(n, layers, 202)]
        cc = circuit.CompiledCircuit(gl, n, cd)
        print("passes", cc.num_passes, "gates", cc.num_gates, "source gates", cc.num_source_gates)
        a.zero_()
        a[0] = 1
        fn = lambda: cc.run(a, in_place=True)   # noqa
    elif sc == "phase":
        ang = torch.rand(2 ** n, dtype=torch.float32 if args.dtype == "c64" else torch.float64, device=dev)
        fn = lambda: ua.simulation.apply_phase(ang, a)   # noqa
    else:
        raise SystemExit("unknown scenario")
    for _ in range(args.reps):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(sc, "ms", e0.elapsed_time(e1))


if __name__ == "__main__":
    main()
