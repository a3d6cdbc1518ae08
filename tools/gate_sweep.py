#!/usr/bin/env python
"""Per-kernel HBM bandwidth sweep on one GPU: every 1-qubit target bit, a set of 2-qubit
pairs, k=3..5 blocks, phase, reductions, fused passes.  Writes gpurun_out/sweep_<tag>.json.
Timing: CUDA events, 3 warm-ups, median of `reps`; states >= 1 GiB so inputs exceed L2."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
import unitair_b200 as ua  # noqa: E402
from unitair_b200 import _engine, circuit  # noqa: E402


def timeit(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
    return float(np.median(ts))


def haar(rng, dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--dtype", default="c64")
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--what", default="gate1,gate2,gatek,phase,reduce,fused,copy")
    args = ap.parse_args()
    n = args.qubits
    cd = torch.complex64 if args.dtype == "c64" else torch.complex128
    npc = np.complex64 if args.dtype == "c64" else np.complex128
    esz = 8 if args.dtype == "c64" else 16
    rng = np.random.default_rng(0)
    dev = torch.device("cuda")
    a = torch.randn(2 ** n, dtype=cd, device=dev)
    a /= a.abs().pow(2).sum().sqrt()
    b = torch.empty_like(a)
    bytes_rw = 2.0 * esz * 2 ** n
    res = {"qubits": n, "dtype": args.dtype, "bytes_rw": bytes_rw}
    what = args.what.split(",")

    def gbs(t, nbytes=bytes_rw):
        return nbytes / t / 1e9

    if "copy" in what:
        res["torch_copy_GBs"] = gbs(timeit(lambda: b.copy_(a)))
    if "gate1" in what:
        u = torch.as_tensor(haar(rng, 2).astype(npc)).to(dev)
        out = {}
        for q in range(n):
            t = timeit(lambda: _engine.launch_gate(b, a, u, n, 1, [q], 1, 1 << n, 0, False))
            out[f"bit{n - 1 - q}"] = round(gbs(t), 1)
        res["gate1_outofplace_GBs_by_bit"] = out
        out = {}
        for q in (0, n // 2, n - 2, n - 1):
            t = timeit(lambda: _engine.launch_gate(a, a, u, n, 1, [q], 1, 1 << n, 0, False))
            out[f"bit{n - 1 - q}"] = round(gbs(t), 1)
        res["gate1_inplace_GBs_by_bit"] = out
    if "gate2" in what:
        u = torch.as_tensor(haar(rng, 4).astype(npc)).to(dev)
        out = {}
        for qs in [(0, 1), (0, n - 1), (n - 1, 0), (n - 2, n - 1), (n // 2, n // 2 + 1), (3, n - 4), (n - 1, n - 3)]:
            t = timeit(lambda: _engine.launch_gate(b, a, u, n, 2, list(qs), 1, 1 << n, 0, False))
            out[str(tuple(n - 1 - q for q in qs))] = round(gbs(t), 1)
        res["gate2_outofplace_GBs_by_bits"] = out
    if "gatek" in what:
        out = {}
        for k in (3, 4, 5):
            u = torch.as_tensor(haar(rng, 2 ** k).astype(npc)).to(dev)
            for qs in [tuple(range(k)), tuple(range(n - k, n)), tuple(int(x) for x in rng.permutation(n)[:k])]:
                t = timeit(lambda: _engine.launch_gate(b, a, u, n, k, list(qs), 1, 1 << n, 0, False), reps=5)
                out[f"k{k}:{tuple(n - 1 - q for q in qs)}"] = {"GBs": round(gbs(t), 1), "ms": round(t * 1e3, 3),
                                                                 "TFLOPs": round(8.0 * 2 ** k * 2 ** n / t / 1e12, 2)}
        res["gatek_outofplace"] = out
    if "phase" in what:
        rdt = torch.float32 if args.dtype == "c64" else torch.float64
        ang = torch.rand(2 ** n, dtype=rdt, device=dev)
        t = timeit(lambda: ua.simulation.apply_phase(ang, a))
        res["phase_full_GBs"] = round(gbs(t, (2 * esz + esz // 2) * 2.0 ** n), 1)
        sc = torch.rand((), dtype=rdt, device=dev)
        t = timeit(lambda: ua.simulation.apply_phase(sc, a))
        res["phase_scalar_GBs"] = round(gbs(t), 1)
        del ang
    if "reduce" in what:
        t = timeit(lambda: ua.norm_squared(a))
        res["norm_squared_GBs"] = round(gbs(t, esz * 2.0 ** n), 1)
        t = timeit(lambda: ua.abs_squared(a))
        res["abs_squared_GBs"] = round(gbs(t, 1.5 * esz * 2.0 ** n), 1)
        t = timeit(lambda: ua.inner_product(a, b))
        res["inner_product_GBs"] = round(gbs(t, 2.0 * esz * 2.0 ** n), 1)
    if "grad" in what:
        out = {}
        for k, qs in ((1, [3]), (1, [n - 1]), (2, [5, 17]), (2, [n - 1, 0]), (3, [2, 9, 20])):
            u = torch.as_tensor(haar(rng, 2 ** k).astype(npc)).to(dev)
            t = timeit(lambda: _engine.launch_gate_grad(b, a, n, k, qs, 1, 1 << n, 0, u.shape), reps=5)
            out[f"k{k}:{qs}"] = {"ms": round(t * 1e3, 3), "GBs": round(gbs(t, 2.0 * esz * 2 ** n), 1)}
        res["gate_grad"] = out
    if "permute" in what:
        out = {}
        for name, perm in (("swap(0,1)", [1, 0] + list(range(2, n))), ("swap(0,n-1)", [n - 1] + list(range(1, n - 1)) + [0]),
                           ("swap(n-2,n-1)", list(range(n - 2)) + [n - 1, n - 2]), ("roll", [n - 1] + list(range(n - 1))),
                           ("random", [int(x) for x in rng.permutation(n)])):
            t = timeit(lambda: circuit.permute_qubits_native(perm, a, n), reps=5)
            out[name] = {"ms": round(t * 1e3, 3), "GBs": round(gbs(t), 1)}
        res["permute_qubits"] = out
    if "fused" in what:
        out = {}
        for tile_bits, low in ((12, 7), (13, 7), (14, 7), (13, 6), (13, 8)):
            if args.dtype == "c128" and tile_bits > 13:
                continue
            geo = circuit.TileGeometry(n, tile_bits, low, tile_bits - low)
            for ngates in (1, 4, 8, 16, 32):
                # 1-qubit gates spread over the tile bits (low + top high bits)
                bits = list(range(low)) + list(range(n - (tile_bits - low), n))
                gl = []
                for i in range(ngates):
                    bpos = bits[(i * 5) % len(bits)]
                    gl.append(([n - 1 - bpos], torch.as_tensor(haar(rng, 2).astype(npc)).to(dev)))
                cc = circuit.CompiledCircuit(gl, n, cd, (), geometry=geo)
                t = timeit(lambda: cc.run(a, in_place=True), reps=5)
                out[f"T{tile_bits}L{low}:g{ngates}"] = {"passes": cc.num_passes, "ms": round(t * 1e3, 3),
                                                       "GBs_per_pass": round(gbs(t / cc.num_passes), 1)}
            # 2-qubit gates
            gl = []
            bits = list(range(low)) + list(range(n - (tile_bits - low), n))
            for i in range(8):
                b0, b1 = bits[(i * 3) % len(bits)], bits[(i * 3 + 7) % len(bits)]
                if b0 == b1:
                    b1 = bits[(i * 3 + 8) % len(bits)]
                gl.append(([n - 1 - b0, n - 1 - b1], torch.as_tensor(haar(rng, 4).astype(npc)).to(dev)))
            cc = circuit.CompiledCircuit(gl, n, cd, (), geometry=geo)
            t = timeit(lambda: cc.run(a, in_place=True), reps=5)
            out[f"T{tile_bits}L{low}:8x2q"] = {"passes": cc.num_passes, "ms": round(t * 1e3, 3),
                                                "GBs_per_pass": round(gbs(t / cc.num_passes), 1)}
        res["fused"] = out
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"sweep_{args.tag}_{args.dtype}_n{n}.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
