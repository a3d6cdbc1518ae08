#!/usr/bin/env python
"""BASELINE.json configs C1-C4 at full size on one GPU (C5 = bench.py --gpus 8).
Writes gpurun_out/configs_<tag>.json.  CUDA events, 2 warm-ups, median of 5."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
import unitair_b200 as ua  # noqa: E402
from bench import haar_unitary, random_circuit  # noqa: E402

dev = torch.device("cuda")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 1e3)
    return float(np.median(ts))


def c1():
    n, B = 10, 1024
    g = torch.Generator(device="cpu").manual_seed(101)
    st = ua.rand_state(n, (B,), generator=g).to(dev)
    theta = (torch.rand(n, generator=g) * 2 * np.pi).to(dev)
    h, cn = ua.gates.hadamard(device=dev), ua.gates.cnot(device=dev)

    def gates():
        gl = [([q], h) for q in range(n)]
        gl += [([q], ua.gates.exp_x(theta[q])) for q in range(n)]
        gl += [([q, q + 1], cn) for q in range(n - 1)]
        return gl

    def per_op():
        psi = st
        for qs, u in gates():
            psi = ua.simulation.apply_operator(u, qs, psi)
        return psi

    def fused():
        return ua.circuit.apply_gates(gates(), st)
    upd = 29 * B * 2.0 ** n
    t1, t2 = timeit(per_op), timeit(fused)
    return {"config": "C1 n=10 B=1024 c64 H+rx+CNOT chain (29 gates)", "per_op_ms": t1 * 1e3, "fused_ms": t2 * 1e3,
            "per_op_updates_per_s": upd / t1, "fused_updates_per_s": upd / t2}


def c2():
    n, layers = 24, 200
    gl = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in random_circuit(n, layers, 202)]
    st = ua.unit_vector(0, num_qubits=n, device=dev)
    cc = ua.circuit.CompiledCircuit(gl, n, torch.complex64)
    t = timeit(lambda: cc.run(st), reps=3, warm=1)

    def per_op(count=360):
        psi = st
        for qs, u in gl[:count]:
            psi = ua.simulation.apply_operator(u, qs, psi)
        return psi
    t2 = timeit(per_op, reps=3, warm=1)
    upd = len(gl) * 2.0 ** n
    return {"config": "C2 n=24 c64 200 layers (7200 gates)", "fused_ms": t * 1e3, "passes": cc.num_passes,
            "fused_updates_per_s": upd / t, "per_op_updates_per_s": 360 * 2.0 ** n / t2,
            "per_op_us_per_gate": t2 / 360 * 1e6,
            "norm_after": float(ua.norm_squared(cc.run(st)))}


def c3():
    n, B, layers = 16, 4096, 20
    g = torch.Generator(device="cpu").manual_seed(303)
    st = ua.rand_state(n, (B,), generator=g).to(dev)
    theta = (torch.rand(layers, n, 2, generator=g) * 2 * np.pi).to(dev).requires_grad_(True)
    cn = ua.gates.cnot(device=dev)
    z0 = torch.where((torch.arange(2 ** n, device=dev) >> (n - 1)) & 1 == 0, 1.0, -1.0)

    def step():
        gl = []
        for l in range(layers):
            for q in range(n):
                gl.append(([q], ua.gates.exp_y(theta[l, q, 0])))
                gl.append(([q], ua.gates.exp_z(theta[l, q, 1])))
            for q in range(n - 1):
                gl.append(([q, q + 1], cn))
        psi = ua.circuit.apply_gates(gl, st, assume_unitary=True)
        loss = ua.diag_expectation_value(z0, psi).sum()
        gth, = torch.autograd.grad(loss, theta)
        return loss, gth, len(gl)
    torch.cuda.reset_peak_memory_stats()
    t = timeit(lambda: step(), reps=3, warm=1)
    loss, gth, ng = step()
    return {"config": "C3 n=16 B=4096 c64 ansatz 20 layers (940 gates) fwd + grad <Z0> (adjoint method)",
            "fwd_bwd_ms": t * 1e3, "updates_per_s_fwd_equiv": ng * B * 2.0 ** n / t,
            "peak_mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30,
            "loss": float(loss), "grad_norm": float(gth.norm())}


def c4():
    n, layers = 30, 10
    rng = np.random.default_rng(404)
    st = torch.zeros(2 ** n, dtype=torch.complex128, device=dev)
    st[0] = 1
    blocks = []
    for l in range(layers):
        perm = rng.permutation(n).tolist()
        blocks.append([(perm[j:j + 5], torch.as_tensor(haar_unitary(rng, 32)).to(dev)) for j in range(0, n, 5)])
    ang = torch.rand(2 ** n, dtype=torch.float64, device=dev) * 2 * np.pi

    def run():
        psi = st
        for l in range(layers):
            for qs, u in blocks[l]:
                psi = ua.simulation.apply_operator(u, qs, psi)
            psi = ua.simulation.apply_phase(ang, psi)
        return psi
    t = timeit(run, reps=3, warm=1)
    nrm = float(ua.norm_squared(run()))
    ops = layers * 7
    return {"config": "C4 n=30 c128 10 layers x (6 Haar U(32) blocks + f64 phase layer)", "ms": t * 1e3,
            "updates_per_s": ops * 2.0 ** n / t, "ms_per_op": t / ops * 1e3, "norm_after": nrm}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c1", "c2", "c3", "c4"]
    res = {}
    for name in which:
        try:
            res[name] = globals()[name]()
        except Exception as e:
            res[name] = {"error": repr(e)[:300]}
        print(name, json.dumps(res[name]), flush=True)
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"configs_{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
