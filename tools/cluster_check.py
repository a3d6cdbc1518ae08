#!/usr/bin/env python
"""A/B of the two fused-pass gate phases on one GPU: parity (register-blocked path vs the
shared-memory-matrix path vs the numpy oracle) and the bench circuit's step time.

    python tools/cluster_check.py [--qubits 30] [--layers 10] [--steps 3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

from bench import random_circuit  # noqa: E402
from unitair_b200 import circuit, _lib  # noqa: E402


def compiled(gates_np, n, dev, cluster, host):
    os.environ["UA_CLUSTER"] = "1" if cluster else "0"
    if host:
        g = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in gates_np]
    else:
        g = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in gates_np]
    return circuit.CompiledCircuit(g, n, torch.complex64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--layers", type=int, default=10)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--skip-old", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = {"arith": os.environ.get("UA_CLUSTER_ARITH", "1")}

    # ---- parity ------------------------------------------------------------------------
    from oracle import unitair_oracle as orc
    rng = np.random.default_rng(7)
    for n, layers in ((6, 3), (11, 3), (13, 4), (16, 3), (21, 4)):
        gates = random_circuit(n, layers, 100 + n)
        st = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        st = (st / np.linalg.norm(st)).astype(np.complex64)
        psi = torch.as_tensor(st).to(dev)
        res = {}
        for name, cl, host in (("old", False, False), ("cluster_dev", True, False), ("cluster_host", True, True)):
            l0 = _lib.launch_count()
            res[name] = compiled(gates, n, dev, cl, host).run(psi).cpu().numpy()
        if n <= 16:
            ref = st
            for qs, u in gates:
                ref = orc.apply_operator(u.astype(np.complex64), qs, ref)
        else:
            ref = res["old"]
        for name in res:
            err = float(np.linalg.norm(res[name] - ref) / np.linalg.norm(ref))
            out[f"parity_n{n}_{name}"] = err
            assert err < 1e-5, (n, name, err)
    # adjoint flag + batch rows through the raw entry is covered by tests; here: batch of states
    n = 12
    gates = random_circuit(n, 2, 5)
    stb = (rng.standard_normal((5, 2 ** n)) + 1j * rng.standard_normal((5, 2 ** n))).astype(np.complex64)
    psi = torch.as_tensor(stb).to(dev)
    os.environ["UA_CLUSTER"] = "1"
    g = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in gates]
    a = circuit.CompiledCircuit(g, n, torch.complex64, (5,)).run(psi).cpu().numpy()
    os.environ["UA_CLUSTER"] = "0"
    b = circuit.CompiledCircuit(g, n, torch.complex64, (5,)).run(psi).cpu().numpy()
    out["parity_batch5_n12"] = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert out["parity_batch5_n12"] < 1e-5

    # ---- timing ------------------------------------------------------------------------
    n = args.qubits
    gates = random_circuit(n, args.layers, 202)
    state = torch.zeros(2 ** n, dtype=torch.complex64, device=dev)
    state[0] = 1
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    for name, cl in (("cluster", True),) + (() if args.skip_old else (("old", False),)):
        cc = compiled(gates, n, dev, cl, False)
        for _ in range(2):
            cc.run(state, in_place=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            cc.run(state, in_place=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[f"{name}_ms_per_step"] = ms
        out[f"{name}_passes"] = cc.num_passes
        out[f"{name}_ms_per_pass"] = ms / cc.num_passes
        out[f"{name}_gates"] = cc.num_gates
        out[f"{name}_norm"] = float(torch.linalg.vector_norm(state).item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
