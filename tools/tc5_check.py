#!/usr/bin/env python
"""Dense 5-qubit complex64 gate: tensor-core path (tcgen05 3xTF32, default) against the numpy
oracle at small sizes and its pass time at 30 qubits.  UA_TC5=0 selects the CUDA-core kernel."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import unitair_b200 as ua
from oracle import unitair_oracle as orc

rng = np.random.default_rng(5)


def haar(dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return (q * (d / np.abs(d))).astype(np.complex64)


out = {"tc5": os.environ.get("UA_TC5", "1")}
worst = 0.0
for n, qs, batch in [(11, [0, 1, 2, 3, 4], ()), (12, [4, 1, 6, 0, 3], ()), (13, [8, 2, 5, 0, 7], ()),
                     (14, [9, 0, 4, 7, 2], (3,)), (16, [11, 3, 0, 8, 5], ()), (12, [7, 6, 5, 4, 3], (2,)),
                     (15, [14, 2, 9, 5, 0], ()), (13, [1, 3, 5, 7, 8], ())]:
    u = haar(32)
    s = rng.standard_normal(batch + (2 ** n,)) + 1j * rng.standard_normal(batch + (2 ** n,))
    s = (s / np.linalg.norm(s, axis=-1, keepdims=True)).astype(np.complex64)
    print("case", n, qs, batch, flush=True)
    got = ua.simulation.apply_operator(torch.from_numpy(u).cuda(), qs, torch.from_numpy(s).cuda()).cpu().numpy()
    ref = orc.apply_operator(u, qs, s)
    err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    print("   rel err", err, flush=True)
    worst = max(worst, err)
out["worst_rel_err"] = worst
n = 30
a = torch.zeros(2 ** n, dtype=torch.complex64, device="cuda"); a[0] = 1
b = torch.empty_like(a)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
from unitair_b200 import _engine
for name, qs in {"contiguous_high": [0, 1, 2, 3, 4], "contiguous_mid": [10, 11, 12, 13, 14], "scattered": [2, 9, 14, 20, 25],
                 "scattered2": [0, 6, 11, 17, 23], "with_low_bits": [27, 28, 29, 3, 8]}.items():
    u = torch.from_numpy(haar(32)).cuda()
    for _ in range(2): _engine.launch_gate(b, a, u, n, 5, qs, 1, 1 << n, 0, False)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): _engine.launch_gate(b, a, u, n, 5, qs, 1, 1 << n, 0, False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[name] = {"ms": round(ms, 3), "GBs": round(16.0 * 2 ** n / ms / 1e6, 1), "TFLOPs": round(256.0 * 2 ** n / ms / 1e9, 1)}
print(json.dumps(out))
