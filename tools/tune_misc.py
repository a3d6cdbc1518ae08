#!/usr/bin/env python
"""Timing of the two kernels round 1 left untuned, at 30 qubits complex64:
ua_permute_bits (random / swap / roll permutations) against the copy peak (16 B per amplitude),
ua_gate_grad (1-/2-qubit gates on low / middle / high bits) against 16 B per amplitude."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import unitair_b200 as ua
from unitair_b200 import _engine

n = int(os.environ.get("N", 30))
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
a = torch.randn(2 ** n, dtype=torch.complex64, device=dev)
g = torch.randn(2 ** n, dtype=torch.complex64, device=dev)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
out = {"n": n}


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


bytes_pass = 16.0 * 2 ** n
perms = {"random_a": rng.permutation(n).tolist(), "random_b": rng.permutation(n).tolist(),
         "swap_high(0,1)": [1, 0] + list(range(2, n)), "swap_low": list(range(n - 2)) + [n - 1, n - 2],
         "swap(0,n-1)": [n - 1] + list(range(1, n - 1)) + [0], "roll_1": list(range(n))[-1:] + list(range(n))[:-1],
         "reverse": list(range(n))[::-1]}
for name, perm in perms.items():
    ms = timeit(lambda: ua.simulation.permute_qubits(perm, a))
    out["permute_" + name] = {"ms": round(ms, 3), "GBs": round(bytes_pass / ms / 1e6, 1)}
for name, qs in {"k1_high": [0], "k1_mid": [15], "k1_low": [n - 1], "k2_high": [0, 1], "k2_mixed": [3, n - 1],
                 "k2_low": [n - 2, n - 1], "k2_mid": [12, 20]}.items():
    k = len(qs)
    ms = timeit(lambda: _engine.launch_gate_grad(g, a, n, k, qs, 1, 2 ** n, 0, (2 ** k, 2 ** k)))
    out["gate_grad_" + name] = {"ms": round(ms, 3), "GBs": round(bytes_pass / ms / 1e6, 1)}
ms = timeit(lambda: torch.empty_like(a).copy_(a))
out["torch_copy"] = {"ms": round(ms, 3), "GBs": round(bytes_pass / ms / 1e6, 1)}
print(json.dumps(out))
