#!/usr/bin/env python
"""Time ONE fused pass of six 2-qubit gates at n qubits for target sets that avoid / hit the
low index bits (shared-memory bank bits 0..3): isolates the bank-conflict cost of the gate phase."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from bench import haar_unitary
from unitair_b200 import circuit

n = int(os.environ.get("N", 30))
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1)
state = torch.zeros(2 ** n, dtype=torch.complex64, device=dev); state[0] = 1
cases = {
    "high_only(7..12)": [(7, 8), (9, 10), (11, 12), (7, 9), (8, 11), (10, 12)],
    "mid(4..6+high)": [(4, 8), (5, 10), (6, 12), (4, 9), (5, 11), (6, 7)],
    "low(0..3+high)": [(0, 8), (1, 10), (2, 12), (3, 9), (0, 11), (1, 7)],
    "lowlow(0..3 pairs)": [(0, 1), (2, 3), (0, 2), (1, 3), (0, 3), (1, 2)],
    "one_gate_high": [(7, 8)],
    "two_gates_high": [(7, 8), (9, 10)],
    "two_gates_4bits": [(7, 8), (9, 10), (7, 9), (8, 10)],
}
out = {}
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for name, pairs in cases.items():
    gates = [([n - 1 - a, n - 1 - b], torch.as_tensor(haar_unitary(rng, 4).astype(np.complex64)).to(dev)) for a, b in pairs]
    cc = circuit.CompiledCircuit(gates, n, torch.complex64, merge=False)
    for _ in range(2): cc.run(state, in_place=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): cc.run(state, in_place=True)
    e1.record(); torch.cuda.synchronize()
    out[name] = {"passes": cc.num_passes, "ms_per_pass": e0.elapsed_time(e1) / 5 / cc.num_passes}
print(json.dumps(out))
