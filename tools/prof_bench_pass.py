#!/usr/bin/env python
"""Run the bench circuit (C2 recipe) once or twice on one GPU -- a short command for ncu:
    ncu --set full --clock-control none --import-source on -k regex:cluster -s 30 -c 3 -o gpurun_out/prof \
        python tools/prof_bench_pass.py --qubits 28
"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from bench import random_circuit
from unitair_b200 import circuit

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=28)
ap.add_argument("--layers", type=int, default=10)
ap.add_argument("--steps", type=int, default=2)
args = ap.parse_args()
n = args.qubits
dev = torch.device("cuda", 0)
gates = random_circuit(n, args.layers, 202)
g = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in gates]
cc = circuit.CompiledCircuit(g, n, torch.complex64)
state = torch.zeros(2 ** n, dtype=torch.complex64, device=dev)
state[0] = 1
for _ in range(args.steps):
    cc.run(state, in_place=True)
torch.cuda.synchronize()
print("passes per step", cc.num_passes, "gates", cc.num_gates)
