#!/usr/bin/env python
"""Deep circuit on a small state: eager launches vs CUDA-graph replay."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200")); sys.path.insert(0, ROOT)
from bench import random_circuit
from unitair_b200 import circuit
dev = torch.device("cuda")
for n, layers in ((16, 100), (20, 100), (24, 100)):
    gl = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in random_circuit(n, layers, 7)]
    cc = circuit.CompiledCircuit(gl, n, torch.complex64)
    st = torch.zeros(2 ** n, dtype=torch.complex64, device=dev); st[0] = 1
    for _ in range(2): cc.run(st, in_place=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): cc.run(st, in_place=True)
    torch.cuda.synchronize(); t_eager = (time.perf_counter() - t0) / 5
    g = cc.capture_graph(st)
    g.replay(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): g.replay()
    torch.cuda.synchronize(); t_graph = (time.perf_counter() - t0) / 5
    upd = len(gl) * 2.0 ** n
    print(f"n={n} gates={len(gl)} passes={cc.num_passes}: eager {t_eager*1e3:.2f} ms ({upd/t_eager:.3e} upd/s)  graph {t_graph*1e3:.2f} ms ({upd/t_graph:.3e} upd/s)")
