#!/usr/bin/env python
"""Stress the register-blocked pass kernel over many sizes (run under `timeout`): prints one line
per case before and after it runs, so a hang is attributable."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from bench import random_circuit
from unitair_b200 import circuit
dev = torch.device("cuda", 0)
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "6,9,11,12,13,14,16,18,21,24,27,30".split(","))]
for n in sizes:
    gates = random_circuit(n, 3, 100 + n)
    g = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in gates]
    os.environ["UA_CLUSTER"] = "1"
    cc = circuit.CompiledCircuit(g, n, torch.complex64)
    os.environ["UA_CLUSTER"] = "0"
    ref = circuit.CompiledCircuit([(qs, u.to(dev)) for qs, u in g], n, torch.complex64)
    st = torch.randn(2 ** n, dtype=torch.complex64, device=dev)
    st /= torch.linalg.vector_norm(st)
    print(f"n={n} passes={cc.num_passes} start", flush=True)
    t0 = time.time()
    for rep in range(5):
        out = cc.run(st)
        torch.cuda.synchronize()
    want = ref.run(st)
    err = float(torch.linalg.vector_norm(out - want) / torch.linalg.vector_norm(want))
    print(f"n={n} done in {time.time()-t0:.2f}s err={err:.2e}", flush=True)
    assert err < 1e-5
print("stress ok")
