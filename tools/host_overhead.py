#!/usr/bin/env python
"""Host-side cost per public-API call on a tiny state (GPU time ~ 0): wall-clock us per call."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
import unitair_b200 as ua  # noqa: E402

dev = torch.device("cuda")
st = ua.rand_state(10, (64,)).to(dev)
h = ua.gates.hadamard(device=dev)
cn = ua.gates.cnot(device=dev)
ang = torch.rand(1024, device=dev)


def bench(name, fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:40s} {1e6 * (t1 - t0) / n:7.1f} us/call (host issue)   {1e6 * (t2 - t0) / n:7.1f} us/call (incl. drain)")


bench("apply_operator 1q", lambda: ua.simulation.apply_operator(h, (3,), st))
bench("apply_operator 2q", lambda: ua.simulation.apply_operator(cn, (3, 7), st))
bench("apply_phase", lambda: ua.simulation.apply_phase(ang, st))
bench("norm_squared", lambda: ua.norm_squared(st))
bench("abs_squared", lambda: ua.abs_squared(st))
bench("apply_all_qubits", lambda: ua.simulation.apply_all_qubits(h, st), n=300)
bench("torch.empty_like (reference point)", lambda: torch.empty_like(st))
bench("torch mul (reference point)", lambda: st * 2.0)
