#!/usr/bin/env python
"""The reference's own torch path on the SAME B200 ("the on-box baseline to
beat"): unmodified unitair from baseline/_ref with CUDA tensors (generic ATen copy kernels +
cuBLAS bmm), against this engine on identical inputs.  One layer of the bench circuit."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
from bench import random_circuit, load_reference  # noqa: E402
import unitair_b200 as ua  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    kind, ref = load_reference()
    assert kind == "reference", "baseline/_ref did not travel"
    dev = torch.device("cuda")
    out = {}
    for n, reps in ((24, 5), (28, 3), (30, 2)):
        gates = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in random_circuit(n, 1, 202)]
        psi0 = torch.zeros(2 ** n, dtype=torch.complex64, device=dev)
        psi0[0] = 1

        def run_ref():
            psi = psi0
            for qs, u in gates:
                psi = ref.apply_operator(operator=u, qubits=qs, state=psi)
            return psi

        def run_ours():
            psi = psi0
            for qs, u in gates:
                psi = ua.simulation.apply_operator(operator=u, qubits=qs, state=psi)
            return psi

        def run_fused():
            return ua.circuit.apply_gates(gates, psi0)
        a, b, c = run_ref(), run_ours(), run_fused()
        err = float((a - b).norm() / a.norm())
        errf = float((a - c).norm() / a.norm())
        del a, b, c
        torch.cuda.empty_cache()
        t_ref, t_ours, t_fused = timed(run_ref, reps), timed(run_ours, reps), timed(run_fused, reps)
        upd = len(gates) * 2.0 ** n
        out[f"n{n}"] = {"gates": len(gates),
                        "reference_torch_cuda_ms": round(t_ref, 3), "reference_updates_per_s": upd / t_ref * 1e3,
                        "ours_per_gate_ms": round(t_ours, 3), "ours_per_gate_updates_per_s": upd / t_ours * 1e3,
                        "ours_fused_ms": round(t_fused, 3), "ours_fused_updates_per_s": upd / t_fused * 1e3,
                        "speedup_per_gate": round(t_ref / t_ours, 2), "speedup_fused": round(t_ref / t_fused, 2),
                        "rel_diff_per_gate_vs_reference": err, "rel_diff_fused_vs_reference": errf,
                        "reference_peak_mem_GiB": None}
        torch.cuda.reset_peak_memory_stats()
        run_ref()
        torch.cuda.synchronize()
        out[f"n{n}"]["reference_peak_mem_GiB"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "reference_torch_cuda.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
