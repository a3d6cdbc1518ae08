#!/usr/bin/env python
"""A/B the bench circuit (C2 recipe, 30 qubits complex64) under named env-knob sets.

    python tools/ab_fused.py [qubits] [layers] [name=K1:V1,K2:V2 ...]

Every configuration is checked against the first one (relative L2 difference of the final
state) so that a faster variant that computes something else is caught immediately.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
sys.path.insert(0, ROOT)
from bench import random_circuit  # noqa: E402
from unitair_b200 import circuit  # noqa: E402

DEFAULT = [
    "base=",
    "merge3=UA_MERGE_MAX_K:3",
    "f2lean=UA_FUSED_F2:1",
    "f2lean_merge3=UA_FUSED_F2:1,UA_MERGE_MAX_K:3",
    "f2_256x2=UA_FUSED_F2:1,UA_FUSED_THREADS:256",
    "f2_128x4=UA_FUSED_F2:1,UA_FUSED_THREADS:128",
]
KNOBS = ("UA_MERGE_MAX_K", "UA_FUSED_F2", "UA_FUSED_THREADS", "UA_FUSED_STAGES", "UA_FUSED_SWZ",
         "UA_TILE_BITS", "UA_TILE_LOW_BITS", "UA_FUSED_TEAM", "UA_FUSED_CTAS", "UA_FUSED_L2PF",
         "UA_FUSED_MMA", "UA_FUSED_VARIANT")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    specs = sys.argv[3:] or DEFAULT
    dev = torch.device("cuda")
    gates = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in random_circuit(n, layers, 202)]
    rng = torch.Generator(device="cpu").manual_seed(7)
    init = torch.randn(2 ** min(n, 24), 2, generator=rng)
    init = torch.view_as_complex(init).to(dev)
    init = init.repeat(2 ** (n - min(n, 24)))
    init /= init.norm()
    ref = None
    rows = []
    for spec in specs:
        name, _, kv = spec.partition("=")
        for k in KNOBS:
            os.environ.pop(k, None)
        for item in filter(None, kv.split(",")):
            k, v = item.split(":")
            os.environ[k] = v
        try:
            cc = circuit.CompiledCircuit(gates, n, torch.complex64)
            state = init.clone()
            cc.run(state, in_place=True)
            if ref is None:
                ref = state.clone()
                err = 0.0
            else:
                err = float((state - ref).norm() / ref.norm())
            for _ in range(2):
                cc.run(state, in_place=True)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                cc.run(state, in_place=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 4
            r = dict(name=name, env=kv, passes=cc.num_passes, blocks=cc.num_gates, ms=round(ms, 2),
                     ms_per_pass=round(ms / cc.num_passes, 3), rel_diff_vs_first=err,
                     updates_per_s=float(len(gates)) * 2 ** n / ms * 1e3)
            del state
        except Exception as e:  # noqa: BLE001
            r = dict(name=name, env=kv, error=str(e)[:200])
        rows.append(r)
        print(json.dumps(r), flush=True)
    tag = os.environ.get("UA_AB_TAG", "ab")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"{tag}_n{n}.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
