#!/usr/bin/env python
"""Per-CTA phase timeline of the fused pass (debug hook ua_debug_set_fused_trace)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))
from unitair_b200 import _lib, circuit  # noqa: E402
sys.path.insert(0, ROOT)
from tools.prof_one import haar  # noqa: E402

n = 30
ngates = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda")
rng = np.random.default_rng(0)
a = torch.randn(2 ** n, dtype=torch.complex64, device=dev)
geo = circuit.TileGeometry(n, 13, 7, 6)
hb = list(range(n - 6, n))
gl = []
for i in range(ngates):
    b0, b1 = hb[i % 6], hb[(i * 2 + 1) % 6]
    if b0 == b1:
        b1 = hb[(i * 2 + 2) % 6]
    gl.append(([n - 1 - b0, n - 1 - b1], torch.as_tensor(haar(rng, 4).astype(np.complex64)).to(dev)))
cc = circuit.CompiledCircuit(gl, n, torch.complex64, (), geometry=geo, merge=False)
for _ in range(2):
    cc.run(a, in_place=True)
torch.cuda.synchronize()
lib = _lib.lib()
lib.ua_debug_set_fused_trace.argtypes = [ctypes.c_void_p]
buf = torch.zeros(148 * 4 * 16 * 4, dtype=torch.int64, device=dev)
lib.ua_debug_set_fused_trace(buf.data_ptr())
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); cc.run(a, in_place=True); e1.record(); torch.cuda.synchronize()
lib.ua_debug_set_fused_trace(None)
print("pass ms", e0.elapsed_time(e1), "gates", ngates)
t = buf.cpu().numpy().reshape(-1, 16, 4).astype(np.float64)
t0 = t[t[:, 0, 0] > 0][:, 0, 0].min()
for cta in (0, 148, 296, 5, 153):
    rows = t[cta]
    if rows[0, 0] == 0:
        continue
    print(f"CTA {cta}: (us since kernel start) wait_begin, landed, gates_done | load_wait, gate_phase, store_drain(prev)")
    for it in range(8):
        wb, ld, gd, sd = (rows[it] - t0) / 1e3
        print(f"  tile {it}: {wb:8.2f} {ld:8.2f} {gd:8.2f} | {ld - wb:6.2f} {gd - ld:6.2f} {(sd - (rows[it-1,2]-t0)/1e3) if it else 0:6.2f}")
valid = t[t[:, 1, 0] > 0]
lw = (valid[:, 1:12, 1] - valid[:, 1:12, 0]) / 1e3
gp = (valid[:, 1:12, 2] - valid[:, 1:12, 1]) / 1e3
sd = (valid[:, 2:12, 3] - valid[:, 1:11, 2]) / 1e3
print("mean over CTAs/tiles: load wait %.2f us, gate phase %.2f us, store drain %.2f us" % (lw.mean(), gp.mean(), sd.mean()))
