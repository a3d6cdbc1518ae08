import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from unitair_b200 import _engine
n = 30
a = torch.zeros(2 ** n, dtype=torch.complex64, device="cuda"); a[0] = 1
b = torch.empty_like(a)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
out = {"debug": os.environ.get("UA_TC5_DEBUG", "0")}
rng = np.random.default_rng(0)
for name, qs in {"contiguous_mid": [10, 11, 12, 13, 14], "scattered": [2, 9, 14, 20, 25]}.items():
    u = torch.from_numpy((rng.standard_normal((32, 32)) + 1j * rng.standard_normal((32, 32))).astype(np.complex64)).cuda()
    for _ in range(2): _engine.launch_gate(b, a, u, n, 5, qs, 1, 1 << n, 0, False)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): _engine.launch_gate(b, a, u, n, 5, qs, 1, 1 << n, 0, False)
    e1.record(); torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / 5, 3)
print(json.dumps(out))
