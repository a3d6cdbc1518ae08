import os, sys
import numpy as np
sys.path.insert(0,'/root/repo/qcware-unitair_b200'); sys.path.insert(0,'/root/repo')
import torch
import unitair_b200 as ua
from oracle import unitair_oracle as orc
rng=np.random.default_rng(1)
n=11; qs=[0,1,2,3,4]
s=(rng.standard_normal(2**n)+1j*rng.standard_normal(2**n)).astype(np.complex64)
mode=os.environ.get("UA_TC5_DEBUG","0")
for name,u in (("identity", np.eye(32,dtype=np.complex64)), ("diag", np.diag(np.arange(1,33)).astype(np.complex64)), ("shift", np.roll(np.eye(32),1,axis=0).astype(np.complex64)), ("i*identity", (1j*np.eye(32)).astype(np.complex64))):
    got=ua.simulation.apply_operator(torch.from_numpy(u).cuda(), qs, torch.from_numpy(s).cuda()).cpu().numpy()
    ref=orc.apply_operator(u,qs,s)
    print(mode, name, "err", np.linalg.norm(got-ref)/np.linalg.norm(ref), "nonzero", np.count_nonzero(got), "got[:3]", got[:3], "ref[:3]", ref[:3], "in[:3]", s[:3], flush=True)
    if name=="identity":
        # where do values land?
        idx=np.argsort(-np.abs(got))[:3]; print("   top got idx", idx, got[idx])
