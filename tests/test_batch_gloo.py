"""Batch-dimension sharding (unitair_b200/batch.py) on CPU with the gloo backend, world 2 and 3:
the helpers are backend-agnostic, so a plain torch model stands in for the engine.  The GPU
version of the same test (config C3 split over the GPUs of the box) is tests/test_gpu_batch.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qcware-unitair_b200"))


def _loss(theta, w, states):
    """A toy 'circuit': shared parameters theta, per-entry parameters w, complex states."""
    phase = torch.exp(1j * (theta.sum() + w)).unsqueeze(-1)              # (B, 1)
    psi = states * phase * torch.cos(theta).repeat(states.shape[-1] // theta.numel())
    return (psi.abs() ** 2 * torch.arange(states.shape[-1])).sum()


def _worker(rank, world, port, batch, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unitair_b200 import batch as ub
    torch.manual_seed(0)
    theta = torch.randn(4, dtype=torch.float64, requires_grad=True)
    w = torch.randn(batch, dtype=torch.float64, requires_grad=True)
    states = torch.randn(batch, 8, dtype=torch.complex128)
    # reference: the whole batch on one rank
    full = _loss(theta, w, states)
    g_theta, g_w = torch.autograd.grad(full, (theta, w))
    # sharded: every rank its rows, one all_reduce for the shared parameter
    sl = ub.batch_slice(batch)
    assert sl == ub.batch_slice(batch, rank, world)
    th = theta.detach().clone().requires_grad_(True)
    wl = ub.shard_batch(w.detach()).clone().requires_grad_(True)
    local = _loss(th, wl, ub.shard_batch(states))
    local.backward()
    ub.all_reduce_gradients([th])
    assert torch.allclose(th.grad, g_theta, rtol=1e-12, atol=1e-12)
    assert torch.allclose(wl.grad, g_w[sl], rtol=1e-12, atol=1e-12)
    got = ub.gather_batch(wl.grad, batch)
    assert torch.allclose(got, g_w, rtol=1e-12, atol=1e-12)
    # complex gradients and a parameter some rank never touches
    c = torch.randn(3, dtype=torch.complex128, requires_grad=True)
    unused = torch.zeros(2, requires_grad=True)
    (c.abs() ** 2).sum().mul(rank + 1).backward()
    ub.all_reduce_gradients([c, unused])
    tot = sum(range(1, world + 1))
    assert torch.allclose(c.grad, 2 * c.detach() * tot)
    assert torch.equal(unused.grad, torch.zeros(2))
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,batch", [(2, 8), (3, 7), (2, 1)])
def test_shard_batch_and_gradient_all_reduce(tmp_path, world, batch):
    port = 29700 + world * 10 + batch
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def test_batch_slice_covers_every_row_once():
    from unitair_b200 import batch as ub
    for batch in (0, 1, 5, 16, 4096):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                s = ub.batch_slice(batch, r, world)
                rows += list(range(s.start, s.stop))
            assert rows == list(range(batch))
    with pytest.raises(ValueError):
        ub.batch_slice(4, 2, 2)
