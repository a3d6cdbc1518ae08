"""Multi-process CPU tests (gloo, world_size 2 and 4) of the sharded-state path: epoch
planning, victim permutation, block exchange, layout restore.  The local gate work is done
by a test engine built on the numpy oracle (the product engine is CUDA-only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import unitair_oracle as orc
from unitair_b200 import sharded


class OracleEngine:
    """Test-only local engine: applies gates / bit permutations to a CPU shard with the oracle."""

    def compile(self, gates_local, n_local, tail_victims=None):
        return [(qs, m.numpy()) for qs, m in gates_local]

    def num_passes(self, compiled):
        return len(compiled)

    def run(self, compiled, shard):
        psi = shard.numpy()
        for qs, u in compiled:
            psi = orc.apply_operator(u, qs, psi)
        shard.copy_(torch.from_numpy(np.ascontiguousarray(psi)))
        return shard

    def permute(self, src_bits, shard, out):
        n = len(src_bits)
        t = shard.numpy().reshape((2,) * n)          # axis i <-> bit n-1-i
        # output bit p takes input bit src[p]  =>  output axis (n-1-p) is input axis (n-1-src[p])
        axes = [0] * n
        for p, s in enumerate(src_bits):
            axes[n - 1 - p] = n - 1 - s
        out.copy_(torch.from_numpy(np.ascontiguousarray(np.transpose(t, axes)).reshape(-1)))
        return out


class ScatterOracleEngine(OracleEngine):
    """Same, plus a CPU stand-in for the fused scatter exchange: the epoch's gates, then every
    block is delivered straight into the peers' SPARE buffers (send/recv here, peer-memory
    stores in the CUDA engine).  Exercises the planner's victim constraints and the
    fence / buffer-flip logic of ShardedCircuit.run."""

    supports_scatter = True
    LOW = 2

    def min_victim_bit(self, n_local, g):
        return min(self.LOW, max(0, n_local - g))

    def scatter_tail(self, compiled, n_local, victim_bits):
        assert list(victim_bits) == sorted(victim_bits) and min(victim_bits) >= self.min_victim_bit(n_local, 0)
        tail = type("Tail", (), {})()
        tail.compiled, tail.victims, tail.num_passes = compiled, list(victim_bits), len(compiled or []) + 1
        return tail

    def run_scatter(self, tail, st, ep, before_scatter=None):
        self.run(tail.compiled or [], st.local)
        if before_scatter is not None:
            before_scatter()
        nl, m = st.n_local, len(tail.victims)
        idx = np.arange(1 << nl)
        block = np.zeros_like(idx)
        for j, v in enumerate(tail.victims):
            block |= ((idx >> v) & 1) << j
        off = idx.copy()
        for v in sorted(tail.victims, reverse=True):
            off = ((off >> (v + 1)) << v) | (off & ((1 << v) - 1))
        staged = np.empty(1 << nl, dtype=np.complex128)
        staged[block * (1 << (nl - m)) + off] = st.local.numpy()
        staged = torch.from_numpy(staged)
        bs = 1 << (nl - m)
        a = sharded.exchange_block_id(st.rank, ep)
        spare = st._spare()
        ops = []
        for b in range(1 << m):
            if b == a:
                spare[a * bs:(a + 1) * bs].copy_(staged[a * bs:(a + 1) * bs])
                continue
            peer = sharded.exchange_peer(st.rank, ep, b)
            ops.append(dist.P2POp(dist.isend, torch.view_as_real(staged[b * bs:(b + 1) * bs]), peer))
            ops.append(dist.P2POp(dist.irecv, torch.view_as_real(spare[b * bs:(b + 1) * bs]), peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _circuit(n, layers, seed):
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(layers):
        for q in range(n):
            gates.append(([q], (rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))) / 1.5))
        perm = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            gates.append(([perm[j], perm[j + 1]],
                          (rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))) / 2.5))
    gates.append((rng.permutation(n)[:3].tolist(), (rng.standard_normal((8, 8)) + 0j) / 3))
    return [(qs, u.astype(np.complex128)) for qs, u in gates]


def _worker(rank, world, port, n, layers, restore, result_path, scatter=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        g = world.bit_length() - 1
        gates_np = _circuit(n, layers, seed=7)
        rng = np.random.default_rng(3)
        full = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        full /= np.linalg.norm(full)
        nl = n - g
        local = torch.from_numpy(full[rank << nl:(rank + 1) << nl].copy())
        st = sharded.ShardedState(local, n)
        gates = [(qs, torch.from_numpy(u)) for qs, u in gates_np]
        engine = ScatterOracleEngine() if scatter else OracleEngine()
        sc = sharded.ShardedCircuit(gates, n, torch.complex128, world, engine=engine, restore=restore,
                                    exchange="p2p" if scatter else "nccl")
        assert (sc.num_fused_swaps > 0) == scatter
        sc.run(st)
        if restore:
            assert st.layout == sharded.identity_layout(n)
            sc.run(st)                      # the plan is replayable
        nrm = st.norm_squared()
        got = st.gather_logical()
        if rank == 0:
            ref = full
            for rep in range(2 if restore else 1):
                for qs, u in gates_np:
                    ref = orc.apply_operator(u, qs, ref)
            err = np.linalg.norm(got.numpy() - ref) / np.linalg.norm(ref)
            nerr = abs(float(nrm) - float(np.vdot(ref, ref).real)) / float(np.vdot(ref, ref).real)
            with open(result_path, "w") as f:
                f.write(f"{err} {nerr} {sc.num_swaps}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,restore", [(2, 6, True), (2, 7, False), (4, 7, True), (4, 8, False)])
def test_sharded_circuit_matches_oracle(tmp_path, world, n, restore):
    path = str(tmp_path / "res.txt")
    mp.spawn(_worker, args=(world, _free_port(), n, 3, restore, path), nprocs=world, join=True)
    err, nerr, swaps = open(path).read().split()
    assert float(err) < 1e-12, err
    assert float(nerr) < 1e-12, nerr
    assert int(swaps) >= 1


@pytest.mark.parametrize("world,n,restore", [(2, 7, True), (4, 8, True), (4, 9, False), (8, 9, False)])
def test_sharded_circuit_scatter_exchange_matches_oracle(tmp_path, world, n, restore):
    """The fused-scatter control flow (exchange delivered by the previous epoch's last pass into
    the peers' spare buffers, fence, buffer flip; send/recv fallback for restore exchanges that
    must evict a low bit) against the oracle."""
    path = str(tmp_path / "res.txt")
    mp.spawn(_worker, args=(world, _free_port(), n, 3, restore, path, True), nprocs=world, join=True)
    err, nerr, swaps = open(path).read().split()
    assert float(err) < 1e-12, err
    assert float(nerr) < 1e-12, nerr
    assert int(swaps) >= 1


def test_epoch_planner_min_victim_bit():
    rng = np.random.default_rng(1)
    n, g, low = 14, 3, 5
    gq = []
    for _ in range(6):
        gq += [[q] for q in range(n)]
        perm = rng.permutation(n).tolist()
        gq += [[perm[j], perm[j + 1]] for j in range(0, n - 1, 2)]
    epochs, end = sharded.plan_epochs(gq, n, g, restore=False, min_victim_bit=low)
    assert sorted(gi for e in epochs for gi in e.gates) == list(range(len(gq)))
    swaps = [e for e in epochs if e.incoming]
    assert swaps
    for e in swaps:
        assert e.victim_bits == sorted(e.victim_bits) and min(e.victim_bits) >= low
        assert len(e.victim_bits) == len(e.rank_bits)


def test_epoch_planner_properties():
    rng = np.random.default_rng(0)
    for n, g in [(8, 1), (9, 2), (10, 3), (12, 3)]:
        gq = []
        for _ in range(5):
            gq += [[q] for q in range(n)]
            perm = rng.permutation(n).tolist()
            gq += [[perm[j], perm[j + 1]] for j in range(0, n - 1, 2)]
        epochs, end = sharded.plan_epochs(gq, n, g, restore=True)
        assert end == sharded.identity_layout(n)
        assert sorted(gi for e in epochs for gi in e.gates) == list(range(len(gq)))
        for e in epochs:
            assert len(e.incoming) == len(e.victims) == len(e.rank_bits) <= g
            assert all(0 <= rb < g for rb in e.rank_bits)
            for bits in e.local_bits:
                assert all(0 <= b < n - g for b in bits)
        # every gate runs after all earlier gates it shares a qubit with
        order = [gi for e in epochs for gi in e.gates]
        pos = {gi: i for i, gi in enumerate(order)}
        last = {}
        for gi, qs in enumerate(gq):
            for q in qs:
                if q in last:
                    assert pos[last[q]] < pos[gi]
                last[q] = gi
