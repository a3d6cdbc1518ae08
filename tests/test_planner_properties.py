"""Hypothesis property tests (CPU) of the host-side planners: whatever the gate list, merging
and pass / epoch planning must leave the circuit's product unchanged.  Checked with the numpy
oracle on small states; modelled on the reference's own property-test style
(tests/hypothesis_strategies/, tests/test_operations.py)."""
import numpy as np
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import unitair_oracle as orc
from unitair_b200 import circuit, sharded

SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))


@st.composite
def gate_lists(draw, max_qubits=7, max_gates=24, max_k=3):
    n = draw(st.integers(3, max_qubits))
    count = draw(st.integers(1, max_gates))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(count):
        k = int(rng.integers(1, min(max_k, n) + 1))
        qs = rng.choice(n, k, replace=False).tolist()
        u = rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))
        gates.append((qs, (u / np.linalg.norm(u, 2)).astype(np.complex128)))
    state = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    return n, gates, state / np.linalg.norm(state)


def _run(gates, state):
    psi = state
    for qs, u in gates:
        psi = orc.apply_operator(np.asarray(u), list(qs), psi)
    return psi


def _close(a, b):
    return np.linalg.norm(a - b) <= 1e-10 * max(1.0, np.linalg.norm(b))


@settings(**SETTINGS)
@given(gate_lists(), st.sampled_from([2, 3]))
def test_merge_gates_preserves_the_product(case, max_k):
    n, gates, state = case
    merged = circuit.merge_gates([(qs, torch.from_numpy(u)) for qs, u in gates], max_k)
    assert all(len(qs) <= max(max_k, 3) for qs, _ in merged)
    assert len(merged) <= len(gates)
    assert _close(_run([(qs, u.numpy()) for qs, u in merged], state), _run(gates, state))


@settings(**SETTINGS)
@given(gate_lists(max_qubits=8), st.integers(0, 3), st.booleans())
def test_pass_planners_preserve_the_product(case, low, tail_first):
    n, gates, state = case
    tile = min(n, low + 3)
    geo = circuit.TileGeometry(n, tile, min(low, tile), tile - min(low, tile))
    bits = [[n - 1 - q for q in qs] for qs, _ in gates]
    forbidden = [n - 1] if (tail_first and geo.low_bits < n - 1 and tile < n) else []
    if tail_first and forbidden:
        passes = circuit.plan_passes_tail_first(bits, geo, forbidden, max_gates=7)
    else:
        passes = circuit.plan_passes(bits, geo, max_gates=7)
    assert sorted(g for p in passes for g in p.gates) == list(range(len(gates)))
    for p in passes:
        if p.direct:
            continue
        tile_bits = set(range(geo.low_bits)) | set(p.high)
        assert len(p.gates) <= 7 and all(set(bits[g]) <= tile_bits for g in p.gates)
    ordered = [gates[g] for p in passes for g in p.gates]
    assert _close(_run(ordered, state), _run(gates, state))


@settings(**SETTINGS)
@given(gate_lists(max_qubits=8, max_k=2), st.integers(1, 2), st.integers(0, 3), st.booleans())
def test_epoch_planner_respects_dependencies_and_locality(case, g, min_victim, restore):
    n, gates, _ = case
    g = min(g, n - 2)
    gq = [qs for qs, _ in gates]
    epochs, end = sharded.plan_epochs(gq, n, g, restore=restore, min_victim_bit=min_victim)
    order = [gi for e in epochs for gi in e.gates]
    assert sorted(order) == list(range(len(gq)))
    pos = {gi: i for i, gi in enumerate(order)}
    last = {}
    for gi, qs in enumerate(gq):
        for q in qs:
            if q in last:
                assert pos[last[q]] < pos[gi]
            last[q] = gi
    for e in epochs:
        assert len(e.incoming) == len(e.victims) == len(e.rank_bits) == len(e.victim_bits) <= g
        assert e.victim_bits == sorted(e.victim_bits)
        for b in e.local_bits:
            assert all(0 <= x < n - g for x in b)
    if restore:
        assert end == sharded.identity_layout(n)


# ------------------------------------------------------------------------------------------
# The epoch plan as data: simulate it on the FULL state in one process (physical index =
# rank bits on top of the local bits) in both exchange formulations and compare with the
# circuit itself.  Pins the meaning of Epoch.perm_src / rank_bits / victim_bits / local_bits.
def _permute_local_bits(full, n, n_local, src):
    """out bit p takes in bit src[p] (local bits only) -- ua_permute_bits semantics."""
    t = full.reshape((2,) * n)                       # axis i <-> physical bit n-1-i
    axes = list(range(n))
    for p, s_ in enumerate(src):
        axes[n - 1 - p] = n - 1 - s_
    return np.transpose(t, axes).reshape(-1)


def _exchange_bits(full, n, n_local, rank_bits):
    """swap rank bit rank_bits[j] with local bit n_local - m + j (the block exchange)."""
    m = len(rank_bits)
    t = full.reshape((2,) * n)
    axes = list(range(n))
    for j, rb in enumerate(rank_bits):
        a, b = n - 1 - (n_local + rb), n - 1 - (n_local - m + j)
        axes[a], axes[b] = axes[b], axes[a]
    return np.transpose(t, axes).reshape(-1)


def _scatter(full, n, n_local, ep):
    """The fused formulation: amplitude (rank, i) goes to rank' = rank with the swapped rank bits
    replaced by the victim-bit values of i, position (own swapped rank bits) * 2^(nl-m) +
    (i with the victim bits squeezed out)  (ua_apply_fused_pass_scatter + CudaEngine.run_scatter)."""
    m = len(ep.incoming)
    out = np.empty_like(full)
    for rank in range(1 << (n - n_local)):
        a = sharded.exchange_block_id(rank, ep)
        for i in range(1 << n_local):
            b = 0
            for j, v in enumerate(ep.victim_bits):
                b |= ((i >> v) & 1) << j
            off = i
            for v in sorted(ep.victim_bits, reverse=True):
                off = ((off >> (v + 1)) << v) | (off & ((1 << v) - 1))
            peer = sharded.exchange_peer(rank, ep, b)
            out[(peer << n_local) | (a << (n_local - m)) | off] = full[(rank << n_local) | i]
    return out


def _simulate_plan(epochs, end_layout, gates, n, g, state, fused):
    n_local = n - g
    full = np.array(state)                             # identity layout: physical == logical
    for ep in epochs:
        if ep.incoming and fused:
            full = _scatter(full, n, n_local, ep)
        else:
            if ep.perm_src is not None:
                full = _permute_local_bits(full, n, n_local, ep.perm_src)
            if ep.incoming:
                full = _exchange_bits(full, n, n_local, ep.rank_bits)
        for gi, bits in zip(ep.gates, ep.local_bits):
            full = orc.apply_operator(gates[gi][1], [n - 1 - p for p in bits], full)
    t = full.reshape((2,) * n)
    axes = [n - 1 - end_layout[q] for q in range(n)]   # logical qubit q sits on physical bit layout[q]
    return np.transpose(t, axes).reshape(-1)


@settings(**SETTINGS)
@given(gate_lists(max_qubits=7, max_k=2, max_gates=18), st.integers(1, 2), st.integers(0, 2), st.booleans(),
       st.booleans())
def test_epoch_plan_reproduces_the_circuit_in_both_exchange_formulations(case, g, min_victim, restore, fused):
    n, gates, state = case
    g = min(g, n - 2)
    epochs, end = sharded.plan_epochs([qs for qs, _ in gates], n, g, restore=restore, min_victim_bit=min_victim)
    got = _simulate_plan(epochs, end, gates, n, g, state, fused)
    assert _close(got, _run(gates, state))


@settings(**SETTINGS)
@given(gate_lists(max_qubits=7, max_k=2, max_gates=14), st.integers(1, 2), st.integers(2, 3), st.booleans())
def test_joint_plan_of_repeated_steps_equals_chained_per_step_plans(case, g, steps, fused):
    """bench.py plans the K timed steps of a sharded run as ONE gate list (the leftover gates of a
    step share the next step's exchange).  That plan and K per-step plans chained through the qubit
    layout must both reproduce K applications of the circuit, and the joint plan never needs more
    exchanges than the chain."""
    n, gates, state = case
    g = min(g, n - 2)
    qubits = [qs for qs, _ in gates]
    want = np.array(state)
    for _ in range(steps):
        want = _run(gates, want)
    epochs, end = sharded.plan_epochs(qubits * steps, n, g, restore=False)
    got = _simulate_plan(epochs, end, gates * steps, n, g, state, fused)
    assert _close(got, want)
    joint_swaps = sum(1 for ep in epochs if ep.incoming)
    # chained per-step plans: each starts in the layout the previous one ended in; simulated on the
    # physical state by composing the plans' epoch lists
    layout, chain_swaps = None, 0
    phys = None
    all_epochs = []
    for _ in range(steps):
        eps, layout = sharded.plan_epochs(qubits, n, g, layout=layout, restore=False)
        chain_swaps += sum(1 for ep in eps if ep.incoming)
        all_epochs.append(eps)
    full = np.array(state)
    n_local = n - g
    for eps in all_epochs:
        for ep in eps:
            if ep.incoming and fused:
                full = _scatter(full, n, n_local, ep)
            else:
                if ep.perm_src is not None:
                    full = _permute_local_bits(full, n, n_local, ep.perm_src)
                if ep.incoming:
                    full = _exchange_bits(full, n, n_local, ep.rank_bits)
            for gi, bits in zip(ep.gates, ep.local_bits):
                full = orc.apply_operator(gates[gi][1], [n - 1 - p for p in bits], full)
    t = full.reshape((2,) * n)
    chained = np.transpose(t, [n - 1 - layout[q] for q in range(n)]).reshape(-1)
    assert _close(chained, want)
    assert joint_swaps <= chain_swaps
