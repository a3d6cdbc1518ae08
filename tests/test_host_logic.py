"""CPU-only tests: the C-ABI library loads and exports every symbol include/unitair_b200.h
declares, the host shim validates like the reference, the pass planner is correct, and the
product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import unitair_b200 as ua
from unitair_b200 import _lib, circuit
from oracle import unitair_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "unitair_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = sorted(set(re.findall(r"\b(ua_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 15, names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert _lib.lib().ua_version() >= 100


def test_no_cpu_fallback():
    st = torch.zeros(8, dtype=torch.complex64)
    op = torch.eye(2, dtype=torch.complex64)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.simulation.apply_operator(op, (0,), st)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.simulation.apply_all_qubits(op, st)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.simulation.apply_phase(torch.zeros(8), st)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.abs_squared(st)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.diag_expectation_value(torch.zeros(8), st)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.simulation.measure(st, 10)
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.HostCircuitStream(3, torch.complex64, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.circuit.apply_gates([([0], op)], st)


def test_validation_matches_reference_error_types():
    st = torch.zeros(8, dtype=torch.complex64)
    op = torch.eye(2, dtype=torch.complex64)
    for bad in [(3,), (-1,), (0, 1)]:
        with pytest.raises(ValueError):
            ua.simulation.apply_operator(op, bad, st)
    with pytest.raises(ValueError):
        ua.simulation.apply_operator(torch.eye(4, dtype=torch.complex64), (1, 1), st)
    with pytest.raises(ua.states.StateShapeError):
        ua.simulation.apply_operator(op, (0,), torch.zeros(6, dtype=torch.complex64))
    assert issubclass(ua.states.StateShapeError, ValueError)
    with pytest.raises(RuntimeError):
        ua.simulation.apply_operator(torch.zeros(3, 3, dtype=torch.complex64), (0,), st)
    with pytest.raises(ValueError):
        ua.simulation.apply_all_qubits(torch.eye(4, dtype=torch.complex64), st)
    with pytest.raises(ValueError):
        ua.simulation.act_first_qubits(torch.eye(4, dtype=torch.complex64), torch.zeros(2, dtype=torch.complex64))


def test_shape_helpers():
    assert ua.count_qubits(torch.zeros(3, 16)) == 4
    assert ua.states.count_qubits_gate_matrix(torch.zeros(5, 8, 8)) == 3
    t = ua.states.to_tensor_layout(torch.arange(24.).reshape(3, 8))
    assert tuple(t.shape) == (3, 2, 2, 2)
    assert torch.equal(ua.states.to_vector_layout(t, 3), torch.arange(24.).reshape(3, 8))
    x = torch.rand(4, 5, 6, 7, 8, 9)
    assert tuple(ua.states.subset_roll_to_back(x, 2).shape) == (6, 7, 8, 9, 4, 5)
    assert tuple(ua.states.subset_roll_to_front(x, 2).shape) == (8, 9, 4, 5, 6, 7)
    # known answers of the reference's test_state_shapes.py (get_qubit_indices doc examples)
    from unitair_b200.states.shapes import get_qubit_indices
    assert get_qubit_indices(0, torch.rand(2, 2), num_qubits=2) == 0
    assert get_qubit_indices(0, torch.rand(500, 17, 2, 2, 2), num_qubits=3) == 2
    assert get_qubit_indices([1, 0], torch.rand(500, 17, 2, 2, 2), num_qubits=3) == [3, 2]
    assert get_qubit_indices([-1, 0], torch.rand(500, 17, 2, 2, 2), num_qubits=3) == [-1, 2]
    from unitair_b200.utils import permutation_to_front, inverse_list_permutation
    assert permutation_to_front(5, [2]) == [2, 0, 1, 3, 4]
    assert permutation_to_front(5, [2, 1]) == [2, 1, 0, 3, 4]
    assert inverse_list_permutation([1, 2, 3, 4, 0]) == [4, 0, 1, 2, 3]


def _apply_plan_with_oracle(gates, state, passes):
    psi = state
    for p in passes:
        for g in p.gates:
            qs, u = gates[g]
            psi = orc.apply_operator(u, qs, psi)
    return psi


@pytest.mark.parametrize("n,dtype", [(6, torch.complex64), (12, torch.complex64), (14, torch.complex128)])
def test_pass_planner_preserves_the_circuit(n, dtype):
    rng = np.random.default_rng(n)
    gates = []
    for layer in range(4):
        for q in range(n):
            gates.append(([q], rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))))
        pi = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            gates.append(([pi[j], pi[j + 1]], rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))))
        gates.append((rng.permutation(n)[:4].tolist(), rng.standard_normal((16, 16)) + 0j))
    gates = [(qs, (u / np.linalg.norm(u, 2)).astype(np.complex128)) for qs, u in gates]
    geo = circuit.TileGeometry(n, min(n, 8), min(n, 4), min(n, 8) - min(n, 4))
    bits = [[n - 1 - q for q in qs] for qs, _ in gates]
    passes = circuit.plan_passes(bits, geo, max_gates=9)
    seen = sorted(g for p in passes for g in p.gates)
    assert seen == list(range(len(gates)))            # every gate exactly once
    for p in passes:
        if p.direct:
            assert len(p.gates) == 1
            continue
        assert len(p.gates) <= 9
        assert len(p.high) == geo.max_high and all(b >= geo.low_bits for b in p.high)
        tile = set(range(geo.low_bits)) | set(p.high)
        for g in p.gates:
            assert set(bits[g]) <= tile
    state = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    ref = state
    for qs, u in gates:
        ref = orc.apply_operator(u, qs, ref)
    got = _apply_plan_with_oracle(gates, state, passes)
    assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref)
    assert len(passes) < len(gates) / 2


def test_default_geometry():
    g = circuit.default_geometry(30, torch.complex64)
    assert g.tile_bits == 13 and g.low_bits == 7 and g.max_high == 6
    g = circuit.default_geometry(5, torch.complex64)
    assert g.tile_bits == 5 and g.low_bits == 5 and g.max_high == 0
    g = circuit.default_geometry(30, torch.complex128)
    assert g.tile_bits == 12 and g.low_bits == 6


def test_tail_first_planner_preserves_the_circuit_and_avoids_forbidden_bits():
    n = 12
    rng = np.random.default_rng(5)
    gates = []
    for layer in range(5):
        for q in range(n):
            gates.append(([q], rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))))
        pi = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            gates.append(([pi[j], pi[j + 1]], rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))))
    gates = [(qs, (u / np.linalg.norm(u, 2)).astype(np.complex128)) for qs, u in gates]
    geo = circuit.TileGeometry(n, 8, 4, 4)
    bits = [[n - 1 - q for q in qs] for qs, _ in gates]
    for forbidden in ([9], [5, 11], [4, 6, 10]):
        passes = circuit.plan_passes_tail_first(bits, geo, forbidden)
        assert sorted(g for p in passes for g in p.gates) == list(range(len(gates)))
        last = passes[-1]
        used_last = {b for g in last.gates for b in bits[g]}
        touches = bool(used_last & set(forbidden))
        # the last gates of this circuit may themselves touch a forbidden bit (then the ban is
        # lifted and the caller adds a copy pass); otherwise the last tile avoids them
        if not touches:
            assert not (set(last.high) & set(forbidden))
        state = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        ref = state
        for qs, u in gates:
            ref = orc.apply_operator(u, qs, ref)
        got = _apply_plan_with_oracle(gates, state, passes)
        assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref)


def test_scatter_tail_copy_pass_geometry():
    """Planning of the pass that scatters a shard to the peers (no CUDA needed for the plan):
    the tile never contains a leaving bit, is full, and keeps the low bits together."""
    for n, victims, dtype in [(30, [10, 19, 28], torch.complex64), (30, [27, 28, 29], torch.complex64),
                              (33, [7], torch.complex64), (20, [12, 13], torch.complex128), (12, [10, 11], torch.complex64)]:
        tail = circuit.ScatterTail(None, n, dtype, victims)
        base = circuit.default_geometry(n, dtype)
        tile = min(base.tile_bits, n - len(victims))
        low = min(base.low_bits, tile)
        assert not tail.reused and tail.num_passes == 1 and tail.launch.ngates == 0
        high = list(tail.launch.high) if tail.launch.high is not None else []
        assert tail.launch.low + len(high) == tile and tail.launch.low == tile - len(high)
        assert len(high) == tile - low
        assert not (set(high) & set(victims)) and all(low <= b < n for b in high)
        assert high == sorted(high)
    with pytest.raises(ValueError):
        circuit.ScatterTail(None, 30, torch.complex64, [3])          # a leaving bit among the low bits
    assert circuit._fill_high({9}, circuit.TileGeometry(10, 8, 7, 1), forbidden=[]) == [9]
    assert circuit._fill_high(set(), circuit.TileGeometry(9, 9, 7, 2), forbidden=[8]) is None


def test_plan_passes_never_spins_when_a_gate_cannot_fit():
    """a gate that fits no tile (no room for its high bits) must become a direct
    launch instead of looping forever."""
    geo = circuit.TileGeometry(12, 5, 4, 1)              # one free high slot only
    gate_bits = [[0, 1], [8, 9], [2, 10], [9, 11, 3]]     # the 2nd and 4th need two high bits
    passes = circuit.plan_passes(gate_bits, geo)
    seen = sorted(g for p in passes for g in p.gates)
    assert seen == [0, 1, 2, 3]
    assert any(p.direct and p.gates == [1] for p in passes)


def test_cluster_geometry_and_window_split():
    g = circuit.default_geometry(30, torch.complex64, cluster=True)
    assert (g.tile_bits, g.low_bits, g.max_high, g.split_low) == (12, 7, 5, 4)
    # the first TMA dimension is exactly bits 0..3; bits 4.. continue as before
    assert circuit.count_windows(7, [7, 8, 9, 10, 11], 0, split_low=4) == 2
    assert circuit.count_windows(7, [7, 8, 9, 10, 11], 0) == 2          # 0..7 and 8..11
    assert circuit.count_windows(7, [9, 11, 13, 15, 17], 0, split_low=4) == 7
    small = circuit.default_geometry(3, torch.complex64, cluster=True)
    assert small.split_low == 0 and small.tile_bits == 3
