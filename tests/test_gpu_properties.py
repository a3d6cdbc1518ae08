"""Property tests on the GPU, modelled on the reference's own Hypothesis suite
(tests/test_operations.py:23-201 and tests/test_unitary.py of qcware/qcware-unitair): same
properties, own strategies, and every drawn case is also compared with the oracle."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import assert_close
from oracle import unitair_oracle as orc

pytestmark = pytest.mark.gpu
SET = settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))


@pytest.fixture(scope="module")
def ua():
    import unitair_b200
    from unitair_b200 import _lib
    _lib.lib()
    return unitair_b200


def dev(x):
    return torch.from_numpy(np.array(x, copy=True)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@st.composite
def batch_dims(draw, max_rank=3, max_size=3):
    return tuple(draw(st.lists(st.integers(1, max_size), min_size=0, max_size=max_rank)))


@st.composite
def states_and_ops(draw, max_qubits=8, max_op_qubits=5, op_max_abs=10.0):
    """(operator, state): operator batch is empty, equal to the state batch, or the state is
    unbatched (the three documented structures, operations.py:88-112)."""
    n = draw(st.integers(1, max_qubits))
    k = draw(st.integers(1, min(n, max_op_qubits)))
    sb = draw(batch_dims())
    mode = draw(st.sampled_from(["shared", "same", "op_only"]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    if mode == "shared":
        ob = ()
    elif mode == "same":
        ob = sb
    else:
        ob, sb = sb, ()
    state = rng.standard_normal(sb + (2 ** n,)) + 1j * rng.standard_normal(sb + (2 ** n,))
    state = (state / np.linalg.norm(state, axis=-1, keepdims=True)).astype(np.complex64)
    op = rng.uniform(-op_max_abs, op_max_abs, ob + (2 ** k, 2 ** k)) + 1j * rng.uniform(-op_max_abs, op_max_abs, ob + (2 ** k, 2 ** k))
    return op.astype(np.complex64), state, n, k, rng


@SET
@given(data=states_and_ops())
def test_apply_operator_matches_act_first_qubits_and_oracle(ua, data):
    op, state, n, k, rng = data
    a = ua.simulation.act_first_qubits(operator=dev(op), state=dev(state))
    b = ua.simulation.apply_operator(operator=dev(op), qubits=range(k), state=dev(state))
    assert torch.equal(a, b)
    assert_close(host(b), orc.apply_operator(op, range(k), state), "c64", factor=3)
    qs = rng.permutation(n)[:k].tolist()
    c = ua.simulation.apply_operator(operator=dev(op), qubits=qs, state=dev(state))
    assert_close(host(c), orc.apply_operator(op, qs, state), "c64", factor=3, what=f"qubits={qs}")


@SET
@given(data=states_and_ops(op_max_abs=1.0))
def test_batch_entries_equal_unbatched_calls(ua, data):
    op, state, n, k, rng = data
    out = ua.simulation.act_first_qubits(operator=dev(op), state=dev(state))
    batch = out.shape[:-1]
    if len(batch) == 0:
        return
    idx = tuple(int(rng.integers(0, s)) for s in batch)
    op_e = op[idx] if op.ndim > 2 else op
    st_e = state[idx] if state.ndim > 1 else state
    single = ua.simulation.act_first_qubits(operator=dev(op_e), state=dev(st_e))
    assert torch.allclose(out[idx], single, rtol=1e-5, atol=1e-6)


@SET
@given(data=states_and_ops(max_op_qubits=1, op_max_abs=1.0))
def test_apply_all_qubits_batching_and_oracle(ua, data):
    op, state, n, k, rng = data
    out = ua.simulation.apply_all_qubits(operator=dev(op), state=dev(state))
    assert_close(host(out), orc.apply_all_qubits(op, state) if op.ndim == 2 or state.ndim > 1 else
                 np.stack([orc.apply_all_qubits(o, state) for o in op.reshape(-1, 2, 2)]).reshape(out.shape),
                 "c64", factor=5)
    batch = out.shape[:-1]
    if len(batch) and state.ndim > 1:
        idx = tuple(int(rng.integers(0, s)) for s in batch)
        single = ua.simulation.apply_all_qubits(dev(op[idx] if op.ndim > 2 else op), dev(state[idx]))
        assert torch.allclose(out[idx], single, rtol=1e-5, atol=1e-6)


@SET
@given(n=st.integers(1, 8), sb=batch_dims(), seed=st.integers(0, 2 ** 31 - 1))
def test_phase_then_inverse_phase_is_identity(ua, n, sb, seed):
    rng = np.random.default_rng(seed)
    state = (rng.standard_normal(sb + (2 ** n,)) + 1j * rng.standard_normal(sb + (2 ** n,))).astype(np.complex64)
    angles = rng.uniform(-20, 20, sb + (2 ** n,)).astype(np.float32)
    s, a = dev(state), dev(angles)
    back = ua.simulation.apply_phase(-a, ua.simulation.apply_phase(a, s))
    assert torch.isclose(back, s, atol=1e-4).all()
    comp = ua.simulation.apply_phase(2 * np.pi - a, ua.simulation.apply_phase(a, s))
    assert torch.isclose(comp, s, atol=1e-4 + float(np.abs(angles).sum()) * .001).all()
    assert_close(host(ua.simulation.apply_phase(a, s)), orc.apply_phase(angles, state), "c64", factor=2)


@SET
@given(n=st.integers(2, 8), sb=batch_dims(), seed=st.integers(0, 2 ** 31 - 1))
def test_swap_properties(ua, n, sb, seed):
    rng = np.random.default_rng(seed)
    state = dev((rng.standard_normal(sb + (2 ** n,)) + 1j * rng.standard_normal(sb + (2 ** n,))).astype(np.complex64))
    i, j = rng.permutation(n)[:2].tolist()
    swapped = ua.simulation.swap(state, qubit_pair=(i, j))
    assert torch.equal(ua.simulation.swap(swapped, qubit_pair=(i, j)), state)        # involutory
    assert torch.equal(ua.simulation.swap(state, qubit_pair=(j, i)), swapped)        # symmetric
    perm = list(range(n))
    perm[i], perm[j] = j, i
    assert torch.equal(ua.simulation.permute_qubits(permutation=perm, state_vector=state), swapped)
    assert np.array_equal(host(swapped), orc.swap(host(state), (i, j)))
