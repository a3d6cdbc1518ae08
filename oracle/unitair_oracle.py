"""CPU oracle for unitair's gate-application hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (qcware/qcware-unitair
v0.3.0).  Nothing in the product package (``qcware-unitair_b200/``) may import it;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and there only as the checker or the thing timed as
the CPU baseline -- never as the shipped path.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the
unmodified reference (``/root/reference/src``) in the build container with
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks every
function below against those vectors and against the known-answer vectors in the
reference's docs (README.rst:165-166, 223-224; docs/tutorial/first_example.rst).

The restatement follows the reference step by step (permute targets to the front,
make contiguous, contract, permute back) so that it is also a fair "port" CPU
baseline.  File:line citations are relative to ``/root/reference/``.

The arithmetic of the reference lives in a third-party dependency, PyTorch ATen
(``torch>=1.8.1``, unpinned in setup.cfg:20; 2.11.0+cu128 installed here):
``einsum``->``bmm`` (operations.py:322), ``exp``/``mul`` (operations.py:41-42),
``sum`` (innerprod.py:46,59,65).  Here they are restated with numpy ``matmul``,
``exp``, ``sum``.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Sequence

import numpy as np


class StateShapeError(ValueError):
    """Mirror of src/unitair/states/shapes.py:149 (a ValueError subclass)."""


# --------------------------------------------------------------------------- #
# shape helpers  (src/unitair/states/shapes.py, src/unitair/utils.py)
# --------------------------------------------------------------------------- #
def count_qubits(state: np.ndarray) -> int:
    """shapes.py:13-27 -- round(log2(last dim)); error if not a power of two."""
    length = state.shape[-1]
    num_bits = round(math.log2(length)) if length > 0 else 0
    if length <= 0 or 2 ** num_bits != length:
        raise StateShapeError(f"last dim {length} is not a power of two")
    return num_bits


def count_qubits_gate_matrix(gate: np.ndarray) -> int:
    """shapes.py:31-47 -- RuntimeError if the last dim is not 2**k."""
    length = gate.shape[-1]
    num_bits = round(math.log2(length)) if length > 0 else 0
    if length <= 0 or 2 ** num_bits != length:
        raise RuntimeError(f"gate size {gate.shape} is not consistent with qubits")
    return num_bits


def count_gate_batch_dims(gate: np.ndarray) -> int:
    """src/unitair/simulation/utils.py:4-20."""
    out = gate.ndim - 2
    if out < 0:
        raise RuntimeError(f"gate with size {gate.shape} is incorrectly shaped")
    return out


def permutation_to_front(n: int, entries: Sequence[int]) -> List[int]:
    """src/unitair/utils.py:5-27 -- listed entries first (in the given order)."""
    arrangement = list(range(n))
    for q in entries:
        del arrangement[arrangement.index(q)]  # ValueError on duplicates
    return list(entries) + arrangement


def inverse_list_permutation(perm: Sequence[int]) -> List[int]:
    """src/unitair/utils.py:30-49."""
    inverse = [0] * len(perm)
    for a, b in enumerate(perm):
        inverse[b] = a
    return inverse


# --------------------------------------------------------------------------- #
# layouts  (src/unitair/states/conversions.py)
# --------------------------------------------------------------------------- #
def to_tensor_layout(state: np.ndarray) -> np.ndarray:
    """conversions.py:5-45 -- (*B, 2**n) -> (*B, 2, ..., 2), contiguous."""
    n = count_qubits(state)
    state = np.ascontiguousarray(state)
    return state.reshape(state.shape[:-1] + (2,) * n)


def to_vector_layout(state_tensor: np.ndarray, num_qubits: int) -> np.ndarray:
    """conversions.py:48-92 -- (*B, 2, ..., 2) -> (*B, 2**n), contiguous."""
    state_tensor = np.ascontiguousarray(state_tensor)
    batch = state_tensor.shape[: state_tensor.ndim - num_qubits]
    return state_tensor.reshape(batch + (2 ** num_qubits,))


def permute_qubits_tensor(permutation, state_tensor, num_qubits, contiguous_output=False):
    """operations.py:626-654 -- permute qubit axes, leave batch axes alone."""
    non_qubit_dims = state_tensor.ndim - num_qubits
    axes = tuple(range(non_qubit_dims)) + tuple(i + non_qubit_dims for i in permutation)
    out = np.transpose(state_tensor, axes)
    return np.ascontiguousarray(out) if contiguous_output else out


# --------------------------------------------------------------------------- #
# the hot path  (src/unitair/simulation/operations.py)
# --------------------------------------------------------------------------- #
def act_first_qubits_tensor(operator, state_tensor, num_qubits, gate_num_qubits=None):
    """operations.py:258-329 -- out[a, r] = sum_b U[a, b] psi[b, r] on leading qubits.

    The three batch structures of operations.py:277-287 (plus right-aligned
    broadcasting, which the reference gets from einsum) are reproduced by rolling
    batch dims to the back (:311-312) and contracting with a broadcasting matmul.
    """
    if gate_num_qubits is None:
        gate_num_qubits = count_qubits_gate_matrix(operator)
    gate_dim = 2 ** gate_num_qubits
    state_nb = state_tensor.ndim - num_qubits
    op_nb = count_gate_batch_dims(operator)
    state_batch = state_tensor.shape[:state_nb]
    op_batch = operator.shape[:op_nb]

    # operations.py:304-309 -- batched operator on a single state: expand the state
    if op_nb > 0 and state_nb == 0:
        state_tensor = np.broadcast_to(state_tensor, op_batch + state_tensor.shape)
        state_nb = op_nb
        state_batch = op_batch

    if op_nb > state_nb:
        # the reference's einsum refuses an operator with more batch dims than the state
        raise RuntimeError("operator batch dims are not broadcastable to the state batch dims")
    # einsum's '...' broadcasting is right-aligned and accepts size-1 dims
    np.broadcast_shapes(op_batch, state_batch[len(state_batch) - op_nb:])
    out_batch = np.broadcast_shapes(op_batch, state_batch)
    if out_batch != tuple(state_batch):
        state_tensor = np.broadcast_to(
            state_tensor, out_batch + state_tensor.shape[state_nb:])
        state_batch = out_batch

    rest = 2 ** (num_qubits - gate_num_qubits)
    # view as (*B, 2**k, rest) and contract: matmul broadcasts the operator over B
    psi = np.ascontiguousarray(state_tensor).reshape(tuple(state_batch) + (gate_dim, rest))
    out = np.matmul(operator, psi)
    return out.reshape(tuple(state_batch) + (2,) * num_qubits)


def apply_operator_tensor(operator, qubits, state_tensor, num_qubits, operator_num_qubits=None):
    """operations.py:151-186 -- permute targets to the front, contract, permute back."""
    qubits = list(qubits)
    if operator_num_qubits is None:
        operator_num_qubits = count_qubits_gate_matrix(operator)
    perm = permutation_to_front(num_qubits, qubits)
    inv_perm = inverse_list_permutation(perm)
    state_tensor = permute_qubits_tensor(perm, state_tensor, num_qubits, contiguous_output=True)
    state_tensor = act_first_qubits_tensor(operator, state_tensor, num_qubits, operator_num_qubits)
    return permute_qubits_tensor(inv_perm, state_tensor, num_qubits, contiguous_output=True)


def apply_operator(operator, qubits: Iterable[int], state):
    """operations.py:45-148 -- validated front for a dense k-qubit operator."""
    operator = np.asarray(operator)
    state = np.asarray(state)
    num_qubits = count_qubits(state)
    qubits = list(qubits)
    if not set(qubits).issubset(range(num_qubits)):          # :128-132
        raise ValueError(f"qubits={qubits} is not consistent with {num_qubits} qubits")
    op_num_qubits = count_qubits_gate_matrix(operator)       # :133
    if len(qubits) != op_num_qubits:                         # :134-138
        raise ValueError(f"cannot apply a {op_num_qubits}-qubit operator to {qubits}")
    if operator.dtype != state.dtype:
        raise RuntimeError("expected operator and state to have the same dtype")
    st = to_tensor_layout(state)
    st = apply_operator_tensor(operator, qubits, st, num_qubits, op_num_qubits)
    return to_vector_layout(st, num_qubits)


def apply_all_qubits(operator, state):
    """operations.py:332-413 -- the same 2x2 (shared or batched) on every qubit.

    The reference walks q = 0..n-1 with a swap-to-front view and un-rolls at the
    end (:395-411); the result equals apply_operator(op, (q,), .) for q = 0..n-1
    in sequence, which is how it is restated here.
    """
    operator = np.asarray(operator)
    state = np.asarray(state)
    if count_qubits_gate_matrix(operator) != 1:              # :355-359
        raise ValueError("expected operator on 1 qubit")
    num_qubits = count_qubits(state)
    out = state
    for q in range(num_qubits):
        out = apply_operator(operator, (q,), out)
    if out is state:
        out = state.copy()
    return out


def apply_phase(angles, state):
    """operations.py:15-42 -- psi_k <- exp(-i angle_k) psi_k with numpy/torch broadcasting."""
    angles = np.asarray(angles)
    state = np.asarray(state)
    # torch type promotion (operations.py:41): the phase factors are computed in the
    # ANGLES' precision (f32 angles -> complex64 factors, f64 -> complex128) and the
    # product is then promoted with the state's dtype.
    cdtype = np.complex128 if angles.dtype == np.float64 else np.complex64
    factors = np.exp(np.asarray(-1j, dtype=cdtype) * angles.astype(cdtype))
    out = factors * state
    want = np.result_type(cdtype, state.dtype)
    return out.astype(want, copy=False)


# --------------------------------------------------------------------------- #
# reductions  (src/unitair/states/innerprod.py)
# --------------------------------------------------------------------------- #
def abs_squared(state):
    """innerprod.py:4-26 -- (conj(s) * s).real."""
    state = np.asarray(state)
    return (np.conj(state) * state).real


def norm_squared(state):
    """innerprod.py:29-46."""
    return np.sum(abs_squared(state), axis=-1)


def diag_expectation_value(diag_values, state):
    """innerprod.py:49-59."""
    return np.sum(abs_squared(state) * np.asarray(diag_values), axis=-1)


def inner_product(state_1, state_2):
    """innerprod.py:62-65 -- left argument conjugated."""
    return np.sum(np.conj(np.asarray(state_1)) * np.asarray(state_2), axis=-1)


# --------------------------------------------------------------------------- #
# qubit permutations (operations.py:506-654) -- "next" rows f-2, used by the
# sharded path's tests.
# --------------------------------------------------------------------------- #
def permute_qubits(permutation, state_vector):
    """operations.py:600-623."""
    state_vector = np.asarray(state_vector)
    n = count_qubits(state_vector)
    st = permute_qubits_tensor(list(permutation), to_tensor_layout(state_vector), n)
    return to_vector_layout(st, n)


def swap(state, qubit_pair):
    """operations.py:506-537."""
    state = np.asarray(state)
    n = count_qubits(state)
    i, j = qubit_pair
    if i == j:
        return to_vector_layout(to_tensor_layout(state), n)
    for q in (i, j):
        if q >= n or q < -n:
            raise ValueError("qubit index out of range")
    perm = list(range(n))
    perm[i], perm[j] = perm[j], perm[i]
    return permute_qubits(perm, state)


def roll_qubits(state, num_steps=1):
    """operations.py:540-597: rolled[a_0..a_{n-1}] = psi[a_k..a_{n-1}, a_0..a_{k-1}] via the axis
    permutation identity[-steps:] + identity[:-steps] (:590-597)."""
    state = np.asarray(state)
    n = count_qubits(state)
    steps = num_steps % n
    if steps == 0:
        return to_vector_layout(to_tensor_layout(state), n)
    identity = list(range(n))
    return permute_qubits(identity[-steps:] + identity[:-steps], state)


def apply_to_qubits(operators, qubits, state):
    """operations.py:416-503 with fuse_single_qubit_operators (gates/matrix_algebra.py:6-51):
    operators on the same qubit are multiplied (later operator on the left), then each fused
    2x2 is applied to its qubit."""
    fused = {}
    for q, op in zip(qubits, operators):
        op = np.asarray(op)
        fused[q] = np.matmul(op, fused[q]) if q in fused else op
    out = np.asarray(state)
    for q, op in fused.items():
        out = apply_operator(op, [q], out)
    return out


def act_last_qubit(single_qubit_operator, state):
    """operations.py:189-233: einsum('ab, ...b -> ...a') on the last qubit axis."""
    state = np.asarray(state)
    n = count_qubits(state)
    return apply_operator(np.asarray(single_qubit_operator), [n - 1], state)


def multi_cz(qubit_pairs, state_vector):
    """operations.py:657-748 -- sign vector prod_pairs (1 - 2*[(i & crit) == crit]) times the state."""
    state_vector = np.asarray(state_vector)
    n = count_qubits(state_vector)
    pairs = np.asarray(qubit_pairs).reshape(-1, 2)
    if (pairs >= n).any():
        raise ValueError("Control/target indices for CZ gate must be less than num_bits.")
    if (pairs[:, 0] == pairs[:, 1]).any():
        raise ValueError("Control and target qubits are not distinct.")
    idx = np.arange(2 ** n)
    phases = np.ones(2 ** n, dtype=np.int64)
    for c, t in pairs:
        crit = 2 ** (n - c - 1) + 2 ** (n - t - 1)           # :729-731
        phases *= 1 - 2 * ((idx & crit) == crit)              # :745-746
    return (phases * state_vector).astype(state_vector.dtype)


def multi_controlled_z(qubits, state_vector):
    """operations.py:751-783 -- permute the listed qubits to the right, flip the sign of the
    all-ones component of that subsystem, permute back."""
    state_vector = np.asarray(state_vector)
    n = count_qubits(state_vector)
    qubits = list(qubits)
    inert = [q for q in range(n) if q not in set(qubits)]
    factors = np.ones(2 ** len(qubits))
    factors[-1] = -1.0
    factors = np.tile(factors, 2 ** len(inert))               # :770-772
    perm = inert + qubits
    reverse = [0] * n
    for q, i in enumerate(perm):
        reverse[i] = q
    out = permute_qubits(perm, state_vector)
    out = (factors * out).astype(state_vector.dtype)
    return permute_qubits(reverse, out)


def multi_controlled_x(state_vector, controls, target):
    """operations.py:786-806 -- H on the target, C...CZ on controls + target, H on the target."""
    state_vector = np.asarray(state_vector)
    h = (np.array([[1, 1], [1, -1]]) * 2 ** -0.5).astype(state_vector.dtype)
    out = apply_operator(h, [target], state_vector)
    out = multi_controlled_z(list(controls) + [target], out)
    return apply_operator(h, [target], out)


# --------------------------------------------------------------------------- #
# closed-form gradients (PyTorch's conjugate-Wirtinger
# convention).  The reference has no source for these (they come from torch's
# tape); they are pinned by tests/golden/grad_*.npz which hold torch-autograd
# results through the unmodified reference.
# --------------------------------------------------------------------------- #
def apply_operator_grads(operator, qubits, state, grad_out):
    """Return (grad_operator, grad_state) of L through out = apply_operator(U, q, psi)."""
    operator = np.asarray(operator)
    state = np.asarray(state)
    grad_out = np.asarray(grad_out)
    n = count_qubits(state)
    k = count_qubits_gate_matrix(operator)
    qubits = list(qubits)
    op_nb = count_gate_batch_dims(operator)
    op_batch = operator.shape[:op_nb]
    state_batch = state.shape[:-1]
    out_batch = grad_out.shape[:-1]

    u_h = np.conj(np.swapaxes(operator, -1, -2))
    g_in = apply_operator(u_h, qubits, grad_out)             # U^H g, batch = out_batch
    # sum over the batch dims the state was broadcast over
    extra = len(out_batch) - len(state_batch)
    if extra > 0:
        g_in = g_in.sum(axis=tuple(range(extra)))
    for ax, (a, b) in enumerate(zip(state_batch, g_in.shape[:-1])):
        if a == 1 and b != 1:
            g_in = g_in.sum(axis=ax, keepdims=True)
    grad_state = g_in

    # grad_U[a, b] = sum_r g[a, r] conj(psi[b, r]) in the `qubits` order
    perm = permutation_to_front(n, qubits)
    rest = 2 ** (n - k)

    def front(x):
        xt = permute_qubits_tensor(perm, to_tensor_layout(x), n, contiguous_output=True)
        return xt.reshape(x.shape[:-1] + (2 ** k, rest))

    g_f = front(grad_out)                                    # (*out_batch, 2^k, rest)
    psi_f = front(np.broadcast_to(state, out_batch + state.shape[-1:]))
    gu = np.matmul(g_f, np.conj(np.swapaxes(psi_f, -1, -2)))  # (*out_batch, 2^k, 2^k)
    extra = len(out_batch) - len(op_batch)
    if extra > 0:
        gu = gu.sum(axis=tuple(range(extra)))
    for ax, (a, b) in enumerate(zip(op_batch, gu.shape[:-2])):
        if a == 1 and b != 1:
            gu = gu.sum(axis=ax, keepdims=True)
    return gu, grad_state


def measurement_probabilities(state):
    """The distribution the reference samples from: Categorical(probs=abs_squared(state))
    (src/unitair/simulation/measurement.py:41-42; Categorical normalises probs by their sum)."""
    p = abs_squared(np.asarray(state)).astype(np.float64)
    return p / p.sum()


def sample_indices(state, uniforms):
    """Inverse-CDF sampling from measurement_probabilities(state): draw u in [0, 1) selects the
    first index whose cumulative probability exceeds u.  The reference draws from the same
    distribution with torch's multinomial sampler (measurement.py:42-43), whose random stream
    cannot be reproduced outside torch: parity for measure() is distributional."""
    state = np.asarray(state)
    p = (state.real.astype(np.float64) ** 2 + state.imag.astype(np.float64) ** 2)
    cdf = np.cumsum(p)
    t = np.asarray(uniforms, dtype=np.float64) * cdf[-1]
    return np.minimum(np.searchsorted(cdf, t, side="right"), len(p) - 1)
