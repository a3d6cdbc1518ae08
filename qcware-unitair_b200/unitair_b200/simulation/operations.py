"""Gate application on states in vector layout -- the reference's public surface
(src/unitair/simulation/operations.py) backed by the B200 engine.

Signatures, keyword names, batch semantics, output shapes and error types follow the
reference (cited per function); the work is done by hand-written sm_100a kernels through
the C ABI in include/unitair_b200.h.  Inputs are never modified; outputs are new
contiguous tensors on the inputs' device and take part in autograd.
"""
import warnings
from typing import Iterable, Tuple

import torch

from .. import _engine
from .. import _lib
from .. import states
from ..utils import inverse_list_permutation


# --------------------------------------------------------------------------- #
def _complex_pair(operator: torch.Tensor, state: torch.Tensor):
    """dtype contract of the reference's bmm: operator and state must share a dtype.

    Real operators on real states work in the reference (result is real); they are
    promoted to complex here and the real part is returned by the caller.
    """
    if operator.dtype != state.dtype:
        raise RuntimeError(
            f"expected scalar type {state.dtype} but found {operator.dtype}: "
            "operator and state must have the same dtype")
    if state.dtype in (torch.complex64, torch.complex128):
        return operator, state, False
    if state.dtype == torch.float32:
        return operator.to(torch.complex64), state.to(torch.complex64), True
    if state.dtype == torch.float64:
        return operator.to(torch.complex128), state.to(torch.complex128), True
    raise RuntimeError(f"unitair_b200: unsupported state dtype {state.dtype}")


def apply_phase(angles: torch.Tensor, state: torch.Tensor):
    """Multiply the kth component of state by e^(-i angles_k)  (operations.py:15-42).

    Batching follows PyTorch multiplication broadcasting between `angles` and `state`,
    and so does dtype promotion: f32 angles with a complex64 state give complex64, f64
    angles or a complex128 state give complex128, a real state is promoted to complex.
    One fused kernel instead of exp + mul + mul.
    """
    if not isinstance(angles, torch.Tensor):
        angles = torch.as_tensor(angles, device=state.device)
    _lib.require_cuda(angles, state)
    if angles.is_complex():
        # exp(-i z) for complex z is not a phase; same formula as the reference, stock torch
        return torch.exp(-1.j * angles) * state
    if angles.dtype not in (torch.float32, torch.float64):
        angles = angles.to(torch.get_default_dtype())
    if not state.is_complex():
        state = state.to(torch.complex128 if state.dtype == torch.float64 else torch.complex64)
    if angles.dtype == torch.float64 or state.dtype == torch.complex128:
        out_dtype = torch.complex128
    else:
        out_dtype = torch.complex64
    if out_dtype == torch.complex128:
        if angles.dtype == torch.float32:
            # the reference computes the factors in f32 and promotes the product
            angles_n = angles.to(torch.float64)
        else:
            angles_n = angles
        state_n = state.to(torch.complex128)
    else:
        angles_n, state_n = angles, state
    return _engine.apply_phase_native(angles_n, state_n)


def apply_operator(operator: torch.Tensor, qubits: Iterable[int], state: torch.Tensor):
    """Apply a dense k-qubit operator to the ordered `qubits` of a state in vector layout.

    Same contract as the reference (operations.py:45-148):
      * `operator` has size (*operator_batch, 2^k, 2^k), `state` (*state_batch, 2^n);
        the gate's most significant index bit acts on qubits[0], so the order matters.
      * batch structures: identical batch dims; unbatched operator on a batch of states;
        batched operator on one state (all operators act on the same state in parallel).
      * ValueError for qubits outside range(n), for len(qubits) != k and for repeated
        qubits; StateShapeError if the state's last dim is not 2^n; RuntimeError if the
        operator's last dim is not 2^k or the dtypes differ.
    One pass over the state (no permute copies): 16 B / 32 B of HBM traffic per amplitude.
    """
    num_qubits = states.count_qubits(state)
    qubits = [int(q) for q in qubits]
    if not set(qubits).issubset(range(num_qubits)):
        raise ValueError(
            f'qubits={qubits} is not consistent with state vector with '
            f'{num_qubits} qubits.')
    op_num_qubits = states.count_qubits_gate_matrix(operator)
    if len(qubits) != op_num_qubits:
        raise ValueError(
            f'Cannot apply operator with {op_num_qubits} to the {len(qubits)} '
            f'qubit sequence {qubits}.')
    if len(set(qubits)) != len(qubits):
        dup = next(q for i, q in enumerate(qubits) if q in qubits[:i])
        raise ValueError(f'{dup} is not in list')   # the reference's list.index failure
    if operator.dim() < 2 or operator.size(-2) != operator.size(-1):
        raise RuntimeError(f'operator with size {tuple(operator.size())} is not a batch of '
                           f'square matrices.')
    _lib.require_cuda(operator, state)
    op_c, st_c, was_real = _complex_pair(operator, state)
    out = _engine.apply_gate(op_c, qubits, st_c, num_qubits, op_num_qubits)
    return out.real.contiguous() if was_real else out


def apply_operator_tensor(operator, qubits, state_tensor, num_qubits, operator_num_qubits=None):
    """Tensor-layout variant (operations.py:151-186)."""
    vec = states.to_vector_layout(state_tensor, num_qubits)
    out = apply_operator(operator, qubits, vec)
    return states.to_tensor_layout(out)


def act_first_qubits(operator: torch.Tensor, state: torch.Tensor):
    """Apply a multi-qubit gate to the first qubits of a state (operations.py:236-255)."""
    num_qubits = states.count_qubits(state)
    gate_num_qubits = states.count_qubits_gate_matrix(operator)
    if num_qubits < gate_num_qubits:
        raise ValueError(
            f'Attempted to apply a {gate_num_qubits}-qubit gate to {num_qubits} qubit(s).')
    return apply_operator(operator, range(gate_num_qubits), state)


def act_first_qubits_tensor(operator: torch.Tensor, state_tensor: torch.Tensor, num_qubits: int,
                            gate_num_qubits=None):
    """Tensor-layout variant of act_first_qubits (operations.py:258-329): the operator acts on
    the first `gate_num_qubits` qubit axes; same three batch structures as apply_operator."""
    if gate_num_qubits is None:
        gate_num_qubits = states.count_qubits_gate_matrix(operator)
    vec = states.to_vector_layout(state_tensor, num_qubits)
    out = apply_operator(operator, range(gate_num_qubits), vec)
    return states.to_tensor_layout(out)


def act_last_qubit(single_qubit_operator: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """Apply a 2x2 operator to the last qubit (operations.py:189-213; deprecated there too)."""
    warnings.warn(
        'act_last_qubit and act_last_qubit_tensor are outdated. These\n'
        'functions will be removed from Unitair in a later release or may\n'
        'be revised to meet our current standards.\n'
        'Please consider using apply_operator instead.')
    num_qubits = states.count_qubits(state)
    return apply_operator(single_qubit_operator, [num_qubits - 1], state)


def act_last_qubit_tensor(single_qubit_operator: torch.Tensor, state_tensor: torch.Tensor) -> torch.Tensor:
    """Tensor-layout variant of act_last_qubit (operations.py:216-233): contracts the matrix with
    the last axis, whatever the other axes are."""
    flat = state_tensor.reshape(-1, 2)
    if flat.shape[0] == 0:
        return state_tensor.clone()
    n = (flat.shape[0].bit_length() - 1) + 1
    if 1 << (n - 1) != flat.shape[0]:
        # not a power of two overall: treat the leading axes as a batch of 1-qubit states
        out = apply_operator(single_qubit_operator, [0], flat)
    else:
        out = apply_operator(single_qubit_operator, [n - 1], flat.reshape(-1))
    return out.reshape(state_tensor.shape)


def apply_all_qubits(operator: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """Apply the same single-qubit operator to every qubit (operations.py:332-413).

    `operator` has size (*batch_dims, 2, 2): shared by the whole batch of states, or one
    operator per batch entry.  The reference walks the qubits in a Python loop (n passes
    over memory); here the 2x2 is applied to all qubits of a shared-memory tile at once,
    so a 30-qubit state needs 4 passes instead of 30.
    """
    if states.count_qubits_gate_matrix(operator) != 1:
        raise ValueError(
            f'Expected operator on 1 qubit, found a '
            f'{states.count_qubits_gate_matrix(operator)} qubit operator.')
    num_qubits = states.count_qubits(state)
    _lib.require_cuda(operator, state)
    op_c, st_c, was_real = _complex_pair(operator, state)
    from .. import circuit
    out = circuit.apply_same_gate_all_qubits(op_c, st_c, num_qubits)
    return out.real.contiguous() if was_real else out


def apply_all_qubits_tensor(operator, state_tensor, num_qubits):
    vec = states.to_vector_layout(state_tensor, num_qubits)
    return states.to_tensor_layout(apply_all_qubits(operator, vec))


# --------------------------------------------------------------------------- #
# "next" rows of SURVEY.md 8(f): single-qubit lists and qubit permutations
# --------------------------------------------------------------------------- #
def apply_to_qubits(operators: Iterable[torch.Tensor], qubits: Iterable[int], state: torch.Tensor):
    """Apply single-qubit gates to the listed qubits (operations.py:416-503).

    Gates on the same qubit are multiplied first (the reference's fusion,
    src/unitair/gates/matrix_algebra.py:6-51); the fused 2x2s then go through the
    shared-memory pass, several qubits per pass.
    """
    from .. import circuit
    num_qubits = states.count_qubits(state)
    fused = {}
    for q, op in zip(qubits, operators):
        q = int(q)
        if q < 0 or q >= num_qubits:
            raise ValueError(f'qubit {q} is not consistent with {num_qubits} qubits.')
        fused[q] = torch.matmul(op, fused[q]) if q in fused else op
    if not fused:
        return states.to_vector_layout(states.to_tensor_layout(state), num_qubits).clone()
    ops = list(fused.values())
    _lib.require_cuda(state, *ops)
    gates = [([q], op) for q, op in fused.items()]
    return circuit.apply_gates(gates, state)


def apply_to_qubits_tensor(operators: Iterable[torch.Tensor], qubits: Iterable[int],
                           state_tensor: torch.Tensor, num_qubits: int):
    """Tensor-layout variant of apply_to_qubits (operations.py:449-503)."""
    vec = states.to_vector_layout(state_tensor, num_qubits)
    return states.to_tensor_layout(apply_to_qubits(operators, qubits, vec))


def permute_qubits_tensor(permutation: Iterable[int], state_tensor: torch.Tensor, num_qubits: int,
                          contiguous_output: bool = False):
    """Tensor-layout variant of permute_qubits (operations.py:626-654).  The reference returns a
    strided view unless contiguous_output=True; here the permutation is one native pass and the
    result is always contiguous (same values, same shape)."""
    vec = states.to_vector_layout(state_tensor, num_qubits)
    return states.to_tensor_layout(permute_qubits(permutation, vec))


def swap_tensor(state_tensor: torch.Tensor, qubit_pair: Tuple[int, int], num_qubits: int):
    """Tensor-layout variant of swap (operations.py:520-537)."""
    if qubit_pair[0] == qubit_pair[1]:
        return state_tensor
    vec = states.to_vector_layout(state_tensor, num_qubits)
    return states.to_tensor_layout(swap(vec, qubit_pair))


def roll_qubits_tensor(state_tensor: torch.Tensor, num_qubits: int, num_steps: int = 1):
    """Tensor-layout variant of roll_qubits (operations.py:565-597):
    rolled[a_0, ..., a_{n-1}] = psi[a_k, ..., a_{n-1}, a_0, ..., a_{k-1}]... with k = n - num_steps."""
    vec = states.to_vector_layout(state_tensor, num_qubits)
    return states.to_tensor_layout(roll_qubits(vec, num_steps))


def permute_qubits(permutation: Iterable[int], state_vector: torch.Tensor):
    """Permute qubits of a state in vector layout (operations.py:600-654).

    Follows torch.permute semantics on the tensor layout: output qubit axis i is input
    qubit axis permutation[i].  Pure data movement, bit-exact, one pass.
    """
    num_qubits = states.count_qubits(state_vector)
    permutation = [int(p) for p in permutation]
    if sorted(permutation) != list(range(num_qubits)):
        raise RuntimeError(f'permutation {permutation} is not a permutation of '
                           f'{num_qubits} qubits')
    _lib.require_cuda(state_vector)
    from .. import circuit
    return circuit.permute_qubits_native(permutation, state_vector, num_qubits)


def swap(state: torch.Tensor, qubit_pair: Tuple[int, int]):
    """Swap a pair of qubits (operations.py:506-537)."""
    num_qubits = states.count_qubits(state)
    i, j = int(qubit_pair[0]), int(qubit_pair[1])
    for q in (i, j):
        if q >= num_qubits or q < -num_qubits:
            raise ValueError('Expected index in {-num_qubits, ..., num_qubits - 1}.\n'
                             f'Num_qubits: {num_qubits}, index: {qubit_pair}.')
    i %= num_qubits
    j %= num_qubits
    perm = list(range(num_qubits))
    perm[i], perm[j] = perm[j], perm[i]
    return permute_qubits(perm, state)


def roll_qubits(state: torch.Tensor, num_steps=1):
    """Cyclic permutation of qubits (operations.py:540-597)."""
    num_qubits = states.count_qubits(state)
    steps = num_steps % num_qubits if num_qubits else 0
    identity = list(range(num_qubits))
    perm = identity[-steps:] + identity[:-steps] if steps else identity
    return permute_qubits(perm, state)


# --------------------------------------------------------------------------- #
# sign-mask diagonal gates: SURVEY.md 8(f-3)
# --------------------------------------------------------------------------- #
class _SignMasks(torch.autograd.Function):
    """out = D psi with D = diag(+-1) from bit masks; D is real and its own adjoint."""

    @staticmethod
    def forward(ctx, state, n, masks):
        ctx.n, ctx.masks = n, masks
        return _sign_masks_launch(state, n, masks)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        return _sign_masks_launch(_engine._aligned(grad), ctx.n, ctx.masks), None, None


def _sign_masks_launch(state, n, masks):
    import ctypes
    dev = state.device
    out = torch.empty_like(state)
    batch = state.numel() >> n
    if batch == 0:
        return out
    src = state
    lib = _lib.lib()
    with _lib.on_device(dev):
        for i in range(0, max(len(masks), 1), 64):        # 64 masks per launch
            chunk = masks[i:i + 64]
            arr = (ctypes.c_ulonglong * max(len(chunk), 1))(*chunk)
            _lib.check(lib.ua_apply_sign_masks(_lib.dtype_code(state.dtype), out.data_ptr(), src.data_ptr(),
                                               n, batch, len(chunk), arr, _lib.stream_ptr(dev)))
            src = out
    return out


def _apply_sign_masks(state_vector, num_qubits, masks):
    _lib.require_cuda(state_vector)
    was_real = not state_vector.is_complex()
    st = state_vector
    if was_real:
        st = st.to(torch.complex128 if st.dtype == torch.float64 else torch.complex64)
    st = _engine._aligned(st)
    if torch.is_grad_enabled() and st.requires_grad:
        out = _SignMasks.apply(st, num_qubits, list(masks))
    else:
        out = _sign_masks_launch(st, num_qubits, list(masks))
    return out.real.contiguous() if was_real else out


def multi_cz(qubit_pairs, state_vector: torch.Tensor, num_bits_memory_cutoff=None):
    """Apply CZ gates to the given qubit pairs (operations.py:657-748).

    `qubit_pairs` has size (2,) or (num_pairs, 2) (tensor or nested list).  The reference
    builds a 2^n sign vector with arange/bitwise_and/prod (and needs a memory cutoff for it);
    here the pair masks are tested inside one streaming kernel, so `num_bits_memory_cutoff`
    is accepted and ignored.
    """
    num_qubits = states.count_qubits(state_vector)
    if isinstance(qubit_pairs, torch.Tensor):
        if qubit_pairs.dim() == 1:
            qubit_pairs = qubit_pairs.view(1, 2)
        elif qubit_pairs.dim() != 2:
            raise ValueError('qubit_pairs must have size (2,) or (num_pairs, 2).')
        pairs = qubit_pairs.tolist()
    else:
        pairs = list(qubit_pairs)
        if pairs and not isinstance(pairs[0], (list, tuple)):
            pairs = [pairs]
    masks = []
    for pair in pairs:
        if len(pair) != 2:
            raise ValueError('qubit_pairs must have size (2,) or (num_pairs, 2).')
        c, t = int(pair[0]), int(pair[1])
        if c >= num_qubits or t >= num_qubits or c < 0 or t < 0:
            raise ValueError('Control/target indices for CZ gate must be less than num_bits.')
        if c == t:
            raise ValueError('Control and target qubits are not distinct.')
        masks.append((1 << (num_qubits - 1 - c)) | (1 << (num_qubits - 1 - t)))
    return _apply_sign_masks(state_vector, num_qubits, masks)


def multi_controlled_z(qubits: Iterable[int], state_vector: torch.Tensor):
    """Apply a CC...CZ gate to the given qubits (operations.py:751-783): the amplitude whose
    listed bits are all 1 changes sign.  One pass instead of two full-state permutations."""
    num_qubits = states.count_qubits(state_vector)
    qubits = [int(q) for q in qubits]
    if not set(qubits).issubset(range(num_qubits)):
        raise ValueError(f'qubits={qubits} is not consistent with {num_qubits} qubits.')
    mask = 0
    for q in set(qubits):
        mask |= 1 << (num_qubits - 1 - q)
    if mask == 0:
        return -state_vector if state_vector.numel() else state_vector.clone()
    return _apply_sign_masks(state_vector, num_qubits, [mask])


def multi_controlled_x(state_vector: torch.Tensor, controls: Iterable[int], target: int):
    """C...CX as H . C...CZ . H on the target (operations.py:786-806)."""
    controls = [int(c) for c in controls]
    dtype = state_vector.dtype if state_vector.is_complex() else (
        torch.complex128 if state_vector.dtype == torch.float64 else torch.complex64)
    from .. import gates as _gates
    h = _gates.hadamard(device=state_vector.device, dtype=dtype)
    sv = state_vector if state_vector.is_complex() else state_vector.to(dtype)
    sv = apply_operator(operator=h, qubits=[target], state=sv)
    sv = multi_controlled_z(qubits=controls + [int(target)], state_vector=sv)
    return apply_operator(operator=h, qubits=[target], state=sv)
