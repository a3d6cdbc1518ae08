"""Shape arithmetic for states and gates: pure host integer math.

Mirrors src/unitair/states/shapes.py (count_qubits :13, count_qubits_gate_matrix :31,
count_batch_dims_tensor :63, subset_roll_to_back/front :121/:135, StateShapeError :149)
with the same error types, so callers and tests written against the reference keep
working.  Qubit q of an n-qubit state is bit (n-1-q) of the vector index.
"""
import enum
import math

import torch


class StateLayout(str, enum.Enum):
    VECTOR = "vector"
    TENSOR = "tensor"


class StateShapeError(ValueError):
    """Raised when a tensor's shape is not a valid state layout (a ValueError subclass)."""

    def __init__(self, data: torch.Tensor = None, expected_layout: StateLayout = None):
        self.data_size = tuple(data.size()) if data is not None else None
        self.layout = StateLayout(expected_layout.lower()) if expected_layout is not None else None

    def __str__(self):
        lines = ["There is a problem with the shape of a state."]
        if self.layout is StateLayout.VECTOR:
            lines.append("Expected VECTOR layout: size (*optional_batch_dims, 2^num_qubits).")
        elif self.layout is StateLayout.TENSOR:
            lines.append("Expected TENSOR layout: size (*optional_batch_dims, 2, 2, ..., 2), "
                         "one 2 per qubit.")
        if self.data_size is not None:
            lines.append(f"Given state has size {self.data_size}")
        return "\n".join(lines)


def _exact_log2(length: int):
    if length <= 0:
        return None
    bits = round(math.log2(length))
    return bits if 2 ** bits == length else None


def count_qubits(state: torch.Tensor) -> int:
    """Number of qubits of a state in vector layout (any batch dims)."""
    bits = _exact_log2(state.size()[-1])
    if bits is None:
        raise StateShapeError(data=state, expected_layout=StateLayout.VECTOR)
    return bits


def count_qubits_gate_matrix(gate: torch.Tensor) -> int:
    """Number of qubits a (*batch, 2^k, 2^k) gate acts on; RuntimeError if not 2^k."""
    bits = _exact_log2(gate.size()[-1])
    if bits is None:
        raise RuntimeError(f"Given gate matrix has size {gate.size()} which "
                           f"is not consistent with any number of qubits.")
    return bits


def hilbert_space_dim(state: torch.Tensor) -> int:
    return 2 ** count_qubits(state)


def count_qubits_tensor(state_tensor: torch.Tensor, num_batch_dims: int) -> int:
    return state_tensor.dim() - num_batch_dims


def count_batch_dims_tensor(state_tensor: torch.Tensor, num_qubits: int) -> int:
    return state_tensor.dim() - num_qubits


def get_qubit_indices(index, state_tensor: torch.Tensor, num_qubits: int):
    """Translate qubit indices to torch dims of a tensor-layout state (shapes.py:72-118).

    Done with plain integers (the reference builds a tensor per call just to range-check).
    """
    batch_dims = state_tensor.dim() - num_qubits

    def one(i):
        i = int(i)
        if i >= num_qubits or i < -num_qubits:
            raise ValueError("Expected index in {-num_qubits, ..., num_qubits - 1}.\n"
                             f"Num_qubits: {num_qubits}, index: {index}.")
        return i + batch_dims if i >= 0 else i

    if isinstance(index, torch.Tensor):
        if index.dim() == 0:
            return torch.tensor(one(index.item()))
        return torch.tensor([one(i) for i in index.tolist()])
    if isinstance(index, (list, tuple)):
        return [one(i) for i in index]
    return one(index)


def subset_roll_to_back(tensor: torch.Tensor, subset_num_dims: int) -> torch.Tensor:
    d = tensor.dim()
    return tensor.permute(list(range(subset_num_dims, d)) + list(range(subset_num_dims)))


def subset_roll_to_front(tensor: torch.Tensor, subset_num_dims: int) -> torch.Tensor:
    d = tensor.dim()
    return tensor.permute(list(range(d - subset_num_dims, d)) + list(range(d - subset_num_dims)))
