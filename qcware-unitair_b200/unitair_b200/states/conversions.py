"""vector layout (*B, 2^n)  <->  tensor layout (*B, 2, ..., 2).

Mirrors src/unitair/states/conversions.py:5-92.  These are views (plus a contiguous
copy when the input is strided); they define the memory order the kernels honour:
the last axis of the vector layout is contiguous and qubit 0 is its most significant bit.
"""
import torch

from . import shapes


def to_tensor_layout(state: torch.Tensor) -> torch.Tensor:
    num_qubits = shapes.count_qubits(state)
    if not state.is_contiguous():
        state = state.contiguous()
    return state.view(state.size()[:-1] + torch.Size([2] * num_qubits))


def to_vector_layout(state_tensor: torch.Tensor, num_qubits: int) -> torch.Tensor:
    if not state_tensor.is_contiguous():
        state_tensor = state_tensor.contiguous()
    batch = state_tensor.size()[:state_tensor.dim() - num_qubits]
    return state_tensor.view(batch + torch.Size((2 ** num_qubits,)))
