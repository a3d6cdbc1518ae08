"""State constructors with the reference's names (src/unitair/initializations.py): one-shot
helpers used to build inputs; not part of the hot path."""
from typing import Optional, Sequence

import torch


def unit_vector(index: int, num_qubits: Optional[int] = None, dim: Optional[int] = None,
                device=torch.device("cpu"), dtype: torch.dtype = torch.complex64):
    if (dim is None) == (num_qubits is None):
        raise TypeError('Specify a unit vector by exactly one of `num_qubits` and `dim`.')
    if dim is None:
        dim = 2 ** num_qubits
    vec = torch.zeros(dim, device=device, dtype=dtype)
    vec[index] = 1
    return vec


def rand_state(num_qubits: int, batch_dims: Optional[Sequence] = None,
               device=torch.device("cpu"), dtype: torch.dtype = torch.complex64,
               requires_grad: bool = False, generator: Optional[torch.Generator] = None):
    """Normalised random state(s): normal entries scaled to unit L2 norm."""
    size = tuple(batch_dims or ()) + (2 ** num_qubits,)
    state = torch.randn(size, device=device, dtype=dtype, generator=generator)
    norm = (state.real ** 2 + (state.imag ** 2 if state.is_complex() else 0)).sum(-1, keepdim=True).sqrt()
    state = state / norm
    return state.requires_grad_(requires_grad)


def uniform_superposition(num_qubits: int, batch_dims: Optional[Sequence] = None,
                          device=torch.device("cpu"), dtype: torch.dtype = torch.complex64,
                          requires_grad: bool = False):
    size = tuple(batch_dims or ()) + (2 ** num_qubits,)
    state = torch.full(size, 2 ** (-num_qubits / 2.0), device=device, dtype=dtype)
    return state.requires_grad_(requires_grad)
