"""Small gate library with the reference's names (src/unitair/gates/gates.py).

Out of the hot path (tiny tensors) but needed to write circuits without the reference
installed.  Constant gates take (device, dtype); parameterised gates take an angle tensor
(0-d or 1-d batch) and stay in ordinary torch autograd, so gradients flow from the state
through the native gate kernels into the angles.
"""
import math
from typing import Optional, Union

import torch


def _const(rows, device, dtype):
    dtype = torch.complex64 if dtype is None else dtype
    if not dtype.is_complex:
        raise TypeError(f'This gate requires a complex dtype, but it ended up with dtype {dtype}.')
    return torch.tensor(rows, dtype=dtype, device=device if device is not None else "cpu")


def hadamard(device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None):
    v = 1.0 / math.sqrt(2.0)
    return _const([[v, v], [v, -v]], device, dtype)


def pauli_x(device=None, dtype=None):
    return _const([[0, 1], [1, 0]], device, dtype)


def pauli_y(device=None, dtype=None):
    return _const([[0, -1j], [1j, 0]], device, dtype)


def pauli_z(device=None, dtype=None):
    return _const([[1, 0], [0, -1]], device, dtype)


def cnot(device=None, dtype=None):
    return _const([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], device, dtype)


cx = cnot


def cz(device=None, dtype=None):
    return _const([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]], device, dtype)


def _angle(angle, dtype):
    if not isinstance(angle, torch.Tensor):
        angle = torch.tensor(float(angle))
    if not dtype.is_complex:
        raise TypeError(f'This parameterized gate is required to be complex, got {dtype}.')
    real = torch.float64 if dtype == torch.complex128 else torch.float32
    return angle.to(real)


def _assemble(entries, dtype):
    """entries: 2x2 nested list of (re, im) tensor pairs with the angle's shape -> (*shape, 2, 2)."""
    rows = [torch.stack([torch.complex(re, im) for re, im in row], dim=-1) for row in entries]
    return torch.stack(rows, dim=-2).to(dtype)


def exp_x(angle: Union[torch.Tensor, float], dtype: torch.dtype = torch.complex64):
    """e^(-i angle X) = [[cos, -i sin], [-i sin, cos]]; angle () or (batch,)."""
    a = _angle(angle, dtype)
    c, s, z = torch.cos(a), torch.sin(a), torch.zeros_like(a)
    return _assemble([[(c, z), (z, -s)], [(z, -s), (c, z)]], dtype)


def exp_y(angle: Union[torch.Tensor, float], dtype: torch.dtype = torch.complex64):
    """e^(-i angle Y) = [[cos, -sin], [sin, cos]]."""
    a = _angle(angle, dtype)
    c, s, z = torch.cos(a), torch.sin(a), torch.zeros_like(a)
    return _assemble([[(c, z), (-s, z)], [(s, z), (c, z)]], dtype)


def exp_z(angle: Union[torch.Tensor, float], dtype: torch.dtype = torch.complex64):
    """e^(-i angle Z) = diag(cos - i sin, cos + i sin)."""
    a = _angle(angle, dtype)
    c, s, z = torch.cos(a), torch.sin(a), torch.zeros_like(a)
    return _assemble([[(c, -s), (z, z)], [(z, z), (c, s)]], dtype)
