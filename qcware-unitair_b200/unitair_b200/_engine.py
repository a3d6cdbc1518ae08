"""Thin launch layer between the reference-shaped Python fronts and the C ABI.

Every function here takes already-validated, contiguous CUDA tensors, allocates the
output with torch's caching allocator and enqueues one native call on the current
stream.  The autograd Functions implement the closed-form backward of SURVEY.md 3.4
with native kernels (adjoint apply, gate-gradient reduction, fused phase backward).
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib as L


# --------------------------------------------------------------------------- #
# raw launches
# --------------------------------------------------------------------------- #
def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (views into odd offsets are copied)."""
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def launch_gate(out, state, gate, n: int, k: int, qubits: Sequence[int], batch: int,
                in_stride: int, gate_stride: int, adjoint: bool):
    dev = state.device
    with L.on_device(dev):
        L.check(L.lib().ua_apply_gate(
            L.dtype_code(state.dtype), out.data_ptr(), state.data_ptr(), gate.data_ptr(),
            n, k, L.int_array(qubits), batch, in_stride, gate_stride, 1 if adjoint else 0,
            L.stream_ptr(dev)))
    return out


def launch_gate_grad(grad_out, psi, n, k, qubits, batch, psi_stride, gate_stride, gate_shape):
    dev = grad_out.device
    code = L.dtype_code(grad_out.dtype)
    grad_gate = torch.empty(gate_shape, dtype=grad_out.dtype, device=dev)
    with L.on_device(dev):
        nbytes = L.lib().ua_gate_grad_workspace_bytes(code, n, k, batch, gate_stride)
        ws, ws_ptr = L.workspace(nbytes, dev)
        L.check(L.lib().ua_gate_grad(
            code, grad_gate.data_ptr(), grad_out.data_ptr(), psi.data_ptr(), n, k,
            L.int_array(qubits), batch, psi_stride, gate_stride, ws_ptr, nbytes,
            L.stream_ptr(dev)))
    return grad_gate


# --------------------------------------------------------------------------- #
# dense gate with autograd
# --------------------------------------------------------------------------- #
class _ApplyGate(torch.autograd.Function):
    """out = U . psi on `qubits`;  backward: grad_psi = U^H g,  grad_U = g psi^H."""

    @staticmethod
    def forward(ctx, gate, state, qubits, n, k, batch, in_stride, gate_stride, out_shape):
        out = torch.empty(out_shape, dtype=state.dtype, device=state.device)
        launch_gate(out, state, gate, n, k, qubits, batch, in_stride, gate_stride, False)
        ctx.meta = (tuple(qubits), n, k, batch, in_stride, gate_stride)
        # like torch's tape, keep the input state only when the gate needs a gradient
        ctx.save_for_backward(gate, state if ctx.needs_input_grad[0] else None)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        gate, state = ctx.saved_tensors
        qubits, n, k, batch, in_stride, gate_stride = ctx.meta
        grad_out = _aligned(grad_out)
        grad_gate = grad_state = None
        if ctx.needs_input_grad[1]:
            g_in = torch.empty_like(grad_out)
            launch_gate(g_in, grad_out, gate, n, k, qubits, batch, 1 << n, gate_stride, True)
            if in_stride == 0 and batch > 1:
                # the state was broadcast over the gate batch: sum over that batch
                g_in = g_in.reshape(batch, 1 << n).sum(dim=0)
            grad_state = g_in
        if ctx.needs_input_grad[0]:
            grad_gate = launch_gate_grad(grad_out, state, n, k, qubits, batch, in_stride,
                                         gate_stride, gate.shape)
        return grad_gate, grad_state, None, None, None, None, None, None, None


def apply_gate(gate: torch.Tensor, qubits: Sequence[int], state: torch.Tensor, n: int, k: int):
    """Dense gate on a validated (operator, qubits, state) triple of complex CUDA tensors.

    Maps the reference's batch structures (src/unitair/simulation/operations.py:88-112,
    277-309) to (batch, in_stride, gate_stride):
      same batch dims        -> in_stride 2^n, gate_stride 4^k
      shared gate            -> in_stride 2^n, gate_stride 0
      batched gate, 1 state  -> in_stride 0,   gate_stride 4^k   (state is not expanded)
    Right-aligned broadcasting (operator batch shorter than the state batch) is
    materialised with expand().contiguous() first.
    """
    dim = 1 << n
    gdim = 1 << k
    op_batch = tuple(gate.shape[:-2])
    st_batch = tuple(state.shape[:-1])
    if not op_batch:
        out_batch = st_batch
        in_stride, gate_stride = dim, 0
    elif not st_batch:
        out_batch = op_batch
        in_stride, gate_stride = 0, gdim * gdim
    else:
        if len(op_batch) > len(st_batch):
            raise RuntimeError(
                f"operator batch dims {op_batch} cannot be broadcast to state batch dims {st_batch}")
        try:
            out_batch = tuple(torch.broadcast_shapes(op_batch, st_batch))
        except RuntimeError as e:
            raise RuntimeError(
                f"operator batch dims {op_batch} and state batch dims {st_batch} are not "
                f"broadcastable: {e}") from None
        if out_batch != op_batch:
            gate = gate.expand(out_batch + (gdim, gdim))
        if out_batch != st_batch:
            state = state.expand(out_batch + (dim,))
        in_stride, gate_stride = dim, gdim * gdim
    batch = 1
    for s in out_batch:
        batch *= s
    out_shape = out_batch + (dim,)
    if batch == 0:
        return torch.empty(out_shape, dtype=state.dtype, device=state.device)
    gate = _aligned(gate)
    state = _aligned(state)
    qubits = list(qubits)
    if torch.is_grad_enabled() and (gate.requires_grad or state.requires_grad):
        return _ApplyGate.apply(gate, state, qubits, n, k, batch, in_stride, gate_stride, out_shape)
    out = torch.empty(out_shape, dtype=state.dtype, device=state.device)
    return launch_gate(out, state, gate, n, k, qubits, batch, in_stride, gate_stride, False)


# --------------------------------------------------------------------------- #
# diagonal phase with autograd
# --------------------------------------------------------------------------- #
def _prod(shape):
    r = 1
    for s in shape:
        r *= int(s)
    return r


def _classify(shape, out_batch, elems):
    """(batch_stride, elem_stride) of an operand broadcast to out_batch + (elems,), or None."""
    shape = tuple(shape)
    if _prod(shape) == 1:
        return (0, 0)                                   # one value for everything
    last, lead = shape[-1], shape[:-1]
    if last == elems and _prod(lead) == 1:
        return (0, 1)                                   # one row shared by every batch entry
    if (1,) * (len(out_batch) - len(lead)) + lead == tuple(out_batch):
        if last == elems:
            return (elems, 1)                           # full size
        if last == 1:
            return (1, 0)                               # one value per batch entry
    return None


def _launch_phase(angles, state, plan, conj):
    batch, elems, a_bs, a_es, s_bs, out_shape = plan
    dev = state.device
    out = torch.empty(out_shape, dtype=state.dtype, device=dev)
    with L.on_device(dev):
        L.check(L.lib().ua_apply_phase(
            L.dtype_code(state.dtype), out.data_ptr(), state.data_ptr(), angles.data_ptr(),
            elems, batch, s_bs, a_bs, a_es, 1 if conj else 0, L.stream_ptr(dev)))
    return out


def _reduce_to(full, strides, batch, elems, shape):
    """Sum a (batch, elems) gradient over the axes an operand was broadcast along."""
    bs, es = strides
    g = full.reshape(batch, elems)
    if bs == 0:
        g = g.sum(dim=0, keepdim=True)
    if es == 0:
        g = g.sum(dim=1, keepdim=True)
    return g.reshape(shape)


class _ApplyPhase(torch.autograd.Function):
    """out = exp(-i angles) * state;  backward: grad_state = exp(+i a) g,
    grad_angles = sum_bcast Im(conj(g) out)   (SURVEY.md 3.4)."""

    @staticmethod
    def forward(ctx, angles, state, plan):
        out = _launch_phase(angles, state, plan, False)
        ctx.plan = plan
        ctx.shapes = (tuple(angles.shape), tuple(state.shape))
        ctx.save_for_backward(angles, state if ctx.needs_input_grad[0] else None)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        angles, state = ctx.saved_tensors
        batch, elems, a_bs, a_es, s_bs, out_shape = ctx.plan
        angle_shape, state_shape = ctx.shapes
        grad_out = _aligned(grad_out)
        dev = grad_out.device
        need_a, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_state = torch.empty_like(grad_out) if need_s else None
        g_angle = torch.empty(out_shape, dtype=angles.dtype, device=dev) if need_a else None
        with L.on_device(dev):
            L.check(L.lib().ua_phase_backward(
                L.dtype_code(grad_out.dtype),
                g_state.data_ptr() if need_s else None,
                g_angle.data_ptr() if need_a else None,
                grad_out.data_ptr(), state.data_ptr() if need_a else None, angles.data_ptr(),
                elems, batch, s_bs, a_bs, a_es, L.stream_ptr(dev)))
        if need_a:
            g_angle = _reduce_to(g_angle, (a_bs, a_es), batch, elems, angle_shape)
        if need_s:
            g_state = _reduce_to(g_state, (s_bs, 1), batch, elems, state_shape)
        return g_angle, g_state, None


def apply_phase_native(angles: torch.Tensor, state: torch.Tensor):
    """exp(-i angles) * state for real `angles` (f32 with c64, f64 with c128) and complex
    `state` on the same CUDA device, with torch broadcasting between the two."""
    out_shape = tuple(torch.broadcast_shapes(tuple(angles.shape), tuple(state.shape)))
    if _prod(out_shape) == 0:
        return torch.empty(out_shape, dtype=state.dtype, device=state.device)
    if len(out_shape) == 0:
        return apply_phase_native(angles.reshape(1), state.reshape(1)).reshape(())
    elems = out_shape[-1]
    out_batch = out_shape[:-1]
    batch = _prod(out_batch)
    a_str = _classify(angles.shape, out_batch, elems)
    if a_str is None:
        angles = angles.expand(out_shape)
        a_str = (elems, 1)
    s_str = _classify(state.shape, out_batch, elems)
    if s_str is None or s_str[1] == 0:
        if elems == 1 and s_str is not None:
            # a column of single-amplitude "states": (bs, 0) is the same memory as (bs, 1)
            s_str = (elems if s_str[0] == 1 else 0, 1)
        else:
            state = state.expand(out_shape)
            s_str = (elems, 1)
    if not angles.is_contiguous():
        angles = angles.contiguous()
    state = _aligned(state)
    plan = (batch, elems, a_str[0], a_str[1], s_str[0], out_shape)
    if torch.is_grad_enabled() and (angles.requires_grad or state.requires_grad):
        return _ApplyPhase.apply(angles, state, plan)
    return _launch_phase(angles, state, plan, False)
