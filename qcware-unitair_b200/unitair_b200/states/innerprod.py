"""Reductions over states (mirror of src/unitair/states/innerprod.py:4-65), each a single
fused read of the state on the GPU instead of 2-4 elementwise/reduce passes."""
import torch

from .. import _lib as L


def _complexify(state):
    if state.is_complex():
        return state
    if state.dtype == torch.float64:
        return state.to(torch.complex128)
    return state.to(torch.complex64)


def _real_dtype(cdtype):
    return torch.float64 if cdtype == torch.complex128 else torch.float32


def _rows(state):
    elems = state.shape[-1]
    batch = 1
    for s in state.shape[:-1]:
        batch *= s
    return batch, elems


def _contig(t):
    return t if t.is_contiguous() else t.contiguous()


def _real_scale(state, g, diag, batch, elems, s_bs, g_bs, g_es, d_bs, scale=2.0):
    """scale * g * diag * state in one native pass (ua_real_scale), or None when the layout does
    not fit it (odd complex64 rows, misaligned views): the caller then uses the eager formula."""
    if state.numel() == 0 or (state.dtype == torch.complex64 and elems % 2):
        return None
    g = _contig(g.to(_real_dtype(state.dtype)))
    out = torch.empty((batch, elems), dtype=state.dtype, device=state.device)
    if (state.data_ptr() | out.data_ptr()) & 15:
        return None
    dev = state.device
    with L.on_device(dev):
        rc = L.lib().ua_real_scale(L.dtype_code(state.dtype), out.data_ptr(), state.data_ptr(), g.data_ptr(),
                                   diag.data_ptr() if diag is not None else None, elems, batch, s_bs,
                                   g_bs, g_es, d_bs, float(scale), L.stream_ptr(dev))
    if rc == L.UA_ERR_UNSUPPORTED:
        return None
    L.check(rc)
    return out


class _AbsSquared(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state):
        dev = state.device
        out = torch.empty(state.shape, dtype=_real_dtype(state.dtype), device=dev)
        if state.numel():
            with L.on_device(dev):
                L.check(L.lib().ua_abs_squared(L.dtype_code(state.dtype), out.data_ptr(),
                                               state.data_ptr(), state.numel(), L.stream_ptr(dev)))
        ctx.save_for_backward(state)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        state, = ctx.saved_tensors
        # d|z|^2 -> 2 g z (conjugate Wirtinger convention), one fused pass
        n = state.numel()
        out = _real_scale(state, _contig(grad), None, 1, n, n, n, 1, 0)
        return out.reshape(state.shape) if out is not None else 2 * grad * state


def abs_squared(state: torch.Tensor):
    """Vector of measurement probabilities (|x_1|^2, ..., |x_N|^2)  (innerprod.py:4-26)."""
    L.require_cuda(state)
    st = _contig(_complexify(state))
    if torch.is_grad_enabled() and st.requires_grad:
        return _AbsSquared.apply(st)
    return _AbsSquared.forward(_Ctx(), st)


class _Ctx:
    needs_input_grad = (False, False)

    def save_for_backward(self, *a):
        pass


class _NormSquared(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state):
        dev = state.device
        batch, elems = _rows(state)
        out = torch.empty(state.shape[:-1], dtype=_real_dtype(state.dtype), device=dev)
        if batch:
            with L.on_device(dev):
                nbytes = L.lib().ua_reduce_workspace_bytes(batch, elems)
                ws, ws_ptr = L.workspace(nbytes, dev)
                L.check(L.lib().ua_norm_squared(L.dtype_code(state.dtype), out.data_ptr(),
                                                state.data_ptr(), elems, batch, ws_ptr, nbytes,
                                                L.stream_ptr(dev)))
        ctx.save_for_backward(state)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        state, = ctx.saved_tensors
        batch, elems = _rows(state)
        out = _real_scale(state, grad.reshape(-1), None, batch, elems, elems, 1, 0, 0) if batch else None
        return out.reshape(state.shape) if out is not None else 2 * grad.unsqueeze(-1) * state


def norm_squared(state: torch.Tensor):
    """L^2 norm squared <state|state> per batch entry (innerprod.py:29-46)."""
    L.require_cuda(state)
    st = _contig(_complexify(state))
    if torch.is_grad_enabled() and st.requires_grad:
        return _NormSquared.apply(st)
    return _NormSquared.forward(_Ctx(), st)


class _DiagExpectation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, diag, state, batch, elems, d_bs, s_bs, out_shape):
        dev = state.device
        out = torch.empty(out_shape, dtype=_real_dtype(state.dtype), device=dev)
        with L.on_device(dev):
            nbytes = L.lib().ua_reduce_workspace_bytes(batch, elems)
            ws, ws_ptr = L.workspace(nbytes, dev)
            L.check(L.lib().ua_diag_expectation(
                L.dtype_code(state.dtype), out.data_ptr(), diag.data_ptr(), state.data_ptr(),
                elems, batch, d_bs, s_bs, ws_ptr, nbytes, L.stream_ptr(dev)))
        ctx.save_for_backward(diag, state)
        ctx.meta = (batch, elems, d_bs, s_bs)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        diag, state = ctx.saved_tensors
        batch, elems, d_bs, s_bs = ctx.meta
        g = grad.reshape(batch, 1)
        d2 = diag.reshape(-1, elems)
        s2 = state.reshape(-1, elems)
        g_state = g_diag = None
        if ctx.needs_input_grad[1]:
            fused = None
            if not (s_bs == 0 and batch > 1):               # a broadcast state needs a sum over the batch
                fused = _real_scale(state, grad.reshape(-1), diag, batch, elems, s_bs, 1, 0, d_bs)
            if fused is not None:
                g_state = fused.reshape(state.shape)
            else:
                g_state = 2 * g * d2 * s2                       # (batch, elems)
                if s_bs == 0 and batch > 1:
                    g_state = g_state.sum(0, keepdim=True)
                g_state = g_state.reshape(state.shape)
        if ctx.needs_input_grad[0]:
            g_diag = g * (s2.real ** 2 + s2.imag ** 2)
            if d_bs == 0 and batch > 1:
                g_diag = g_diag.sum(0, keepdim=True)
            g_diag = g_diag.reshape(diag.shape)
        return g_diag, g_state, None, None, None, None, None


def diag_expectation_value(diag_values: torch.Tensor, state: torch.Tensor):
    """Expectation value of a diagonal operator: sum_k d_k |psi_k|^2  (innerprod.py:49-59).

    `diag_values` broadcasts against (*batch_dims, 2^n) like the reference's product.
    """
    L.require_cuda(diag_values, state)
    st = _complexify(state)
    rdt = _real_dtype(st.dtype)
    if diag_values.is_complex():
        # the reference would return a complex sum; keep its formula with stock torch ops
        return torch.sum(abs_squared(state) * diag_values, dim=-1)
    if diag_values.dtype == torch.float64 and rdt == torch.float32:
        st = st.to(torch.complex128)           # torch promotion: f32 probabilities * f64 diag
        rdt = torch.float64
    diag = diag_values.to(rdt)
    out_shape = tuple(torch.broadcast_shapes(tuple(diag.shape), tuple(st.shape)))
    elems = out_shape[-1]
    out_batch = out_shape[:-1]
    batch = 1
    for s in out_batch:
        batch *= s

    def norm(t):
        shape = tuple(t.shape)
        lead = 1
        for s in shape[:-1]:
            lead *= s
        if shape[-1] == elems and lead == 1:
            return _contig(t), 0
        if shape == out_shape:
            return _contig(t), elems
        return t.expand(out_shape).contiguous(), elems

    if batch * elems == 0:
        return torch.zeros(out_batch, dtype=rdt, device=st.device)
    diag, d_bs = norm(diag)
    st, s_bs = norm(st)
    args = (diag, st, batch, elems, d_bs, s_bs, out_batch)
    if torch.is_grad_enabled() and (diag.requires_grad or st.requires_grad):
        return _DiagExpectation.apply(*args)
    return _DiagExpectation.forward(_Ctx(), *args)


class _InnerProduct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, batch, elems, a_bs, b_bs, out_shape):
        dev = a.device
        out = torch.empty(out_shape, dtype=a.dtype, device=dev)
        with L.on_device(dev):
            nbytes = L.lib().ua_reduce_workspace_bytes(batch, elems)
            ws, ws_ptr = L.workspace(nbytes, dev)
            L.check(L.lib().ua_inner_product(
                L.dtype_code(a.dtype), out.data_ptr(), a.data_ptr(), b.data_ptr(),
                elems, batch, a_bs, b_bs, ws_ptr, nbytes, L.stream_ptr(dev)))
        ctx.save_for_backward(a, b)
        ctx.meta = (batch, elems, a_bs, b_bs)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        a, b = ctx.saved_tensors
        batch, elems, a_bs, b_bs = ctx.meta
        g = grad.reshape(batch, 1)
        a2 = a.reshape(-1, elems)
        b2 = b.reshape(-1, elems)
        g_a = g_b = None
        if ctx.needs_input_grad[0]:       # out = sum conj(a) b  ->  grad_a = conj(g) b
            g_a = g.conj() * b2
            if a_bs == 0 and batch > 1:
                g_a = g_a.sum(0, keepdim=True)
            g_a = g_a.reshape(a.shape)
        if ctx.needs_input_grad[1]:       # grad_b = g a
            g_b = g * a2
            if b_bs == 0 and batch > 1:
                g_b = g_b.sum(0, keepdim=True)
            g_b = g_b.reshape(b.shape)
        return g_a, g_b, None, None, None, None, None


def inner_product(state_1: torch.Tensor, state_2: torch.Tensor):
    """<state_1|state_2> per batch entry; the left entry is conjugated (innerprod.py:62-65)."""
    L.require_cuda(state_1, state_2)
    a = _complexify(state_1)
    b = _complexify(state_2)
    if a.dtype != b.dtype:
        a = a.to(torch.complex128)
        b = b.to(torch.complex128)
    out_shape = tuple(torch.broadcast_shapes(tuple(a.shape), tuple(b.shape)))
    elems = out_shape[-1]
    out_batch = out_shape[:-1]
    batch = 1
    for s in out_batch:
        batch *= s
    if batch * elems == 0:
        return torch.zeros(out_batch, dtype=a.dtype, device=a.device)

    def norm(t):
        shape = tuple(t.shape)
        lead = 1
        for s in shape[:-1]:
            lead *= s
        if shape[-1] == elems and lead == 1:
            return _contig(t), 0
        if shape == out_shape:
            return _contig(t), elems
        return t.expand(out_shape).contiguous(), elems

    a, a_bs = norm(a)
    b, b_bs = norm(b)
    args = (a, b, batch, elems, a_bs, b_bs, out_batch)
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return _InnerProduct.apply(*args)
    return _InnerProduct.forward(_Ctx(), *args)
