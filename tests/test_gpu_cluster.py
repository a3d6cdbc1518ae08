"""GPU parity tests of the register-blocked fused pass (csrc/ua_cluster.cu, entry points
ua_apply_fused_pass_hostmats / ua_apply_fused_pass_scatter_hostmats): against the numpy oracle,
against the shared-memory-matrix pass kernel on the same plan, through the C ABI directly, and by
size-independent properties at the full 30-qubit size.  Tolerance: 1e-5 relative (complex64).
"""
import numpy as np
import pytest
import torch

from conftest import assert_close, rel_err
from oracle import unitair_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ua():
    import unitair_b200
    from unitair_b200 import _lib
    _lib.lib()
    return unitair_b200


def haar(rng, dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return (q * (d / np.abs(d))).astype(np.complex64)


def rnd_state(rng, n, batch=()):
    s = rng.standard_normal(tuple(batch) + (2 ** n,)) + 1j * rng.standard_normal(tuple(batch) + (2 ** n,))
    s /= np.linalg.norm(s, axis=-1, keepdims=True)
    return s.astype(np.complex64)


def random_gates(rng, n, count, one_qubit_share=0.3):
    gates = []
    for _ in range(count):
        if n == 1 or rng.random() < one_qubit_share:
            gates.append(([int(rng.integers(n))], haar(rng, 2)))
        else:
            a, b = rng.choice(n, 2, replace=False)
            gates.append(([int(a), int(b)], haar(rng, 4)))
    return gates


def oracle_run(gates, st):
    for qs, u in gates:
        st = orc.apply_operator(u, qs, st)
    return st


def compiled(gates, n, batch_shape=(), on_host=False, cluster=True, monkeypatch=None, **kw):
    from unitair_b200 import circuit
    monkeypatch.setenv("UA_CLUSTER", "1" if cluster else "0")
    monkeypatch.setattr(circuit, "SMALL_STATE_AMPS", 0)     # device gates take the register-blocked path at every size
    tg = [(qs, torch.from_numpy(u) if on_host else torch.from_numpy(u).cuda()) for qs, u in gates]
    return circuit.CompiledCircuit(tg, n, torch.complex64, batch_shape, **kw)


@pytest.mark.parametrize("n", [4, 5, 6, 8, 11, 12, 13, 15, 17])
@pytest.mark.parametrize("on_host", [False, True])
def test_register_blocked_pass_matches_oracle(ua, monkeypatch, n, on_host):
    rng = np.random.default_rng(1000 + n)
    gates = random_gates(rng, n, 5 * n)
    st = rnd_state(rng, n)
    cc = compiled(gates, n, on_host=on_host, monkeypatch=monkeypatch)
    assert cc.cluster and cc.mats_host is not None, "complex64 shared 1-/2-qubit circuits take the register-blocked path"
    from unitair_b200 import _lib
    before = _lib.launch_count()
    out = cc.run(torch.from_numpy(st).cuda())
    assert _lib.launch_count() - before == cc.num_passes
    assert_close(out.cpu().numpy(), oracle_run(gates, st), "c64", factor=2, what=f"n={n}")
    # the shared-memory-matrix kernel on the same gate list agrees too
    old = compiled(gates, n, cluster=False, monkeypatch=monkeypatch).run(torch.from_numpy(st).cuda())
    assert_close(out.cpu().numpy(), old.cpu().numpy(), "c64", factor=2)


@pytest.mark.parametrize("n,batch", [(6, (3,)), (10, (2, 3)), (12, (5,)), (13, (4,))])
def test_register_blocked_pass_batched_states(ua, monkeypatch, n, batch):
    rng = np.random.default_rng(77 + n)
    gates = random_gates(rng, n, 3 * n)
    st = rnd_state(rng, n, batch)
    out = compiled(gates, n, batch_shape=batch, monkeypatch=monkeypatch).run(torch.from_numpy(st).cuda())
    assert_close(out.cpu().numpy(), oracle_run(gates, st), "c64", factor=2)


def test_only_one_qubit_gates_and_unmerged_lists(ua, monkeypatch):
    """1-qubit gates in clusters (types 6..9) and lists that are not merged first."""
    n = 14
    rng = np.random.default_rng(5)
    gates = [([q], haar(rng, 2)) for q in range(n)] + [([int(rng.integers(n))], haar(rng, 2)) for _ in range(20)]
    st = rnd_state(rng, n)
    for merge in (True, False):
        out = compiled(gates, n, monkeypatch=monkeypatch, merge=merge).run(torch.from_numpy(st).cuda())
        assert_close(out.cpu().numpy(), oracle_run(gates, st), "c64", factor=2, what=f"merge={merge}")
    gates2 = random_gates(rng, n, 60, one_qubit_share=0.5)
    out = compiled(gates2, n, monkeypatch=monkeypatch, merge=False).run(torch.from_numpy(st).cuda())
    assert_close(out.cpu().numpy(), oracle_run(gates2, st), "c64", factor=3)


def test_mixed_with_three_qubit_and_big_gates(ua, monkeypatch):
    """Passes that hold a 3-qubit gate stay on the shared-memory-matrix kernel, bigger gates go
    through the direct kernel, the rest through the register-blocked pass: one circuit."""
    n = 13
    rng = np.random.default_rng(9)
    gates = random_gates(rng, n, 12) + [([2, 9, 5], haar(rng, 8))] + random_gates(rng, n, 12) + \
        [([0, 3, 7, 11], haar(rng, 16))] + random_gates(rng, n, 12) + [([12, 1, 4, 6, 8, 10], haar(rng, 64))] + \
        random_gates(rng, n, 6)
    st = rnd_state(rng, n)
    out = compiled(gates, n, monkeypatch=monkeypatch).run(torch.from_numpy(st).cuda())
    assert_close(out.cpu().numpy(), oracle_run(gates, st), "c64", factor=3)


def test_in_place_and_cuda_graph(ua, monkeypatch):
    n = 16
    rng = np.random.default_rng(16)
    gates = random_gates(rng, n, 40)
    st = torch.from_numpy(rnd_state(rng, n)).cuda()
    cc = compiled(gates, n, monkeypatch=monkeypatch)
    ref = cc.run(cc.run(cc.run(st)))
    buf = st.clone()
    graph = cc.capture_graph(buf)           # the warm-up run applies the circuit once
    graph.replay()
    graph.replay()
    torch.cuda.synchronize()
    assert_close(buf.cpu().numpy(), ref.cpu().numpy(), "c64", factor=3)


def test_c_abi_hostmats_direct_call_and_adjoint(ua):
    """ua_apply_fused_pass_hostmats through ctypes: forward, adjoint (U^H), error codes."""
    from unitair_b200 import _lib as L
    lib = L.lib()
    n, low, high = 13, 7, [8, 9, 10, 11, 12]
    rng = np.random.default_rng(3)
    us = [haar(rng, 4), haar(rng, 2), haar(rng, 4)]
    bits = [(12, 3), (9,), (0, 8)]                     # index-bit positions, gate order (MSB first)
    qubits = [[n - 1 - b for b in bb] for bb in bits]
    mats = np.concatenate([u.reshape(-1) for u in us])
    offs = [0, 16, 20]
    st = rnd_state(rng, n)
    d_in = torch.from_numpy(st).cuda()
    d_out = torch.empty_like(d_in)
    flat = []
    for bb in bits:
        flat += list(bb) + [0] * (3 - len(bb))
    stream = L.stream_ptr(d_in.device)

    def call(adjoint, dtype=0, ks=(2, 1, 2), src=d_in, dst=d_out):
        return lib.ua_apply_fused_pass_hostmats(dtype, dst.data_ptr(), src.data_ptr(), 1 << n, n, low, len(high),
                                                L.int_array(high), 3, L.int_array(list(ks)), L.int_array(flat),
                                                L.ll_array(offs), mats.ctypes.data, adjoint, stream)
    assert call(0) == 0
    torch.cuda.synchronize()
    ref = st
    for qs, u in zip(qubits, us):
        ref = orc.apply_operator(u, qs, ref)
    assert_close(d_out.cpu().numpy(), ref, "c64")
    # adjoint of the pass applied to its output: the gates are unitary but the pass applies the
    # adjoint matrices in the SAME order, so undo them one pass per gate, last gate first
    cur = d_out.clone()
    for g in (2, 1, 0):
        k = len(bits[g])
        rc = lib.ua_apply_fused_pass_hostmats(0, cur.data_ptr(), cur.data_ptr(), 1 << n, n, low, len(high),
                                              L.int_array(high), 1, L.int_array([k]),
                                              L.int_array(list(bits[g]) + [0] * (3 - k)), L.ll_array([offs[g]]),
                                              mats.ctypes.data, 1, stream)
        assert rc == 0
    torch.cuda.synchronize()
    assert_close(cur.cpu().numpy(), st, "c64", factor=2)
    # this path is complex64 / k <= 2 only: the caller falls back to ua_apply_fused_pass
    assert call(0, dtype=1) == L.UA_ERR_UNSUPPORTED
    assert call(0, ks=(2, 1, 3)) == L.UA_ERR_UNSUPPORTED
    assert "supported" in lib.ua_last_error().decode() or "complex64" in lib.ua_last_error().decode()


def test_full_size_properties_30_qubits(ua, monkeypatch):
    """BASELINE size (30 qubits, 8 GiB): norm preservation, agreement with the per-gate kernel on
    sampled amplitudes is impossible to store twice, so use linearity-free invariants: the
    circuit followed by its inverse returns the start state; a basis state stays normalised."""
    n = 30
    free, _ = torch.cuda.mem_get_info()
    if free < 20 * 2 ** 30:
        pytest.skip("needs 20 GiB of device memory")
    rng = np.random.default_rng(30)
    gates = random_gates(rng, n, 90, one_qubit_share=0.4)
    inverse = [(qs, np.ascontiguousarray(u.conj().T)) for qs, u in reversed(gates)]
    st = torch.zeros(1 << n, dtype=torch.complex64, device="cuda")
    st[12345] = 1
    fwd = compiled(gates, n, on_host=True, monkeypatch=monkeypatch)
    bwd = compiled(inverse, n, on_host=True, monkeypatch=monkeypatch)
    fwd.run(st, in_place=True)
    nrm = float(ua.norm_squared(st).item())
    assert abs(nrm - 1) < 1e-4
    assert float(st[12345].abs()) < 0.9, "the circuit must have moved the state"
    bwd.run(st, in_place=True)
    assert abs(float(st[12345].real) - 1) < 1e-4 and abs(float(st[12345].imag)) < 1e-4
    st[12345] = 0
    assert float(torch.linalg.vector_norm(st)) < 1e-4


# --------------------------------------------------------------------------- round-1 advisor findings
def test_six_qubit_gate_in_the_middle_of_a_circuit(ua, monkeypatch):
    """A 6..10-qubit gate after other gates used to call the out-of-place-only generic kernel
    with out == in (EngineError); it now goes through a scratch buffer."""
    n = 12
    rng = np.random.default_rng(12)
    gates = random_gates(rng, n, 8) + [([1, 3, 5, 7, 9, 11], haar(rng, 64))] + random_gates(rng, n, 8)
    st = rnd_state(rng, n, (2,))
    for in_place in (False, True):
        buf = torch.from_numpy(st).cuda()
        out = compiled(gates, n, batch_shape=(2,), monkeypatch=monkeypatch).run(buf, in_place=in_place)
        assert_close(out.cpu().numpy(), oracle_run(gates, st), "c64", factor=3, what=f"in_place={in_place}")
    tg = [(qs, torch.from_numpy(u).cuda()) for qs, u in gates]
    out = ua.circuit.apply_gates(tg, torch.from_numpy(st).cuda())
    assert_close(out.cpu().numpy(), oracle_run(gates, st), "c64", factor=3)


def test_adjoint_gradient_of_shared_gate_next_to_batched_gates(ua):
    """assume_unitary=True with a trainable SHARED gate among per-entry gates: its gradient is the
    sum over the batch (used to raise in reshape)."""
    n, B = 6, 5
    rng = np.random.default_rng(60)
    shared = torch.from_numpy(haar(rng, 4)).cuda().requires_grad_(True)
    shared1 = torch.from_numpy(haar(rng, 2)).cuda().requires_grad_(True)
    batched = torch.from_numpy(np.stack([haar(rng, 4) for _ in range(B)])).cuda().requires_grad_(True)
    batched1 = torch.from_numpy(np.stack([haar(rng, 2) for _ in range(B)])).cuda().requires_grad_(True)
    st = torch.from_numpy(rnd_state(rng, n, (B,))).cuda()
    diag = torch.from_numpy(rng.standard_normal(2 ** n).astype(np.float32)).cuda()

    def gate_list():
        return [([0, 3], shared), ([2], batched1), ([4, 1], batched), ([5], shared1), ([3, 2], shared), ([0, 5], batched)]

    def loss_of(out):
        return ua.diag_expectation_value(diag, out).sum()

    loss_of(ua.circuit.apply_gates(gate_list(), st, assume_unitary=True)).backward()
    got = [p.grad.clone() for p in (shared, shared1, batched, batched1)]
    for p in (shared, shared1, batched, batched1):
        p.grad = None
    out = st
    for qs, m in gate_list():                      # per-gate autograd path (one node per gate)
        out = ua.simulation.apply_operator(m, qs, out)
    loss_of(out).backward()
    want = [p.grad for p in (shared, shared1, batched, batched1)]
    for g, w, name in zip(got, want, ("shared 2q", "shared 1q", "batched 2q", "batched 1q")):
        assert g.shape == w.shape, name
        assert rel_err(g.cpu().numpy(), w.cpu().numpy()) < 2e-5, name


def test_measure_rejects_zero_norm_state(ua):
    st = torch.zeros(1 << 8, dtype=torch.complex64, device="cuda")
    with pytest.raises(ValueError):
        ua.simulation.measure(st, 10)
    st[3] = float("nan")
    with pytest.raises(ValueError):
        ua.simulation.measure(st, 10)


def test_native_backward_of_real_reductions(ua):
    """abs_squared / norm_squared / diag_expectation_value backward in one native pass
    (ua_real_scale) against torch autograd of the reference formulas (innerprod.py:26,46,59)."""
    rng = np.random.default_rng(8)
    from unitair_b200 import _lib
    for cdt, rdt, tol in ((torch.complex64, torch.float32, 1e-5), (torch.complex128, torch.float64, 1e-12)):
        for shape in ((64,), (3, 32), (2, 3, 16), (5,)):           # (5,): odd complex64 rows use the eager formula
            x = torch.from_numpy(rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).to(cdt).cuda()
            d = torch.from_numpy(rng.standard_normal(shape[-1])).to(rdt).cuda()
            db = torch.from_numpy(rng.standard_normal(shape)).to(rdt).cuda()
            w = torch.from_numpy(rng.standard_normal(shape)).to(rdt).cuda()
            wr = torch.from_numpy(rng.standard_normal(shape[:-1] or (1,))).to(rdt).cuda().reshape(shape[:-1])
            cases = [
                (lambda s: (ua.abs_squared(s) * w).sum(), lambda s: ((s.real ** 2 + s.imag ** 2) * w).sum()),
                (lambda s: (ua.norm_squared(s) * wr).sum(), lambda s: ((s.real ** 2 + s.imag ** 2).sum(-1) * wr).sum()),
                (lambda s: (ua.diag_expectation_value(d, s) * wr).sum(), lambda s: (((s.real ** 2 + s.imag ** 2) * d).sum(-1) * wr).sum()),
                (lambda s: (ua.diag_expectation_value(db, s) * wr).sum(), lambda s: (((s.real ** 2 + s.imag ** 2) * db).sum(-1) * wr).sum()),
            ]
            for ours, ref in cases:
                a = x.clone().requires_grad_(True)
                b = x.clone().requires_grad_(True)
                before = _lib.launch_count()
                ours(a).backward()
                launched = _lib.launch_count() - before
                ref(b).backward()
                assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < tol
                if shape[-1] % 2 == 0:
                    assert launched >= 2, "forward and backward must both be native kernels"
