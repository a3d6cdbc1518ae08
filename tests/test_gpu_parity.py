"""GPU parity tests: the CUDA engine (through the C ABI) against the numpy oracle and the
golden vectors produced by the unmodified reference.

Tolerances are BASELINE.json's: 1e-5 relative for complex64, 1e-12 for complex128
(norm-wise relative error), gradients included.  Pure data movement is bit-exact.
"""
import itertools
import math

import numpy as np
import pytest
import torch

from conftest import assert_close, rel_err
from oracle import unitair_oracle as orc

pytestmark = pytest.mark.gpu

CD = {"c64": torch.complex64, "c128": torch.complex128}
NPC = {"c64": np.complex64, "c128": np.complex128}


@pytest.fixture(scope="module")
def ua():
    import unitair_b200
    from unitair_b200 import _lib
    _lib.lib()   # fail loudly if the native library is missing
    return unitair_b200


def dev(x):
    return torch.from_numpy(np.array(x, copy=True)).cuda()   # keeps 0-d arrays 0-d


def host(t):
    return t.detach().cpu().numpy()


def rnd_c(rng, shape, dt, scale=1.0):
    return (scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))).astype(NPC[dt])


def rnd_state(rng, n, batch, dt):
    s = rng.standard_normal(tuple(batch) + (2 ** n,)) + 1j * rng.standard_normal(tuple(batch) + (2 ** n,))
    s /= np.linalg.norm(s, axis=-1, keepdims=True)
    return s.astype(NPC[dt])


def haar(rng, dim, dt):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return (q * (d / np.abs(d))).astype(NPC[dt])


# --------------------------------------------------------------------------- golden
def test_apply_operator_golden(ua, golden):
    arr = golden.arrays("apply_operator")
    for c in golden.manifest["apply_operator"]:
        k = c["key"]
        out = ua.simulation.apply_operator(operator=dev(arr[k + "_op"]), qubits=c["qubits"],
                                           state=dev(arr[k + "_state"]))
        ref = arr[k + "_out"]
        assert tuple(out.shape) == ref.shape and out.is_contiguous(), c
        assert_close(host(out), ref, c["dtype"], what=str(c))


def test_apply_all_qubits_golden(ua, golden):
    arr = golden.arrays("apply_all")
    for c in golden.manifest["apply_all_qubits"]:
        k = c["key"]
        out = ua.simulation.apply_all_qubits(operator=dev(arr[k + "_op"]), state=dev(arr[k + "_state"]))
        assert tuple(out.shape) == arr[k + "_out"].shape, c
        assert_close(host(out), arr[k + "_out"], c["dtype"], factor=3, what=str(c))


def test_apply_phase_golden(ua, golden):
    arr = golden.arrays("phase")
    for c in golden.manifest["apply_phase"]:
        k = c["key"]
        out = ua.simulation.apply_phase(dev(arr[k + "_angles"]), dev(arr[k + "_state"]))
        ref = arr[k + "_out"]
        assert tuple(out.shape) == ref.shape, c
        assert str(out.dtype) == c["out_dtype"], c
        key = "c64" if c["angle_dtype"] == "f32" else ref.dtype
        assert_close(host(out), ref, key, what=str(c))


def test_reductions_golden(ua, golden):
    arr = golden.arrays("reductions")
    for c in golden.manifest["reductions"]:
        k = c["key"]
        st, st2 = dev(arr[k + "_state"]), dev(arr[k + "_state2"])
        dt = c["dtype"]
        a2 = ua.abs_squared(st)
        assert a2.dtype == (torch.float32 if dt == "c64" else torch.float64)
        assert_close(host(a2), arr[k + "_abs2"], dt)
        assert_close(host(ua.norm_squared(st)), arr[k + "_norm2"], dt)
        assert_close(host(ua.diag_expectation_value(dev(arr[k + "_diag"]), st)), arr[k + "_dexp"], dt, factor=5)
        assert_close(host(ua.diag_expectation_value(dev(arr[k + "_diagb"]), st)), arr[k + "_dexpb"], dt, factor=5)
        ip = ua.inner_product(st, st2)
        assert ip.dtype == CD[dt]
        assert_close(host(ip), arr[k + "_inner"], dt, factor=5)


def test_grads_apply_operator_golden(ua, golden):
    arr = golden.arrays("grads")
    for c in golden.manifest["grads_apply_operator"]:
        k = c["key"]
        op = dev(arr[k + "_op"]).requires_grad_(True)
        st = dev(arr[k + "_state"]).requires_grad_(True)
        w = dev(arr[k + "_w"])
        out = ua.simulation.apply_operator(operator=op, qubits=c["qubits"], state=st)
        loss = (out * w.conj()).real.sum() + (out.abs() ** 2).sum() * 0.5
        g_op, g_st = torch.autograd.grad(loss, (op, st))
        assert_close(host(g_op), arr[k + "_gop"], c["dtype"], factor=5, what="grad_op " + str(c))
        assert_close(host(g_st), arr[k + "_gstate"], c["dtype"], factor=5, what="grad_state " + str(c))


def test_grads_phase_expectation_golden(ua, golden):
    arr = golden.arrays("grads")
    for c in golden.manifest["grads_phase_expectation"]:
        k = c["key"]
        ang = dev(arr[k + "_angles"]).requires_grad_(True)
        st = dev(arr[k + "_state"]).requires_grad_(True)
        w = dev(arr[k + "_w"])
        diag = dev(arr[k + "_diag"])
        out = ua.simulation.apply_phase(ang, st)
        loss = (out * w.conj()).real.sum() + ua.diag_expectation_value(diag, out).sum()
        g_ang, g_st = torch.autograd.grad(loss, (ang, st))
        assert_close(host(loss), arr[k + "_loss"], c["dtype"], factor=5)
        assert tuple(g_ang.shape) == arr[k + "_gangles"].shape
        assert_close(host(g_ang), arr[k + "_gangles"], c["dtype"], factor=10, what="grad_angles " + str(c))
        assert_close(host(g_st), arr[k + "_gstate"], c["dtype"], factor=10, what="grad_state " + str(c))


def test_docs_known_answers(ua):
    c64 = torch.complex64
    q = torch.tensor([[1, 5 - 1j], [5 + 1j, -1]], dtype=c64).cuda()
    ket0 = torch.tensor([1, 0], dtype=c64).cuda()
    ket1 = torch.tensor([0, 1], dtype=c64).cuda()
    np.testing.assert_allclose(host(ua.simulation.apply_operator(q, (0,), ket0)), [1, 5 + 1j])
    np.testing.assert_allclose(host(ua.simulation.apply_operator(q, (0,), torch.stack([ket0, ket1]))),
                               [[1, 5 + 1j], [5 - 1j, -1]])
    h = ua.gates.hadamard(device="cuda")
    np.testing.assert_allclose(host(ua.simulation.apply_operator(h, (0,), ket0)), [0.70710678] * 2, rtol=1e-6)
    s = ua.unit_vector(0, num_qubits=2, device="cuda")
    s = ua.simulation.apply_operator(h, (0,), s)
    s = ua.simulation.apply_operator(ua.gates.cnot(device="cuda"), (0, 1), s)
    np.testing.assert_allclose(host(s), [0.70710678, 0, 0, 0.70710678], rtol=1e-6, atol=1e-7)
    s = ua.unit_vector(0, num_qubits=3, device="cuda")
    assert int(ua.simulation.apply_operator(ua.gates.pauli_x(device="cuda"), (0,), s).abs().argmax()) == 4


# --------------------------------------------------------------------------- seeded vs oracle
@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_single_qubit_every_target(ua, dt):
    rng = np.random.default_rng(11)
    n = 14
    st = rnd_state(rng, n, (), dt)
    for q in range(n):
        u = rnd_c(rng, (2, 2), dt)
        out = ua.simulation.apply_operator(dev(u), (q,), dev(st))
        assert_close(host(out), orc.apply_operator(u, (q,), st), dt, what=f"q={q}")


@pytest.mark.parametrize("dt", ["c64", "c128"])
@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_multi_qubit_random_targets(ua, dt, k):
    rng = np.random.default_rng(100 + k)
    n = 13
    st = rnd_state(rng, n, (), dt)
    combos = [tuple(range(k)), tuple(range(n - k, n)), tuple(reversed(range(n - k, n)))]
    for _ in range(8):
        combos.append(tuple(rng.permutation(n)[:k].tolist()))
    for qs in combos:
        u = rnd_c(rng, (2 ** k, 2 ** k), dt)
        out = ua.simulation.apply_operator(dev(u), qs, dev(st))
        assert_close(host(out), orc.apply_operator(u, qs, st), dt, what=f"qubits={qs}")


@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_batch_structures(ua, dt):
    rng = np.random.default_rng(5)
    n = 9
    for k in (1, 2, 3):
        qs = tuple(rng.permutation(n)[:k].tolist())
        for sb, ob in [((5,), ()), ((5,), (5,)), ((), (7,)), ((2, 3), (2, 3)), ((3, 2), (2,)),
                       ((3, 1, 2), ()), ((1,), (1,)), ((4, 1), (4, 1))]:
            u = rnd_c(rng, tuple(ob) + (2 ** k, 2 ** k), dt)
            st = rnd_state(rng, n, sb, dt)
            out = ua.simulation.apply_operator(dev(u), qs, dev(st))
            ref = orc.apply_operator(u, qs, st)
            assert tuple(out.shape) == ref.shape
            assert_close(host(out), ref, dt, what=f"k={k} sb={sb} ob={ob}")


def test_tiny_states_and_odd_batches(ua):
    rng = np.random.default_rng(6)
    for dt in ("c64", "c128"):
        for n in (1, 2, 3, 4):
            for k in range(1, n + 1):
                for batch in [(), (1,), (3,), (257,)]:
                    qs = tuple(rng.permutation(n)[:k].tolist())
                    u = rnd_c(rng, (2 ** k, 2 ** k), dt)
                    st = rnd_state(rng, n, batch, dt)
                    out = ua.simulation.apply_operator(dev(u), qs, dev(st))
                    assert_close(host(out), orc.apply_operator(u, qs, st), dt, what=f"n={n} k={k} {batch}")


def test_noncontiguous_and_misaligned_inputs(ua):
    rng = np.random.default_rng(7)
    n = 6
    big = dev(rnd_state(rng, n, (4, 3), "c64"))
    st = big.transpose(0, 1)                       # non-contiguous batch
    u = dev(rnd_c(rng, (2, 2), "c64"))
    out = ua.simulation.apply_operator(u, (2,), st)
    assert_close(host(out), orc.apply_operator(host(u), (2,), host(st)), "c64")
    assert out.is_contiguous()
    # gate as a permuted view (like the reference's nested_stack(roll=True) output)
    ub = dev(rnd_c(rng, (2, 2, 3), "c64")).permute(2, 0, 1)
    stb = dev(rnd_state(rng, n, (3,), "c64"))
    out = ua.simulation.apply_operator(ub, (4,), stb)
    assert_close(host(out), orc.apply_operator(host(ub), (4,), host(stb)), "c64")
    # slice with an 8-byte (not 16-byte) aligned start
    flat = dev(rnd_c(rng, (130,), "c64"))
    view = flat[1:129]
    out = ua.simulation.apply_operator(u, (0,), view)
    assert_close(host(out), orc.apply_operator(host(u), (0,), host(view)), "c64")
    # inputs are never modified
    before = stb.clone()
    ua.simulation.apply_operator(u, (1,), stb)
    assert torch.equal(before, stb)


def test_real_dtype_states(ua):
    rng = np.random.default_rng(8)
    st = rng.standard_normal((3, 16)).astype(np.float32)
    u = rng.standard_normal((4, 4)).astype(np.float32)
    out = ua.simulation.apply_operator(dev(u), (3, 1), dev(st))
    assert out.dtype == torch.float32
    assert_close(host(out), orc.apply_operator(u, (3, 1), st), "c64")


def test_generic_large_k(ua):
    rng = np.random.default_rng(9)
    n, k = 9, 6
    qs = tuple(rng.permutation(n)[:k].tolist())
    for dt in ("c64", "c128"):
        u = rnd_c(rng, (64, 64), dt, 0.3)
        st = rnd_state(rng, n, (2,), dt)
        out = ua.simulation.apply_operator(dev(u), qs, dev(st))
        assert_close(host(out), orc.apply_operator(u, qs, st), dt, factor=3)


def test_errors(ua):
    st = torch.zeros(8, dtype=torch.complex64, device="cuda")
    op = torch.eye(2, dtype=torch.complex64, device="cuda")
    with pytest.raises(ValueError):
        ua.simulation.apply_operator(op, (3,), st)
    with pytest.raises(ValueError):
        ua.simulation.apply_operator(op, (-1,), st)
    with pytest.raises(ValueError):
        ua.simulation.apply_operator(op, (0, 1), st)
    with pytest.raises(ValueError):
        ua.simulation.apply_operator(torch.eye(4, dtype=torch.complex64, device="cuda"), (1, 1), st)
    with pytest.raises(ua.states.StateShapeError):
        ua.simulation.apply_operator(op, (0,), torch.zeros(6, dtype=torch.complex64, device="cuda"))
    with pytest.raises(RuntimeError):
        ua.simulation.apply_operator(torch.zeros(3, 3, dtype=torch.complex64, device="cuda"), (0,), st)
    with pytest.raises(RuntimeError):
        ua.simulation.apply_operator(op.to(torch.complex128), (0,), st)
    with pytest.raises(ValueError):
        ua.simulation.apply_all_qubits(torch.eye(4, dtype=torch.complex64, device="cuda"), st)
    with pytest.raises(RuntimeError):   # no CPU path
        ua.simulation.apply_operator(op.cpu(), (0,), st.cpu())
    with pytest.raises(RuntimeError):
        ua.simulation.apply_operator(
            torch.zeros(3, 2, 2, 2, dtype=torch.complex64, device="cuda"), (0,),
            torch.zeros(2, 8, dtype=torch.complex64, device="cuda"))


# --------------------------------------------------------------------------- C ABI level
def test_c_abi_inplace_and_adjoint(ua):
    from unitair_b200 import _engine
    rng = np.random.default_rng(12)
    for dt in ("c64", "c128"):
        n = 12
        for k in (1, 2, 3, 5):
            qs = rng.permutation(n)[:k].tolist()
            u = rnd_c(rng, (2 ** k, 2 ** k), dt)
            st = rnd_state(rng, n, (3,), dt)
            ref = orc.apply_operator(u, qs, st)
            ref_adj = orc.apply_operator(np.conj(u.T), qs, st)
            buf = dev(st)
            _engine.launch_gate(buf, buf, dev(u), n, k, qs, 3, 1 << n, 0, False)     # in place
            assert_close(host(buf), ref, dt, what=f"inplace k={k}")
            buf = dev(st)
            out = torch.empty_like(buf)
            _engine.launch_gate(out, buf, dev(u), n, k, qs, 3, 1 << n, 0, True)      # adjoint
            assert_close(host(out), ref_adj, dt, what=f"adjoint k={k}")


def test_c_abi_rejects_bad_arguments(ua):
    from unitair_b200 import _engine, _lib
    st = torch.zeros(16, dtype=torch.complex64, device="cuda")
    g = torch.eye(2, dtype=torch.complex64, device="cuda")
    with pytest.raises(_lib.EngineError):
        _engine.launch_gate(st.clone(), st, g, 4, 1, [4], 1, 16, 0, False)
    with pytest.raises(_lib.EngineError):
        _engine.launch_gate(st.clone(), st, g, 4, 2, [1, 1], 1, 16, 0, False)
    with pytest.raises(_lib.EngineError):
        _engine.launch_gate(st.clone(), st, g, 4, 1, [0], 1, 8, 0, False)
    assert _lib.launch_count() > 0


# --------------------------------------------------------------------------- phase / reductions
@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_phase_seeded(ua, dt):
    rng = np.random.default_rng(13)
    rdt = np.float32 if dt == "c64" else np.float64
    for ash, ssh in [((1 << 14,), (1 << 14,)), ((1 << 10,), (7, 1 << 10)), ((7, 1), (7, 1 << 10)),
                     ((7, 1 << 10), (7, 1 << 10)), ((), (3, 64)), ((3, 5, 16), (16,)),
                     ((5, 1, 8), (5, 3, 8)), ((6,), (6,)), ((3, 1), (3, 1)), ((4, 3), (4, 3))]:
        ang = (rng.random(ash) * 2 * np.pi).astype(rdt)
        st = rnd_c(rng, ssh, dt)
        out = ua.simulation.apply_phase(dev(ang), dev(st))
        ref = orc.apply_phase(ang, st)
        assert tuple(out.shape) == ref.shape
        assert_close(host(out), ref, dt, what=f"{ash} {ssh}")
    # reference's own property tests (tests/test_operations.py:23-40, tests/test_unitary.py:10-36)
    st = dev(rnd_c(rng, (3, 256), dt))
    ang = dev((rng.random((3, 256)) * 6).astype(rdt))
    back = ua.simulation.apply_phase(-ang, ua.simulation.apply_phase(ang, st))
    assert torch.allclose(back, st, atol=1e-4)
    rot = ua.simulation.apply_phase(ang, st)
    basic = torch.complex(ang.cos() * st.real + ang.sin() * st.imag, -ang.sin() * st.real + ang.cos() * st.imag)
    assert torch.isclose(rot, basic).all()


@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_reductions_seeded(ua, dt):
    rng = np.random.default_rng(14)
    rdt = np.float32 if dt == "c64" else np.float64
    for shape in [(1 << 20,), (3, 1 << 16), (4096, 64), (5, 3, 2), (1 << 18,)]:
        a = rnd_c(rng, shape, dt)
        b = rnd_c(rng, shape, dt)
        d = rng.standard_normal(shape).astype(rdt)
        assert_close(host(ua.abs_squared(dev(a))), orc.abs_squared(a), dt)
        # the oracle sums in the low precision; compare with a float64 evaluation
        a128, b128 = a.astype(np.complex128), b.astype(np.complex128)
        assert_close(host(ua.norm_squared(dev(a))), orc.norm_squared(a128).astype(rdt), dt)
        assert_close(host(ua.diag_expectation_value(dev(d), dev(a))),
                     orc.diag_expectation_value(d.astype(np.float64), a128).astype(rdt), dt, factor=20)
        assert_close(host(ua.inner_product(dev(a), dev(b))),
                     orc.inner_product(a128, b128).astype(NPC[dt]), dt, factor=20)


def test_reduction_grads_match_torch(ua):
    rng = np.random.default_rng(15)
    a = dev(rnd_c(rng, (3, 64), "c128")).requires_grad_(True)
    b = dev(rnd_c(rng, (3, 64), "c128")).requires_grad_(True)
    d = dev(rng.standard_normal((64,)))
    w = dev(rnd_c(rng, (3,), "c128"))
    wr = dev(rng.standard_normal((3, 64)))

    def loss(fn_abs, fn_norm, fn_diag, fn_inner):
        return ((fn_abs(a) * wr).sum() + (fn_norm(a) * wr[:, 0]).sum() + fn_diag(d, a).sum() * 0.7
                + (fn_inner(a, b) * w.conj()).real.sum())

    l1 = loss(ua.abs_squared, ua.norm_squared, ua.diag_expectation_value, ua.inner_product)
    g1 = torch.autograd.grad(l1, (a, b))
    l2 = loss(lambda s: (s.conj() * s).real, lambda s: (s.conj() * s).real.sum(-1),
              lambda dd, s: ((s.conj() * s).real * dd).sum(-1), lambda x, y: (x.conj() * y).sum(-1))
    g2 = torch.autograd.grad(l2, (a, b))
    assert_close(host(l1), host(l2), "c128", factor=10)
    assert_close(host(g1[0]), host(g2[0]), "c128", factor=10)
    assert_close(host(g1[1]), host(g2[1]), "c128", factor=10)


# --------------------------------------------------------------------------- gate gradients
@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_gate_grad_seeded(ua, dt):
    rng = np.random.default_rng(16)
    n = 11
    for k, sb, ob in [(1, (), ()), (2, (), ()), (3, (), ()), (4, (), ()), (5, (), ()),
                      (1, (6,), ()), (1, (6,), (6,)), (2, (6,), (6,)), (1, (), (6,)), (2, (64,), ())]:
        qs = rng.permutation(n)[:k].tolist()
        u = rnd_c(rng, tuple(ob) + (2 ** k, 2 ** k), dt)
        st = rnd_state(rng, n, sb, dt)
        uo = dev(u).requires_grad_(True)
        so = dev(st).requires_grad_(True)
        out = ua.simulation.apply_operator(uo, qs, so)
        g = rnd_c(rng, tuple(out.shape), dt)
        g_u, g_s = torch.autograd.grad(out, (uo, so), grad_outputs=dev(g))
        ref_u, ref_s = orc.apply_operator_grads(u.astype(np.complex128), qs, st.astype(np.complex128),
                                                g.astype(np.complex128))
        assert_close(host(g_u), ref_u.astype(NPC[dt]), dt, factor=5, what=f"gU k={k} {sb} {ob}")
        assert_close(host(g_s), ref_s.astype(NPC[dt]), dt, factor=5, what=f"gS k={k} {sb} {ob}")


# --------------------------------------------------------------------------- circuits
def _run_c2(ua, arr, m, fused):
    gates = [(g["qubits"], dev(arr[f"c2_g{g['g']}"])) for g in m["c2"]["gates"]]
    psi = dev(arr["c2_state"])
    if fused:
        return ua.circuit.apply_gates(gates, psi)
    for qs, u in gates:
        psi = ua.simulation.apply_operator(u, qs, psi)
    return psi


def test_circuits_golden(ua, golden):
    arr = golden.arrays("circuits")
    m = golden.manifest["circuits"]
    g = ua.gates
    # C1
    psi = dev(arr["c1_state"])
    theta = dev(arr["c1_theta"])
    n = m["c1"]["n"]
    h = g.hadamard(device="cuda")
    for q in range(n):
        psi = ua.simulation.apply_operator(h, (q,), psi)
    for q in range(n):
        psi = ua.simulation.apply_operator(g.exp_x(theta[q]), (q,), psi)
    cn = g.cnot(device="cuda")
    for q in range(n - 1):
        psi = ua.simulation.apply_operator(cn, (q, q + 1), psi)
    assert_close(host(psi), arr["c1_out"], "c64", factor=5, what="C1")
    # C2 per-op and fused
    assert_close(host(_run_c2(ua, arr, m, False)), arr["c2_out"], "c64", factor=10, what="C2 per-op")
    assert_close(host(_run_c2(ua, arr, m, True)), arr["c2_out"], "c64", factor=10, what="C2 fused")
    # C4
    psi = dev(arr["c4_state"])
    nb = len(m["c4"]["blocks"]) // m["c4"]["layers"]
    for l in range(m["c4"]["layers"]):
        for b in m["c4"]["blocks"][l * nb:(l + 1) * nb]:
            psi = ua.simulation.apply_operator(dev(arr[f"c4_g{b['g']}"]), b["qubits"], psi)
        psi = ua.simulation.apply_phase(dev(arr[f"c4_ang{l}"]), psi)
    assert_close(host(psi), arr["c4_out"], "c128", factor=10, what="C4")


def test_circuit_c3_gradients_golden(ua, golden):
    arr = golden.arrays("circuits")
    m = golden.manifest["circuits"]["c3"]
    n, layers = m["n"], m["layers"]
    g = ua.gates
    cn = g.cnot(device="cuda")
    z0 = torch.where((torch.arange(2 ** n, device="cuda") >> (n - 1)) & 1 == 0, 1.0, -1.0)
    for tag, theta_np in (("c3", arr["c3_theta"]), ("c3b", arr["c3b_theta"])):
        theta = dev(theta_np).requires_grad_(True)
        psi = dev(arr["c3_state"])
        for l in range(layers):
            for q in range(n):
                t = theta[l, q] if tag == "c3" else theta[:, l, q]
                psi = ua.simulation.apply_operator(g.exp_y(t[..., 0]), (q,), psi)
                psi = ua.simulation.apply_operator(g.exp_z(t[..., 1]), (q,), psi)
            for q in range(n - 1):
                psi = ua.simulation.apply_operator(cn, (q, q + 1), psi)
        loss = ua.diag_expectation_value(z0, psi).sum()
        g_theta, = torch.autograd.grad(loss, theta)
        assert_close(host(psi), arr[tag + "_out"], "c64", factor=10, what=tag + " state")
        assert_close(host(loss), arr[tag + "_loss"], "c64", factor=20, what=tag + " loss")
        assert_close(host(g_theta), arr[tag + "_gtheta"], "c64", factor=30, what=tag + " grad")


@pytest.mark.parametrize("dt", ["c64", "c128"])
@pytest.mark.parametrize("n", [5, 10, 16, 19])
def test_fused_passes_match_per_gate(ua, dt, n):
    rng = np.random.default_rng(17 + n)
    st = rnd_state(rng, n, (), dt)
    gates = []
    for layer in range(3):
        for q in range(n):
            gates.append(([q], haar(rng, 2, dt)))
        pi = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            gates.append(([pi[j], pi[j + 1]], haar(rng, 4, dt)))
        if n >= 3:
            tri = rng.permutation(n)[:3].tolist()
            gates.append((tri, haar(rng, 8, dt)))
    dgates = [(qs, dev(u)) for qs, u in gates]
    fused = ua.circuit.apply_gates(dgates, dev(st))
    psi = dev(st)
    for qs, u in dgates:
        psi = ua.simulation.apply_operator(u, qs, psi)
    assert_close(host(fused), host(psi), dt, factor=5, what="fused vs per-gate")
    if n <= 16:
        ref = st
        for qs, u in gates:
            ref = orc.apply_operator(u, qs, ref)
        assert_close(host(fused), ref, dt, factor=10, what="fused vs oracle")


def test_fused_passes_batched(ua):
    rng = np.random.default_rng(23)
    n, B = 8, 5
    st = rnd_state(rng, n, (B,), "c64")
    gates = []
    for q in range(n):
        gates.append(([q], rnd_c(rng, (B, 2, 2), "c64")))       # per-entry gate
        gates.append(([q, (q + 3) % n], haar(rng, 4, "c64")))   # shared gate
    fused = ua.circuit.apply_gates([(qs, dev(u)) for qs, u in gates], dev(st))
    ref = st
    for qs, u in gates:
        ref = orc.apply_operator(u, qs, ref)
    assert_close(host(fused), ref, "c64", factor=10)


@pytest.mark.parametrize("dt", ["c64", "c128"])
def test_apply_all_qubits_seeded(ua, dt):
    rng = np.random.default_rng(18)
    for n, sb, ob in [(15, (), ()), (17, (), ()), (9, (4,), (4,)), (9, (4,), ()), (6, (), (3,)), (20, (), ())]:
        u = rnd_c(rng, tuple(ob) + (2, 2), dt, 0.7)
        st = rnd_state(rng, n, sb, dt)
        out = ua.simulation.apply_all_qubits(dev(u), dev(st))
        if n <= 17:
            ref = orc.apply_all_qubits(u, st)
        else:
            psi = dev(st)
            for q in range(n):
                psi = ua.simulation.apply_operator(dev(u), (q,), psi)
            ref = host(psi)
        assert tuple(out.shape) == ref.shape
        assert_close(host(out), ref, dt, factor=5, what=f"n={n} {sb} {ob}")
    # property from the reference's tests (tests/test_unitary.py:38-52): equals apply_to_qubits
    rot = ua.gates.exp_x(0.3).cuda()
    st = dev(rnd_state(rng, 7, (), "c64"))
    a = ua.simulation.apply_all_qubits(rot, st)
    b = ua.simulation.apply_to_qubits([rot] * 7, range(7), st)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)


def test_apply_all_qubits_autograd(ua):
    rng = np.random.default_rng(19)
    theta = torch.tensor(0.37, device="cuda", dtype=torch.float64, requires_grad=True)
    st = dev(rnd_state(rng, 5, (3,), "c128"))
    out = ua.simulation.apply_all_qubits(ua.gates.exp_y(theta, dtype=torch.complex128), st)
    w = dev(rng.standard_normal((3, 32)))
    loss = (ua.abs_squared(out) * w).sum()
    g, = torch.autograd.grad(loss, theta)
    eps = 1e-6
    with torch.no_grad():
        lp = (ua.abs_squared(ua.simulation.apply_all_qubits(ua.gates.exp_y(theta + eps, dtype=torch.complex128), st)) * w).sum()
        lm = (ua.abs_squared(ua.simulation.apply_all_qubits(ua.gates.exp_y(theta - eps, dtype=torch.complex128), st)) * w).sum()
    assert abs(float(g) - float((lp - lm) / (2 * eps))) < 1e-6 * max(1.0, abs(float(g)))


# --------------------------------------------------------------------------- permutations
def test_swap_and_permute(ua):
    rng = np.random.default_rng(20)
    for dt in ("c64", "c128"):
        for n, batch in [(1, ()), (5, ()), (8, (3,)), (12, ()), (4, (2, 3))]:
            st = rnd_state(rng, n, batch, dt)
            for _ in range(4):
                perm = rng.permutation(n).tolist()
                out = ua.simulation.permute_qubits(perm, dev(st))
                assert np.array_equal(host(out), orc.permute_qubits(perm, st)), (n, perm)
            if n >= 2:
                i, j = rng.permutation(n)[:2].tolist()
                sw = ua.simulation.swap(dev(st), (i, j))
                assert np.array_equal(host(sw), orc.swap(st, (i, j)))
                assert torch.equal(ua.simulation.swap(sw, (i, j)), dev(st))        # involutive
                assert torch.equal(ua.simulation.swap(dev(st), (j, i)), sw)        # symmetric
    # known cases from the reference (tests/test_unitary.py:94-104)
    s = torch.tensor([2.24 + .3j, 1. + .2j, 1. + .2j, 73. - .13j], device="cuda")
    assert torch.equal(ua.simulation.swap(s, (0, 1)), s)


# --------------------------------------------------------------------------- full size properties
def test_full_size_properties_30_qubits(ua):
    """BASELINE.json's 30-qubit complex64 size: size-independent properties only."""
    free, _ = torch.cuda.mem_get_info()
    n = 30 if free > 40 * 2 ** 30 else 27
    rng = np.random.default_rng(21)
    psi = torch.zeros(2 ** n, dtype=torch.complex64, device="cuda")
    psi[0] = 1
    h = ua.gates.hadamard(device="cuda")
    psi = ua.simulation.apply_all_qubits(h, psi)          # H^n |0> = uniform superposition
    assert abs(float(ua.norm_squared(psi)) - 1.0) < 1e-5
    amp = 2.0 ** (-n / 2)
    assert float((psi.real - amp).abs().max()) < 1e-3 * amp and float(psi.imag.abs().max()) < 1e-3 * amp
    # unitary gates on high / low / mixed targets keep the norm; U then U^H is the identity
    from unitair_b200 import _engine
    for qs in [(0,), (n - 1,), (0, n - 1), (3, n - 2, 11), (n // 2, 1)]:
        k = len(qs)
        u = dev(haar(rng, 2 ** k, "c64"))
        out = ua.simulation.apply_operator(u, qs, psi)
        assert abs(float(ua.norm_squared(out)) - 1.0) < 2e-5, qs
        back = torch.empty_like(out)
        _engine.launch_gate(back, out, u, n, k, list(qs), 1, 1 << n, 0, True)
        err = float(ua.norm_squared(back - psi)) ** 0.5
        assert err < 1e-5, (qs, err)
        del out, back
    # linearity on a 5-qubit block
    u5 = dev(haar(rng, 32, "c64"))
    qs = (2, n - 1, 7, n - 3, 12)
    out = ua.simulation.apply_operator(u5, qs, psi)
    assert abs(float(ua.norm_squared(out)) - 1.0) < 5e-5
    z = torch.where((torch.arange(2 ** 10, device="cuda") & 1) == 0, 1.0, -1.0).repeat(2 ** (n - 10))
    ez = float(ua.diag_expectation_value(z, psi))
    assert abs(ez) < 1e-5


# --------------------------------------------------------------------------- adjoint-method circuit autograd
def test_adjoint_circuit_gradients(ua, golden):
    """circuit.apply_gates(assume_unitary=True): one autograd node, no saved intermediate states.
    Checked against the reference's torch-autograd gradient (golden C3) and against the
    per-gate autograd path on a complex128 circuit."""
    arr = golden.arrays("circuits")
    m = golden.manifest["circuits"]["c3"]
    n, layers = m["n"], m["layers"]
    g = ua.gates
    cn = g.cnot(device="cuda")
    z0 = torch.where((torch.arange(2 ** n, device="cuda") >> (n - 1)) & 1 == 0, 1.0, -1.0)
    for tag in ("c3", "c3b"):
        theta = dev(arr[tag + "_theta"]).requires_grad_(True)
        gates = []
        for l in range(layers):
            for q in range(n):
                t = theta[l, q] if tag == "c3" else theta[:, l, q]
                gates.append(([q], g.exp_y(t[..., 0])))
                gates.append(([q], g.exp_z(t[..., 1])))
            for q in range(n - 1):
                gates.append(([q, q + 1], cn))
        psi = ua.circuit.apply_gates(gates, dev(arr["c3_state"]), assume_unitary=True)
        loss = ua.diag_expectation_value(z0, psi).sum()
        g_theta, = torch.autograd.grad(loss, theta)
        assert_close(host(psi), arr[tag + "_out"], "c64", factor=10, what=tag + " state")
        assert_close(host(g_theta), arr[tag + "_gtheta"], "c64", factor=50, what=tag + " adjoint grad")
    # complex128, state gradient too, vs the per-gate path
    rng = np.random.default_rng(31)
    n = 9
    st0 = dev(rnd_state(rng, n, (3,), "c128"))
    mats = [dev(haar(rng, 2 ** k, "c128")) for k in (1, 2, 1, 3, 2, 1, 2)]
    qls = [[4], [1, 7], [0], [8, 2, 5], [3, 4], [8], [6, 0]]
    w = dev(rnd_c(rng, (3, 2 ** n), "c128"))
    res = []
    for unitary in (True, False):
        st = st0.clone().requires_grad_(True)
        ms = [mm.clone().requires_grad_(True) for mm in mats]
        out = ua.circuit.apply_gates(list(zip(qls, ms)), st, assume_unitary=unitary)
        loss = (out * w.conj()).real.sum()
        res.append(torch.autograd.grad(loss, [st] + ms))
    for a, b in zip(*res):
        assert_close(host(a), host(b), "c128", factor=50, what="adjoint vs per-gate")


# --------------------------------------------------------------------------- sign-mask diagonal gates
def test_multi_cz_family(ua, golden):
    arr = golden.arrays("diag")
    for c in golden.manifest["diag"]:
        k = c["key"]
        st = dev(arr[k + "_state"])
        if c["fn"] == "multi_cz":
            out = ua.simulation.multi_cz(torch.tensor(c["arg"], device="cuda"), st)
        elif c["fn"] == "multi_controlled_z":
            out = ua.simulation.multi_controlled_z(c["arg"], st)
        else:
            out = ua.simulation.multi_controlled_x(st, controls=c["arg"][0], target=c["arg"][1])
        assert tuple(out.shape) == arr[k + "_out"].shape and str(out.dtype) == c["out_dtype"]
        assert_close(host(out), arr[k + "_out"], "c64", what=str(c))
    # equivalence with the dense CZ gate (the reference has no tests for these functions)
    rng = np.random.default_rng(40)
    for dt in ("c64", "c128"):
        st = dev(rnd_state(rng, 11, (3,), dt))
        a = ua.simulation.multi_cz([[2, 9], [0, 10]], st)
        cz = ua.gates.cz(device="cuda", dtype=CD[dt])
        b = ua.simulation.apply_operator(cz, (0, 10), ua.simulation.apply_operator(cz, (2, 9), st))
        assert torch.equal(a, b)                       # sign flips are exact
        ccz = ua.simulation.multi_controlled_z([1, 4, 7], st)
        assert_close(host(ccz), orc.multi_controlled_z([1, 4, 7], host(st)), dt)
    many = [[i, (i + 3) % 12] for i in range(12)] * 7        # > 64 masks -> several launches
    st = dev(rnd_state(rng, 12, (), "c64"))
    assert_close(host(ua.simulation.multi_cz(many, st)), orc.multi_cz(many, host(st)), "c64")
    with pytest.raises(ValueError):
        ua.simulation.multi_cz([[0, 12]], st)
    with pytest.raises(ValueError):
        ua.simulation.multi_cz([[3, 3]], st)
    # gradient: D is its own adjoint
    s2 = dev(rnd_state(rng, 6, (), "c128")).requires_grad_(True)
    w = dev(rnd_c(rng, (64,), "c128"))
    g, = torch.autograd.grad((ua.simulation.multi_cz([[0, 5]], s2) * w.conj()).real.sum(), s2)
    assert_close(host(g), orc.multi_cz([[0, 5]], host(w)), "c128")


def test_cuda_graph_replay(ua):
    """CompiledCircuit.capture_graph: replaying the graph equals running the circuit again."""
    rng = np.random.default_rng(50)
    n = 14
    gates = []
    for _ in range(4):
        for q in range(n):
            gates.append(([q], dev(haar(rng, 2, "c64"))))
        pi = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            gates.append(([pi[j], pi[j + 1]], dev(haar(rng, 4, "c64"))))
    cc = ua.circuit.CompiledCircuit(gates, n, torch.complex64)
    st0 = dev(rnd_state(rng, n, (), "c64"))
    ref = cc.run(cc.run(cc.run(st0)))                 # three applications, out of place
    buf = st0.clone()
    graph = cc.capture_graph(buf)                     # warm-up run applies the circuit once...
    graph.replay()                                    # ...and two replays make three
    graph.replay()
    torch.cuda.synchronize()
    assert_close(host(buf), host(ref), "c64", factor=5)


# --------------------------------------------------------------------------- scatter tail
def _scatter_reference(ref, n, victims):
    """Where ua_apply_fused_pass_scatter puts amplitude i: block = values of the victim bits,
    offset = index with the victim bits squeezed out."""
    idx = torch.arange(1 << n, device=ref.device)
    block = torch.zeros_like(idx)
    for j, v in enumerate(victims):
        block |= ((idx >> v) & 1) << j
    off = idx.clone()
    for v in sorted(victims, reverse=True):
        off = ((off >> (v + 1)) << v) | (off & ((1 << v) - 1))
    out = torch.empty_like(ref)
    out[block * (1 << (n - len(victims))) + off] = ref
    return out


@pytest.mark.parametrize("dt", ["c64", "c128"])
@pytest.mark.parametrize("n,victims", [(16, [9]), (18, [8, 15]), (19, [7, 12, 18]), (15, [14]), (12, [10, 11])])
def test_fused_pass_scatter_matches_permute_and_split(ua, dt, n, victims):
    """The pass that folds a global-qubit exchange into the last fused pass (peer-memory
    stores): with all destinations in local memory it must equal circuit + bit permutation
    (pure data movement after the gates: bit-exact against the in-place circuit)."""
    from unitair_b200 import circuit
    rng = np.random.default_rng(n * 7 + len(victims))
    st = dev(rnd_state(rng, n, (), dt))
    low = circuit.default_geometry(n, CD[dt]).low_bits
    victims = [v for v in victims if v >= low] or [n - 1]
    m = len(victims)
    block_bytes = (1 << (n - m)) * st.element_size()

    def gates_avoiding(avoid, count):
        gl = []
        free = [q for q in range(n) if (n - 1 - q) not in avoid]
        for _ in range(count):
            a, b = rng.choice(free, 2, replace=False)
            gl.append(([int(a), int(b)], dev(haar(rng, 4, dt))))
            gl.append(([int(rng.choice(free))], dev(haar(rng, 2, dt))))
        return gl

    cases = {
        "no circuit (pure scatter copy)": None,
        "last pass reusable": gates_avoiding(set(victims), 6),
        "last pass touches a victim bit (extra copy pass)":
            gates_avoiding(set(), 5) + [([n - 1 - victims[0], (n - 1 - victims[0] + 1) % n], dev(haar(rng, 4, dt)))],
    }
    for name, gl in cases.items():
        cc = circuit.CompiledCircuit(gl, n, CD[dt]) if gl else None
        ref = cc.run(st) if cc is not None else st
        tail = circuit.ScatterTail(cc, n, CD[dt], victims)
        if name == "last pass reusable" and cc.geo.tile_bits <= n - m:
            assert tail.reused, "a last pass that avoids the victim bits must be reused"
        out = torch.full_like(st, float("nan"))
        work = st.clone()
        tail.run(work, [out.data_ptr() + b * block_bytes for b in range(1 << m)],
                 visit_xor=int(rng.integers(0, 1 << m)))
        torch.cuda.synchronize()
        want = _scatter_reference(ref, n, victims)
        if tail.reused:      # same gates, different tile shape: rounding may differ in the last pass
            assert_close(host(out), host(want), dt, what=name)
        else:
            assert torch.equal(torch.view_as_real(out), torch.view_as_real(want)), name


def test_fused_pass_scatter_rejects_bad_arguments(ua):
    from unitair_b200 import _lib as L
    n = 14
    st = torch.zeros(1 << n, dtype=torch.complex64, device="cuda")
    out = torch.zeros_like(st)
    half = (1 << (n - 1)) * 8
    dst = L.ptr_array([out.data_ptr(), out.data_ptr() + half])
    lib = L.lib()
    stream = L.stream_ptr(st.device)

    def call(low, high, victims, total=1 << n):
        return lib.ua_apply_fused_pass_scatter(0, st.data_ptr(), total, n, low, len(high), L.int_array(high) if high else None,
                                               0, None, None, None, None, len(victims), L.int_array(victims), dst, 1, stream)
    assert call(7, [7, 8, 9, 10, 11, 12], [13]) == 0
    assert call(7, [7, 8, 9, 10, 11, 13], [13]) != 0        # scatter bit inside the tile
    assert call(7, [8, 9, 10], [3]) != 0                    # scatter bit among the low bits
    assert call(7, [7, 8, 9, 10, 11, 12, 13], [6]) != 0
    assert call(7, [7, 8], [13], total=2 << n) != 0         # one state only
    torch.cuda.synchronize()


# --------------------------------------------------------------------------- measure
@pytest.mark.parametrize("dt", ["c64", "c128"])
@pytest.mark.parametrize("n", [3, 11, 12, 17, 20])
def test_sample_indices_match_oracle_inverse_cdf(ua, dt, n):
    from unitair_b200.simulation import measurement
    rng = np.random.default_rng(n)
    st = rnd_c(rng, (2 ** n,), dt, scale=0.7)        # not normalised: the sampler normalises
    if n == 17:
        st[: 2 ** 16] = 0                            # empty blocks in front
    u = rng.random(4096)
    u[:3] = [0.0, 0.5, 1.0 - 2 ** -53]
    got = host(measurement.sample_indices(dev(st), len(u), uniforms=torch.from_numpy(u)))
    want = orc.sample_indices(st, u)
    # identical except for draws that land within rounding of a CDF step (different fp64
    # summation order): those may move to a neighbouring index with non-zero probability
    bad = np.nonzero(got != want)[0]
    assert len(bad) <= 2, (len(bad), got[bad][:5], want[bad][:5])
    p = np.abs(st.astype(np.complex128)) ** 2
    cdf = np.cumsum(p)
    for i in bad:
        t = u[i] * cdf[-1]
        assert abs(cdf[min(got[i], want[i])] - t) < 1e-9 * cdf[-1]
    assert np.all(p[got] > 0)


def test_measure_histogram(ua):
    rng = np.random.default_rng(3)
    n = 10
    st = rnd_state(rng, n, (), "c64")
    torch.manual_seed(5)
    num = 400000
    raw = ua.simulation.measure(dev(st), num, raw_output=True)
    assert sum(raw.values()) == num and all(isinstance(k, int) for k in raw)
    p = orc.measurement_probabilities(st)
    counts = np.zeros(2 ** n)
    for k, c in raw.items():
        counts[k] = c
    chi2 = np.sum((counts - num * p) ** 2 / (num * p))
    assert chi2 < 2 ** n + 6 * math.sqrt(2 * 2 ** n), chi2
    hist = ua.simulation.measure(dev(st), 5000)
    assert isinstance(hist, ua.simulation.MeasurementHistogram) and hist.num_qubits == n
    cs = list(hist.histogram.values())
    assert cs == sorted(cs, reverse=True) and sum(cs) == 5000
    assert all(len(k) == n and set(k) <= {"0", "1"} for k in hist.histogram)
    assert hist["0" * n] >= 0
    with pytest.raises(KeyError):
        hist["012"]
    # README Bell state: only |00> and |11> are ever observed
    bell = torch.tensor([2 ** -0.5, 0, 0, 2 ** -0.5], dtype=torch.complex64, device="cuda")
    hb = ua.simulation.measure(bell, 1000)
    assert hb.observed_samples <= {"00", "11"} and hb["00"] + hb["11"] == 1000
    with pytest.raises(RuntimeError, match="CUDA"):
        ua.simulation.measure(torch.zeros(4, dtype=torch.complex64), 3)


# --------------------------------------------------------------------------- host pipeline
def test_host_circuit_stream_matches_apply_gates(ua):
    """HostCircuitStream: jobs whose upload, circuit and download overlap on three streams must
    deliver exactly what circuit.apply_gates gives for each job on its own."""
    n = 18
    rng = np.random.default_rng(21)
    jobs = []
    for j in range(7):
        st = torch.from_numpy(rnd_state(rng, n, (), "c64")).pin_memory()
        gl = []
        for _ in range(12):
            a, b = rng.choice(n, 2, replace=False)
            gl.append(([int(a), int(b)], torch.from_numpy(haar(rng, 4, "c64")).pin_memory()))
            gl.append(([int(rng.integers(n))], torch.from_numpy(haar(rng, 2, "c64")).pin_memory()))
        jobs.append((gl, st, torch.empty(2 ** n, dtype=torch.complex64).pin_memory()))
    hs = ua.HostCircuitStream(n, torch.complex64, "cuda", depth=3)
    for gl, h_in, h_out in jobs:
        hs.submit(gl, h_in, h_out)
    hs.drain()
    for gl, h_in, h_out in jobs:
        ref = ua.circuit.apply_gates([(qs, u.cuda()) for qs, u in gl], h_in.cuda())
        # host operators are merged on the host (no upload, no synchronisation), device operators
        # on the device: the merged 4x4 products differ in the last bit, the states by ~1e-7
        assert_close(h_out.numpy(), host(ref), "c64")
        same = ua.circuit.apply_gates(gl, h_in.cuda())            # host operators here too: bit-exact
        assert torch.equal(torch.view_as_real(h_out), torch.view_as_real(same.cpu()))
    # a compiled plan can be reused, and a second round after drain() works
    cc = None
    outs = [torch.empty(2 ** n, dtype=torch.complex64).pin_memory() for _ in range(4)]
    for o in outs:
        cc = hs.submit(jobs[0][0], jobs[0][1], o, compiled=cc)
    hs.drain()
    for o in outs:
        assert torch.equal(torch.view_as_real(o), torch.view_as_real(jobs[0][2]))
    with pytest.raises(RuntimeError):
        ua.HostCircuitStream(n, torch.complex64, "cpu")
