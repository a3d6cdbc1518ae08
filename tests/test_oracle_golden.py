"""Pin the numpy oracle (oracle/unitair_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py) and the reference docs' known answers.
CPU only."""
import numpy as np
import pytest

from oracle import unitair_oracle as orc
from conftest import assert_close


def test_apply_operator_golden(golden):
    arr = golden.arrays("apply_operator")
    cases = golden.manifest["apply_operator"]
    assert len(cases) > 60
    for c in cases:
        k = c["key"]
        out = orc.apply_operator(arr[k + "_op"], c["qubits"], arr[k + "_state"])
        assert out.shape == arr[k + "_out"].shape and out.dtype == arr[k + "_out"].dtype
        assert_close(out, arr[k + "_out"], c["dtype"], what=str(c))


def test_apply_all_qubits_golden(golden):
    arr = golden.arrays("apply_all")
    for c in golden.manifest["apply_all_qubits"]:
        k = c["key"]
        out = orc.apply_all_qubits(arr[k + "_op"], arr[k + "_state"])
        assert out.shape == arr[k + "_out"].shape
        assert_close(out, arr[k + "_out"], c["dtype"], factor=3, what=str(c))


def test_apply_phase_golden(golden):
    arr = golden.arrays("phase")
    for c in golden.manifest["apply_phase"]:
        k = c["key"]
        out = orc.apply_phase(arr[k + "_angles"], arr[k + "_state"])
        ref = arr[k + "_out"]
        assert out.shape == ref.shape and out.dtype == ref.dtype, c
        # the factors carry the ANGLES' precision (f32 angles on a c128 state -> f32 accuracy)
        key = "c64" if c["angle_dtype"] == "f32" else ref.dtype
        assert_close(out, ref, key, what=str(c))


def test_reductions_golden(golden):
    arr = golden.arrays("reductions")
    for c in golden.manifest["reductions"]:
        k = c["key"]
        st, st2 = arr[k + "_state"], arr[k + "_state2"]
        assert_close(orc.abs_squared(st), arr[k + "_abs2"], c["dtype"])
        assert_close(orc.norm_squared(st), arr[k + "_norm2"], c["dtype"])
        assert_close(orc.diag_expectation_value(arr[k + "_diag"], st), arr[k + "_dexp"], c["dtype"], factor=5)
        assert_close(orc.diag_expectation_value(arr[k + "_diagb"], st), arr[k + "_dexpb"], c["dtype"], factor=5)
        assert_close(orc.inner_product(st, st2), arr[k + "_inner"], c["dtype"], factor=5)


def test_grads_golden(golden):
    arr = golden.arrays("grads")
    for c in golden.manifest["grads_apply_operator"]:
        k = c["key"]
        gu, gs = orc.apply_operator_grads(arr[k + "_op"], c["qubits"], arr[k + "_state"], arr[k + "_gout"])
        assert gu.shape == arr[k + "_gop"].shape, c
        assert gs.shape == arr[k + "_gstate"].shape, c
        assert_close(gu, arr[k + "_gop"], c["dtype"], factor=5, what="grad_op " + str(c))
        assert_close(gs, arr[k + "_gstate"], c["dtype"], factor=5, what="grad_state " + str(c))


def test_circuit_c2_c4_golden(golden):
    arr = golden.arrays("circuits")
    m = golden.manifest["circuits"]
    psi = arr["c2_state"]
    for g in m["c2"]["gates"]:
        psi = orc.apply_operator(arr[f"c2_g{g['g']}"], g["qubits"], psi)
    assert_close(psi, arr["c2_out"], "c64", factor=10)
    psi = arr["c4_state"]
    nb = len(m["c4"]["blocks"]) // m["c4"]["layers"]
    for l in range(m["c4"]["layers"]):
        for g in m["c4"]["blocks"][l * nb:(l + 1) * nb]:
            psi = orc.apply_operator(arr[f"c4_g{g['g']}"], g["qubits"], psi)
        psi = orc.apply_phase(arr[f"c4_ang{l}"], psi)
    assert_close(psi, arr["c4_out"], "c128", factor=10)


def test_diag_golden(golden):
    arr = golden.arrays("diag")
    for c in golden.manifest["diag"]:
        k = c["key"]
        st = arr[k + "_state"]
        if c["fn"] == "multi_cz":
            out = orc.multi_cz(c["arg"], st)
        elif c["fn"] == "multi_controlled_z":
            out = orc.multi_controlled_z(c["arg"], st)
        else:
            out = orc.multi_controlled_x(st, c["arg"][0], c["arg"][1])
        assert out.shape == arr[k + "_out"].shape
        assert_close(out, arr[k + "_out"], "c64", what=str(c))


def test_docs_known_answers():
    """Known-answer vectors in the reference docs."""
    q = np.array([[1, 5 - 1j], [5 + 1j, -1]], dtype=np.complex64)   # first_example.rst:74-75
    ket0 = np.array([1, 0], dtype=np.complex64)
    ket1 = np.array([0, 1], dtype=np.complex64)
    np.testing.assert_allclose(orc.apply_operator(q, (0,), ket0), [1, 5 + 1j])
    batch = np.stack([ket0, ket1])                                   # :123-124, 187-189
    np.testing.assert_allclose(orc.apply_operator(q, (0,), batch), [[1, 5 + 1j], [5 - 1j, -1]])
    h = np.array([[1, 1], [1, -1]], dtype=np.complex64) * 2 ** -0.5   # README.rst:165-166
    np.testing.assert_allclose(orc.apply_operator(h, (0,), ket0), [0.70710678, 0.70710678], rtol=1e-6)
    cnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex64)
    s = np.zeros(4, np.complex64); s[0] = 1                          # README.rst:223-224 (Bell state)
    s = orc.apply_operator(h, (0,), s)
    s = orc.apply_operator(cnot, (0, 1), s)
    np.testing.assert_allclose(s, [0.70710678, 0, 0, 0.70710678], rtol=1e-6, atol=1e-7)
    # X on qubit 0 of |000> -> index 4 (qubit 0 is the most significant bit; conversions.py:43-45)
    x = np.array([[0, 1], [1, 0]], dtype=np.complex64)
    s = np.zeros(8, np.complex64); s[0] = 1
    assert np.argmax(np.abs(orc.apply_operator(x, (0,), s))) == 4
    # qubit order matters: CNOT on (1,0) != CNOT on (0,1)  (operations.py:82-86)
    s = np.zeros(4, np.complex64); s[1] = 1   # |01>
    assert np.argmax(np.abs(orc.apply_operator(cnot, (1, 0), s))) == 3
    assert np.argmax(np.abs(orc.apply_operator(cnot, (0, 1), s))) == 1


def test_oracle_errors():
    st = np.zeros(8, np.complex64)
    op = np.eye(2, dtype=np.complex64)
    with pytest.raises(ValueError):
        orc.apply_operator(op, (3,), st)
    with pytest.raises(ValueError):
        orc.apply_operator(op, (-1,), st)
    with pytest.raises(ValueError):
        orc.apply_operator(op, (0, 1), st)
    with pytest.raises(ValueError):
        orc.apply_operator(np.eye(4, dtype=np.complex64), (1, 1), st)
    with pytest.raises(orc.StateShapeError):
        orc.apply_operator(op, (0,), np.zeros(6, np.complex64))
    with pytest.raises(RuntimeError):
        orc.apply_operator(np.zeros((3, 3), np.complex64), (0,), st)
    with pytest.raises(ValueError):
        orc.apply_all_qubits(np.eye(4, dtype=np.complex64), st)


def test_oracle_sampling_distribution_matches_reference_probabilities():
    """measure(): the oracle's inverse CDF samples the distribution the reference hands to
    torch.distributions.Categorical (abs_squared, normalised), measurement.py:41-42."""
    rng = np.random.default_rng(12)
    n = 6
    st = (rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)).astype(np.complex64) * 0.3
    p = orc.measurement_probabilities(st)
    assert abs(p.sum() - 1.0) < 1e-12
    assert np.allclose(p * np.sum(np.abs(st.astype(np.complex128)) ** 2), np.abs(st.astype(np.complex128)) ** 2, rtol=1e-6)
    u = rng.random(200000)
    idx = orc.sample_indices(st, u)
    counts = np.bincount(idx, minlength=2 ** n)
    chi2 = np.sum((counts - len(u) * p) ** 2 / (len(u) * p))
    assert chi2 < 2 ** n + 6 * np.sqrt(2 * 2 ** n), chi2          # mean 63, sigma ~11
    # deterministic corner cases: a basis state, and u = 0 / u -> 1
    e3 = np.zeros(8, dtype=np.complex64)
    e3[3] = 1
    assert set(orc.sample_indices(e3, rng.random(50)).tolist()) == {3}
    assert orc.sample_indices(st, [0.0])[0] == 0
    assert orc.sample_indices(st, [1.0 - 1e-16])[0] == 2 ** n - 1


def test_layout_golden(golden):
    """Qubit permutations, tensor-layout fronts, apply_to_qubits, act_last_qubit and the measure()
    probability vector, all produced by the unmodified reference (make_golden.py: make_layout)."""
    arr = golden.arrays("layout")
    seen = set()
    for c in golden.manifest["layout"]:
        k, fn, n = c["key"], c["fn"], c["n"]
        st, ref = arr[k + "_state"], arr[k + "_out"]
        seen.add(fn)
        flat = ref.reshape(st.shape) if ref.size == st.size and fn != "measure_probs" else ref
        if fn in ("roll_qubits", "roll_qubits_tensor"):
            out = orc.roll_qubits(st, c["arg"])
            assert np.array_equal(out, flat), c
        elif fn == "swap_tensor":
            assert np.array_equal(orc.swap(st, tuple(c["arg"])), flat), c
        elif fn == "permute_qubits_tensor":
            assert np.array_equal(orc.permute_qubits(c["arg"], st), flat), c
        elif fn == "act_first_qubits_tensor":
            out = orc.apply_operator(arr[k + "_op"], list(range(c["arg"])), st)
            assert_close(out, ref.reshape(out.shape), ref.dtype, what=str(c))
        elif fn == "apply_operator_tensor":
            out = orc.apply_operator(arr[k + "_op"], c["arg"], st)
            assert_close(out, ref.reshape(out.shape), ref.dtype, what=str(c))
        elif fn == "apply_all_qubits_tensor":
            out = orc.apply_all_qubits(arr[k + "_op"], st)
            assert_close(out, ref.reshape(out.shape), ref.dtype, factor=3, what=str(c))
        elif fn == "apply_to_qubits":
            out = orc.apply_to_qubits(list(arr[k + "_ops"]), c["arg"], st)
            assert_close(out, flat, ref.dtype, factor=3, what=str(c))
        elif fn == "act_last_qubit":
            assert_close(orc.act_last_qubit(arr[k + "_op"], st), flat, ref.dtype, what=str(c))
        elif fn == "measure_probs":
            p = orc.measurement_probabilities(st)
            assert p.shape == ref.shape
            assert np.allclose(p, ref, rtol=(1e-5 if st.dtype == np.complex64 else 1e-12), atol=0), c
    assert seen == {"roll_qubits", "roll_qubits_tensor", "swap_tensor", "permute_qubits_tensor",
                    "act_first_qubits_tensor", "apply_operator_tensor", "apply_all_qubits_tensor",
                    "apply_to_qubits", "act_last_qubit", "measure_probs"}
