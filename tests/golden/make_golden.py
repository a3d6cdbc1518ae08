"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONPATH=/root/reference/src python tests/golden/make_golden.py

It imports qcware/qcware-unitair v0.3.0 from /root/reference/src, feeds it seeded
inputs on CPU and stores inputs + outputs (and torch-autograd gradients through the
reference path) as small .npz fixtures next to this script, with a JSON manifest
describing every case.  tests/test_oracle_golden.py pins the numpy oracle against
them; tests/test_gpu_parity.py pins the CUDA engine against them.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
import unitair  # noqa: E402
import unitair.simulation as sim  # noqa: E402
import unitair.gates as gates  # noqa: E402
from unitair import states  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)

CD = {"c64": torch.complex64, "c128": torch.complex128}
RD = {"c64": torch.float32, "c128": torch.float64}


def rnd_c(gen, shape, dtype, scale=1.0):
    re = torch.randn(shape, generator=gen, dtype=torch.float64)
    im = torch.randn(shape, generator=gen, dtype=torch.float64)
    return (scale * torch.complex(re, im)).to(dtype)


def rnd_state(gen, n, batch, dtype):
    s = rnd_c(gen, tuple(batch) + (2 ** n,), torch.complex128)
    s = s / s.abs().pow(2).sum(-1, keepdim=True).sqrt()
    return s.to(dtype)


def haar(gen, dim, dtype):
    z = rnd_c(gen, (dim, dim), torch.complex128)
    q, r = torch.linalg.qr(z)
    d = torch.diagonal(r)
    q = q * (d / d.abs()).unsqueeze(0)
    return q.to(dtype)


def npy(t):
    return t.detach().cpu().numpy()


def make_apply_operator(arrays, manifest):
    gen = torch.Generator().manual_seed(1001)
    cases = []
    idx = 0
    specs = []
    # (n, k, state_batch, op_batch, dtype)
    for dt in ("c64", "c128"):
        for n in (1, 2, 3, 5, 8):
            for k in range(1, min(n, 5) + 1):
                specs.append((n, k, (), (), dt))
        specs += [
            (6, 1, (3,), (), dt), (6, 2, (3,), (3,), dt), (6, 3, (), (4,), dt),
            (5, 2, (2, 3), (2, 3), dt), (5, 1, (3, 2), (2,), dt), (7, 4, (2,), (), dt),
            (4, 2, (3, 1, 2), (), dt), (8, 5, (2,), (2,), dt), (3, 3, (5,), (5,), dt),
            (10, 2, (), (), dt), (10, 1, (4,), (4,), dt), (11, 3, (), (), dt),
        ]
    for (n, k, sb, ob, dt) in specs:
        for rep in range(2 if n <= 5 else 1):
            perm = torch.randperm(n, generator=gen)[:k].tolist()
            op = rnd_c(gen, tuple(ob) + (2 ** k, 2 ** k), CD[dt], scale=2.0)
            st = rnd_state(gen, n, sb, CD[dt])
            out = sim.apply_operator(operator=op, qubits=perm, state=st)
            key = f"ao{idx}"
            arrays[key + "_op"] = npy(op)
            arrays[key + "_state"] = npy(st)
            arrays[key + "_out"] = npy(out)
            cases.append(dict(key=key, n=n, k=k, qubits=perm, dtype=dt,
                              state_batch=list(sb), op_batch=list(ob)))
            idx += 1
    manifest["apply_operator"] = cases


def make_apply_all(arrays, manifest):
    gen = torch.Generator().manual_seed(1002)
    cases = []
    specs = [(1, (), (), "c64"), (3, (), (), "c64"), (7, (), (), "c64"), (7, (), (), "c128"),
             (5, (3,), (), "c64"), (5, (3,), (3,), "c64"), (4, (2, 2), (2, 2), "c128"),
             (6, (), (3,), "c64"), (12, (), (), "c64"), (13, (2,), (2,), "c128")]
    for i, (n, sb, ob, dt) in enumerate(specs):
        op = rnd_c(gen, tuple(ob) + (2, 2), CD[dt])
        st = rnd_state(gen, n, sb, CD[dt])
        out = sim.apply_all_qubits(operator=op, state=st)
        key = f"aa{i}"
        arrays[key + "_op"] = npy(op)
        arrays[key + "_state"] = npy(st)
        arrays[key + "_out"] = npy(out)
        cases.append(dict(key=key, n=n, dtype=dt, state_batch=list(sb), op_batch=list(ob)))
    manifest["apply_all_qubits"] = cases


def make_phase(arrays, manifest):
    gen = torch.Generator().manual_seed(1003)
    cases = []
    # (angle_shape, state_shape, angle_dtype, state_dtype)
    specs = [((32,), (32,), "f32", "c64"), ((32,), (32,), "f64", "c128"),
             ((32,), (3, 32), "f32", "c64"), ((3, 32), (3, 32), "f32", "c64"),
             ((3, 1), (3, 32), "f32", "c64"), ((), (2, 3, 16), "f32", "c64"),
             ((4, 5, 8), (8,), "f64", "c128"), ((5, 1, 8), (5, 3, 8), "f32", "c64"),
             ((16,), (16,), "f64", "c64"), ((16,), (16,), "f32", "c128"),
             ((2, 4096), (2, 4096), "f32", "c64"), ((4096,), (4096,), "f64", "c128"),
             ((8,), (8,), "f32", "f32")]
    fd = {"f32": torch.float32, "f64": torch.float64}
    for i, (ash, ssh, adt, sdt) in enumerate(specs):
        ang = (torch.rand(ash, generator=gen, dtype=torch.float64) * 2 * np.pi).to(fd[adt])
        if sdt in CD:
            st = rnd_c(gen, ssh, CD[sdt])
        else:
            st = torch.randn(ssh, generator=gen, dtype=torch.float64).to(fd[sdt])
        out = sim.apply_phase(ang, st)
        key = f"ph{i}"
        arrays[key + "_angles"] = npy(ang)
        arrays[key + "_state"] = npy(st)
        arrays[key + "_out"] = npy(out)
        cases.append(dict(key=key, angle_shape=list(ash), state_shape=list(ssh),
                          angle_dtype=adt, state_dtype=sdt, out_dtype=str(out.dtype)))
    manifest["apply_phase"] = cases


def make_reductions(arrays, manifest):
    gen = torch.Generator().manual_seed(1004)
    cases = []
    specs = [((8,), "c64"), ((3, 16), "c64"), ((2, 3, 32), "c128"), ((4096,), "c64"),
             ((5, 2048), "c128"), ((1 << 15,), "c64"), ((7, 2), "c64")]
    for i, (shape, dt) in enumerate(specs):
        st = rnd_c(gen, shape, CD[dt])
        st2 = rnd_c(gen, shape, CD[dt])
        diag = torch.randn(shape, generator=gen, dtype=torch.float64).to(RD[dt])
        diag_b = torch.randn(shape[-1:], generator=gen, dtype=torch.float64).to(RD[dt])
        key = f"rd{i}"
        arrays[key + "_state"] = npy(st)
        arrays[key + "_state2"] = npy(st2)
        arrays[key + "_diag"] = npy(diag)
        arrays[key + "_diagb"] = npy(diag_b)
        arrays[key + "_abs2"] = npy(states.abs_squared(st))
        arrays[key + "_norm2"] = npy(states.norm_squared(st))
        arrays[key + "_dexp"] = npy(states.diag_expectation_value(diag, st))
        arrays[key + "_dexpb"] = npy(states.diag_expectation_value(diag_b, st))
        arrays[key + "_inner"] = npy(states.inner_product(st, st2))
        cases.append(dict(key=key, shape=list(shape), dtype=dt))
    manifest["reductions"] = cases


def make_grads(arrays, manifest):
    """torch-autograd gradients through the unmodified reference path."""
    gen = torch.Generator().manual_seed(1005)
    cases = []
    specs = [(4, 1, (), (), "c128"), (4, 2, (), (), "c128"), (5, 3, (), (), "c128"),
             (5, 1, (3,), (), "c128"), (5, 2, (3,), (3,), "c128"), (5, 1, (), (4,), "c128"),
             (6, 2, (2, 3), (2, 3), "c64"), (6, 1, (4,), (), "c64"), (7, 4, (2,), (), "c128"),
             (6, 5, (), (), "c128"), (9, 2, (), (), "c64"), (5, 1, (3, 2), (2,), "c128")]
    for i, (n, k, sb, ob, dt) in enumerate(specs):
        perm = torch.randperm(n, generator=gen)[:k].tolist()
        op = rnd_c(gen, tuple(ob) + (2 ** k, 2 ** k), CD[dt]).requires_grad_(True)
        st = rnd_state(gen, n, sb, CD[dt]).requires_grad_(True)
        out = sim.apply_operator(operator=op, qubits=perm, state=st)
        w = rnd_c(gen, tuple(out.shape), CD[dt])
        loss = (out * w.conj()).real.sum() + (out.abs() ** 2).sum() * 0.5
        g_out, = torch.autograd.grad(loss, out, retain_graph=True)
        g_op, g_st = torch.autograd.grad(loss, (op, st))
        key = f"gr{i}"
        arrays[key + "_op"] = npy(op)
        arrays[key + "_state"] = npy(st)
        arrays[key + "_w"] = npy(w)
        arrays[key + "_gout"] = npy(g_out)
        arrays[key + "_gop"] = npy(g_op)
        arrays[key + "_gstate"] = npy(g_st)
        cases.append(dict(key=key, n=n, k=k, qubits=perm, dtype=dt,
                          state_batch=list(sb), op_batch=list(ob)))
    manifest["grads_apply_operator"] = cases

    # phase + diag expectation gradients
    pcases = []
    pspecs = [((16,), (16,), "c128"), ((16,), (3, 16), "c128"), ((3, 1), (3, 16), "c128"),
              ((), (2, 8), "c128"), ((2, 32), (2, 32), "c64")]
    for i, (ash, ssh, dt) in enumerate(pspecs):
        ang = (torch.rand(ash, generator=gen, dtype=torch.float64) * 6.28).to(RD[dt]).requires_grad_(True)
        st = rnd_c(gen, ssh, CD[dt]).requires_grad_(True)
        diag = torch.randn(ssh, generator=gen, dtype=torch.float64).to(RD[dt])
        out = sim.apply_phase(ang, st)
        w = rnd_c(gen, tuple(out.shape), CD[dt])
        loss = (out * w.conj()).real.sum() + states.diag_expectation_value(diag, out).sum()
        g_ang, g_st = torch.autograd.grad(loss, (ang, st))
        key = f"gp{i}"
        arrays[key + "_angles"] = npy(ang)
        arrays[key + "_state"] = npy(st)
        arrays[key + "_w"] = npy(w)
        arrays[key + "_diag"] = npy(diag)
        arrays[key + "_gangles"] = npy(g_ang)
        arrays[key + "_gstate"] = npy(g_st)
        arrays[key + "_loss"] = npy(loss)
        pcases.append(dict(key=key, angle_shape=list(ash), state_shape=list(ssh), dtype=dt))
    manifest["grads_phase_expectation"] = pcases


def make_circuits(arrays, manifest):
    """Scaled-down versions of BASELINE.json's configs C1-C4."""
    gen = torch.Generator().manual_seed(1006)
    out_cases = {}

    # C1: n=10 (batch reduced to 8): H on every qubit, exp_x(theta_q), CNOT chain
    n, B = 10, 8
    st = rnd_state(gen, n, (B,), torch.complex64)
    theta = (torch.rand(n, generator=gen, dtype=torch.float64) * 2 * np.pi).to(torch.float32)
    psi = st
    h = gates.hadamard()
    for q in range(n):
        psi = sim.apply_operator(h, (q,), psi)
    for q in range(n):
        psi = sim.apply_operator(gates.exp_x(theta[q]), (q,), psi)
    cn = gates.cnot()
    for q in range(n - 1):
        psi = sim.apply_operator(cn, (q, q + 1), psi)
    arrays["c1_state"] = npy(st)
    arrays["c1_theta"] = npy(theta)
    arrays["c1_out"] = npy(psi)
    out_cases["c1"] = dict(n=n, batch=B)

    # C2: n=12, 6 layers, Haar U(2) on every qubit then Haar U(4) on random ordered pairs
    n, layers = 12, 6
    st = rnd_state(gen, n, (), torch.complex64)
    psi = st
    glist = []
    gi = 0
    for l in range(layers):
        for q in range(n):
            u = haar(gen, 2, torch.complex64)
            arrays[f"c2_g{gi}"] = npy(u)
            glist.append(dict(g=gi, qubits=[q])); gi += 1
            psi = sim.apply_operator(u, (q,), psi)
        pi = torch.randperm(n, generator=gen).tolist()
        for j in range(0, n - 1, 2):
            u = haar(gen, 4, torch.complex64)
            arrays[f"c2_g{gi}"] = npy(u)
            glist.append(dict(g=gi, qubits=[pi[j], pi[j + 1]])); gi += 1
            psi = sim.apply_operator(u, (pi[j], pi[j + 1]), psi)
    arrays["c2_state"] = npy(st)
    arrays["c2_out"] = npy(psi)
    out_cases["c2"] = dict(n=n, layers=layers, gates=glist)

    # C3: n=6, B=4, 3 layers ry/rz + CNOT ladder, loss = sum_b <Z0>, grad wrt theta
    n, B, layers = 6, 4, 3
    st = rnd_state(gen, n, (B,), torch.complex64)
    theta = (torch.rand(layers, n, 2, generator=gen, dtype=torch.float64) * 2 * np.pi).to(torch.float32)
    theta.requires_grad_(True)
    z0 = torch.where((torch.arange(2 ** n) >> (n - 1)) & 1 == 0, 1.0, -1.0).to(torch.float32)
    psi = st
    for l in range(layers):
        for q in range(n):
            psi = sim.apply_operator(gates.exp_y(theta[l, q, 0]), (q,), psi)
            psi = sim.apply_operator(gates.exp_z(theta[l, q, 1]), (q,), psi)
        for q in range(n - 1):
            psi = sim.apply_operator(cn, (q, q + 1), psi)
    loss = states.diag_expectation_value(z0, psi).sum()
    g_theta, = torch.autograd.grad(loss, theta)
    arrays["c3_state"] = npy(st)
    arrays["c3_theta"] = npy(theta)
    arrays["c3_out"] = npy(psi)
    arrays["c3_loss"] = npy(loss)
    arrays["c3_gtheta"] = npy(g_theta)
    out_cases["c3"] = dict(n=n, batch=B, layers=layers)

    # C3 variant B: per-entry theta (batched gates)
    theta_b = (torch.rand(B, layers, n, 2, generator=gen, dtype=torch.float64) * 2 * np.pi).to(torch.float32)
    theta_b.requires_grad_(True)
    psi = st
    for l in range(layers):
        for q in range(n):
            psi = sim.apply_operator(gates.exp_y(theta_b[:, l, q, 0]), (q,), psi)
            psi = sim.apply_operator(gates.exp_z(theta_b[:, l, q, 1]), (q,), psi)
        for q in range(n - 1):
            psi = sim.apply_operator(cn, (q, q + 1), psi)
    loss = states.diag_expectation_value(z0, psi).sum()
    g_theta_b, = torch.autograd.grad(loss, theta_b)
    arrays["c3b_theta"] = npy(theta_b)
    arrays["c3b_out"] = npy(psi)
    arrays["c3b_loss"] = npy(loss)
    arrays["c3b_gtheta"] = npy(g_theta_b)

    # C4: n=10 c128, 2 layers of two Haar U(32) blocks on ordered 5-tuples + f64 phase layer
    n, layers = 10, 2
    st = rnd_state(gen, n, (), torch.complex128)
    psi = st
    blocks = []
    bi = 0
    for l in range(layers):
        pi = torch.randperm(n, generator=gen).tolist()
        for j in range(0, n, 5):
            u = haar(gen, 32, torch.complex128)
            arrays[f"c4_g{bi}"] = npy(u)
            blocks.append(dict(g=bi, qubits=pi[j:j + 5])); bi += 1
            psi = sim.apply_operator(u, pi[j:j + 5], psi)
        ang = torch.rand(2 ** n, generator=gen, dtype=torch.float64) * 2 * np.pi
        arrays[f"c4_ang{l}"] = npy(ang)
        psi = sim.apply_phase(ang, psi)
    arrays["c4_state"] = npy(st)
    arrays["c4_out"] = npy(psi)
    out_cases["c4"] = dict(n=n, layers=layers, blocks=blocks)
    manifest["circuits"] = out_cases


def make_diag(arrays, manifest):
    """multi_cz / multi_controlled_z / multi_controlled_x / swap / permute."""
    gen = torch.Generator().manual_seed(1007)
    cases = []
    specs = [("multi_cz", 4, (), [[0, 1]]), ("multi_cz", 6, (3,), [[0, 5], [2, 3], [5, 1]]),
             ("multi_cz", 5, (), [3, 1]), ("multi_cz", 9, (2, 2), [[8, 0], [4, 7], [0, 4], [1, 2]]),
             ("multi_controlled_z", 5, (), [0, 2, 4]), ("multi_controlled_z", 7, (3,), [6, 1]),
             ("multi_controlled_z", 6, (), [5, 4, 3, 2, 1, 0]), ("multi_controlled_z", 3, (), [1]),
             ("multi_controlled_x", 5, (), [[0, 3], 2]), ("multi_controlled_x", 6, (2,), [[5], 0]),
             ("multi_controlled_x", 4, (), [[1, 2, 3], 0])]
    for i, (fn, n, batch, arg) in enumerate(specs):
        st = rnd_state(gen, n, batch, torch.complex64)
        if fn == "multi_cz":
            out = sim.multi_cz(torch.tensor(arg), st)
        elif fn == "multi_controlled_z":
            out = sim.multi_controlled_z(arg, st)
        else:
            out = sim.multi_controlled_x(st, controls=arg[0], target=arg[1])
        key = f"dg{i}"
        arrays[key + "_state"] = npy(st)
        arrays[key + "_out"] = npy(out)
        cases.append(dict(key=key, fn=fn, n=n, batch=list(batch), arg=arg, out_dtype=str(out.dtype)))
    manifest["diag"] = cases


def make_layout(arrays, manifest):
    """Qubit permutations and the tensor-layout fronts (operations.py:189-233, 258-329, 416-654),
    apply_operator_tensor / apply_all_qubits_tensor (:151-186, :369-413) and the probability
    vector measure() samples from (measurement.py:41-43)."""
    from unitair.simulation import operations as ops
    import warnings
    gen = torch.Generator().manual_seed(1008)
    cases = []

    def add(fn, n, batch, arg, st, out, extra=None):
        key = f"lo{len(cases)}"
        arrays[key + "_state"] = npy(st)
        arrays[key + "_out"] = npy(out.contiguous())
        if extra:
            for k, v in extra.items():
                arrays[key + "_" + k] = npy(v)
        cases.append(dict(key=key, fn=fn, n=n, batch=list(batch), arg=arg))

    for n, batch, steps in [(1, (), 1), (4, (), 1), (4, (), 3), (5, (2,), 2), (6, (), -1), (6, (3, 2), 7),
                            (7, (), 0), (8, (), 5), (11, (), 4)]:
        st = rnd_state(gen, n, batch, torch.complex64)
        add("roll_qubits", n, batch, steps, st, sim.operations.roll_qubits(st, num_steps=steps))
    for n, batch, steps in [(5, (), 2), (6, (2,), 5)]:
        st = rnd_state(gen, n, batch, torch.complex128)
        t = states.to_tensor_layout(st)
        add("roll_qubits_tensor", n, batch, steps, st, ops.roll_qubits_tensor(t, n, steps))
    for n, batch, pair in [(5, (), (0, 4)), (6, (3,), (2, 2)), (7, (), (5, 1))]:
        st = rnd_state(gen, n, batch, torch.complex64)
        t = states.to_tensor_layout(st)
        add("swap_tensor", n, batch, list(pair), st, ops.swap_tensor(t, pair, n))
    for n, batch in [(5, ()), (7, (2,)), (9, ())]:
        st = rnd_state(gen, n, batch, torch.complex64)
        perm = torch.randperm(n, generator=gen).tolist()
        t = states.to_tensor_layout(st)
        add("permute_qubits_tensor", n, batch, perm, st, ops.permute_qubits_tensor(perm, t, n, contiguous_output=True))
    for n, k, sb, ob in [(5, 2, (), ()), (6, 3, (3,), ()), (6, 1, (3,), (3,)), (5, 2, (), (4,))]:
        st = rnd_state(gen, n, sb, torch.complex64)
        op = rnd_c(gen, tuple(ob) + (2 ** k, 2 ** k), torch.complex64)
        t = states.to_tensor_layout(st)
        add("act_first_qubits_tensor", n, sb, k, st, ops.act_first_qubits_tensor(op, t, n, k), {"op": op})
    for n, k, sb, ob in [(5, 2, (), ()), (6, 3, (2,), ()), (6, 2, (2,), (2,))]:
        st = rnd_state(gen, n, sb, torch.complex128)
        op = rnd_c(gen, tuple(ob) + (2 ** k, 2 ** k), torch.complex128)
        qs = torch.randperm(n, generator=gen)[:k].tolist()
        t = states.to_tensor_layout(st)
        add("apply_operator_tensor", n, sb, qs, st, ops.apply_operator_tensor(op, qs, t, n), {"op": op})
    for n, sb, ob in [(4, (), ()), (6, (3,), ()), (5, (2,), (2,))]:
        st = rnd_state(gen, n, sb, torch.complex64)
        op = rnd_c(gen, tuple(ob) + (2, 2), torch.complex64)
        t = states.to_tensor_layout(st)
        add("apply_all_qubits_tensor", n, sb, None, st, ops.apply_all_qubits_tensor(op, t, n), {"op": op})
    for n, sb, ob, qs in [(5, (), (), [0, 3, 3, 1]), (6, (2,), (2,), [5, 0, 5]), (4, (), (), [2]), (7, (3,), (), [6, 1, 0, 1, 6])]:
        st = rnd_state(gen, n, sb, torch.complex64)
        oplist = [rnd_c(gen, tuple(ob) + (2, 2), torch.complex64) for _ in qs]
        out = sim.apply_to_qubits(oplist, qs, st)
        t = states.to_tensor_layout(st)
        out_t = ops.apply_to_qubits_tensor(oplist, qs, t, n)
        assert torch.equal(states.to_vector_layout(out_t.contiguous(), n), out)
        add("apply_to_qubits", n, sb, qs, st, out, {"ops": torch.stack(oplist)})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n, sb in [(1, ()), (4, ()), (6, (3,))]:
            st = rnd_state(gen, n, sb, torch.complex64)
            op = rnd_c(gen, (2, 2), torch.complex64)
            add("act_last_qubit", n, sb, None, st, sim.act_last_qubit(op, st), {"op": op})
    # the distribution measure() samples from: Categorical(probs=abs_squared(state)).probs
    for n, dt in [(3, torch.complex64), (8, torch.complex64), (10, torch.complex128)]:
        st = rnd_c(gen, (2 ** n,), dt, scale=0.7)           # not normalised: Categorical normalises
        probs = torch.distributions.Categorical(probs=states.abs_squared(st)).probs
        add("measure_probs", n, (), None, st, probs)
    manifest["layout"] = cases


def seeded_state(seed, n, batch=(), dtype=np.complex64):
    """numpy-seeded random normalised state: the tests regenerate it instead of storing it."""
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(tuple(batch) + (2 ** n,)) + 1j * rng.standard_normal(tuple(batch) + (2 ** n,))
    s /= np.linalg.norm(s, axis=-1, keepdims=True)
    return s.astype(dtype)


def seeded_haar(rng, dim, dtype):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return (q * (d / np.abs(d))).astype(dtype)


def make_fullsize(arrays, manifest):
    """BASELINE.json's configs at (or near) full size.  Inputs come from numpy seeds (the tests
    regenerate them); only compact outputs of the reference are stored: sampled amplitudes,
    per-entry expectation values, the loss and the theta gradient."""
    torch.set_num_threads(8)
    out_cases = {}
    cn = gates.cnot()

    # C2: 24 qubits, 2 layers of the recipe (Haar U(2) on every qubit + Haar U(4) on random pairs)
    n, layers, seed = 24, 2, 2024
    rng = np.random.default_rng(seed)
    psi = torch.from_numpy(seeded_state(seed + 1, n))
    glist = []
    for l in range(layers):
        for q in range(n):
            u = seeded_haar(rng, 2, np.complex64)
            glist.append(([q], u))
        perm = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            glist.append(([perm[j], perm[j + 1]], seeded_haar(rng, 4, np.complex64)))
    for qs, u in glist:
        psi = sim.apply_operator(torch.from_numpy(u), qs, psi)
    idx = np.random.default_rng(seed + 2).choice(2 ** n, 4096, replace=False)
    arrays["c2_idx"] = idx.astype(np.int64)
    arrays["c2_amps"] = npy(psi)[idx]
    arrays["c2_norm2"] = npy(states.norm_squared(psi))
    out_cases["c2"] = dict(n=n, layers=layers, seed=seed, gates=len(glist))
    del psi

    # C4: 20 qubits complex128, 2 layers of four Haar U(32) blocks on ordered 5-tuples + f64 phase layer
    n, layers, seed = 20, 2, 4040
    rng = np.random.default_rng(seed)
    psi = torch.from_numpy(seeded_state(seed + 1, n, dtype=np.complex128))
    for l in range(layers):
        perm = rng.permutation(n).tolist()
        for j in range(0, n, 5):
            u = seeded_haar(rng, 32, np.complex128)
            psi = sim.apply_operator(torch.from_numpy(u), perm[j:j + 5], psi)
        ang = rng.random(2 ** n) * 2 * np.pi
        psi = sim.apply_phase(torch.from_numpy(ang), psi)
    idx = np.random.default_rng(seed + 2).choice(2 ** n, 4096, replace=False)
    arrays["c4_idx"] = idx.astype(np.int64)
    arrays["c4_amps"] = npy(psi)[idx]
    out_cases["c4"] = dict(n=n, layers=layers, seed=seed)
    del psi

    # C3: 16 qubits, 20 layers (ry, rz on every qubit + CNOT ladder = 940 gates), batch 64, loss =
    # sum_b <Z_0>, gradient w.r.t. the shared theta by torch autograd through the reference, in
    # batch chunks of 8 (the tape holds one state per gate)
    n, B, layers, seed = 16, 64, 20, 3030
    theta_np = (np.random.default_rng(seed).random((layers, n, 2)) * 2 * np.pi).astype(np.float32)
    st_np = seeded_state(seed + 1, n, (B,))
    z0_64 = torch.where((torch.arange(2 ** n) >> (n - 1)) & 1 == 0, 1.0, -1.0).to(torch.float64)
    # float32 = the reference path under test; float64 (same inputs, promoted) = the yardstick
    # that says how much of a float32 difference is rounding of an ill-conditioned sum
    for tag, cdt, rdt in (("", torch.complex64, torch.float32), ("64", torch.complex128, torch.float64)):
        st = torch.from_numpy(st_np).to(cdt)
        theta = torch.from_numpy(theta_np).to(rdt).requires_grad_(True)
        z0 = z0_64.to(rdt)
        cn = gates.cnot().to(cdt)
        loss_total = 0.0
        chunk_grads, ez = [], []
        for c0 in range(0, B, 8):
            psi = st[c0:c0 + 8]
            for l in range(layers):
                for q in range(n):
                    psi = sim.apply_operator(gates.exp_y(theta[l, q, 0]).to(cdt), (q,), psi)
                    psi = sim.apply_operator(gates.exp_z(theta[l, q, 1]).to(cdt), (q,), psi)
                for q in range(n - 1):
                    psi = sim.apply_operator(cn, (q, q + 1), psi)
            e = states.diag_expectation_value(z0, psi)
            loss = e.sum()
            g, = torch.autograd.grad(loss, theta)
            chunk_grads.append(npy(g))
            loss_total += float(loss.detach())
            ez.append(npy(e))
            print("  c3", tag or "32", "chunk", c0, float(loss.detach()), flush=True)
        arrays["c3_ez" + tag] = np.concatenate(ez)
        arrays["c3_loss" + tag] = np.float64(loss_total)
        arrays["c3_gtheta_chunks" + tag] = np.stack(chunk_grads)          # (8, layers, n, 2): per chunk of 8 states
        arrays["c3_gtheta" + tag] = np.stack(chunk_grads).astype(np.float64).sum(0).astype(npy(theta).dtype)
    arrays["c3_theta"] = theta_np
    out_cases["c3"] = dict(n=n, batch=B, layers=layers, seed=seed, chunk=8)
    manifest["fullsize"] = out_cases


ALL = [("apply_operator", make_apply_operator), ("apply_all", make_apply_all),
       ("phase", make_phase), ("reductions", make_reductions),
       ("grads", make_grads), ("circuits", make_circuits), ("diag", make_diag),
       ("layout", make_layout), ("fullsize", make_fullsize)]


def main():
    """make_golden.py [name ...]  -- regenerate all fixtures, or only the named ones."""
    only = set(sys.argv[1:])
    mpath = os.path.join(HERE, "manifest.json")
    manifest = {}
    if only and os.path.exists(mpath):
        with open(mpath) as f:
            manifest = json.load(f)
    manifest["reference"] = "qcware/qcware-unitair v0.3.0 (/root/reference/src)"
    manifest["torch"] = torch.__version__
    for name, fn in ALL:
        if only and name not in only:
            continue
        arrays = {}
        fn(arrays, manifest)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **arrays)
        print(name, len(arrays), "arrays")
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
