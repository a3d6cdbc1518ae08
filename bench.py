#!/usr/bin/env python
"""Headline benchmark: gate-amplitude updates/s of a random 1-/2-qubit circuit on a
30-qubit complex64 state per B200 (weak scaling: 30 + log2(N) qubits sharded over N GPUs by
the high-order qubits), with the HBM roofline of the dominant kernel and the reference's
CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" applies the whole `--layers`-layer circuit (SURVEY.md 8d recipe "C2": a Haar
U(2) on every qubit, then Haar U(4) on the ordered pairs of a random qubit permutation) to
the state resident in HBM.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "qcware-unitair_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "gate_amplitude_updates_per_s"
UNIT = "updates/s"


# --------------------------------------------------------------------------- workload
def haar_unitary(rng, dim):
    z = rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))
    q, r = np.linalg.qr(z)
    d = np.diagonal(r)
    return q * (d / np.abs(d))


def random_circuit(n, layers, seed):
    """[(qubits, 2^k x 2^k complex128 ndarray)] -- the C2 layer recipe."""
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(layers):
        for q in range(n):
            gates.append(([q], haar_unitary(rng, 2)))
        perm = rng.permutation(n).tolist()
        for j in range(0, n - 1, 2):
            gates.append(([perm[j], perm[j + 1]], haar_unitary(rng, 4)))
    return gates


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[0]))
                    out["sm_max_mhz"] = float(parts[1])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), parts[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(kernel):
    """dram bytes per launch from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# --------------------------------------------------------------------------- reference / CPU arm
def load_reference():
    """The unmodified reference if it travelled (baseline/_ref), else the numpy oracle port."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "unitair")):
        sys.path.insert(0, ref_dir)
        try:
            import unitair.simulation as sim  # noqa
            return "reference", sim
        except Exception:
            sys.path.remove(ref_dir)
    from oracle import unitair_oracle as orc
    return "port", orc


def cpu_layer_rate(n, layers, seed, reps, warm):
    """gate-amplitude updates/s of the reference's CPU path (simulation.apply_operator once per
    gate, /root/reference/src/unitair/simulation/operations.py:45) on an n-qubit sample of the
    bench recipe, with every host core torch can use."""
    import torch
    kind, mod = load_reference()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gates = random_circuit(n, layers, seed)
    if kind == "reference":
        state = torch.zeros(2 ** n, dtype=torch.complex64)
        state[0] = 1
        tg = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in gates]

        def run(psi):
            for qs, u in tg:
                psi = mod.apply_operator(operator=u, qubits=qs, state=psi)
            return psi
        threads = torch.get_num_threads()
    else:
        state = np.zeros(2 ** n, dtype=np.complex64)
        state[0] = 1
        tg = [(qs, u.astype(np.complex64)) for qs, u in gates]

        def run(psi):
            for qs, u in tg:
                psi = mod.apply_operator(u, qs, psi)
            return psi
        threads = 1
    times = []
    psi = state
    for i in range(warm + reps):
        t0 = time.perf_counter()
        psi = run(psi)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    updates = len(gates) * float(2 ** n)
    what = ("unmodified reference (baseline/_ref, unitair.simulation.apply_operator per gate)" if kind == "reference"
            else "numpy port of the reference (oracle/unitair_oracle.py; baseline/_ref did not travel)")
    return {"value": updates / float(np.median(times)), "unit": UNIT, "cores": threads,
            "host_cores": cores, "kind": kind,
            "sample": f"{what}: {layers} layer(s) of the bench recipe ({len(gates)} gates) on a {n}-qubit "
                      f"complex64 state, median of {reps} after {warm} warm-up, torch CPU threads={threads}",
            "sample_qubits": n, "sample_layers": layers, "sample_gates": len(gates),
            "sample_seconds": float(np.median(times))}, float(np.median(times))


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host
    cores.  A step of this arm is a BOUNDED SAMPLE of the workload (the 30-qubit state would need
    ~40 GiB and minutes per layer on the host): `config` names both the workload the other arm
    runs and the sample that was timed here; value is updates/s, which is size-independent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_qubits
    base, t_step = cpu_layer_rate(n, args.cpu_layers, args.seed, reps=max(1, args.steps),
                                  warm=max(1, min(args.warmup, 2)))
    cfg = workload_config(args, args.gpus)
    cfg["timed_sample"] = {"qubits": n, "layers": args.cpu_layers, "gates": base["sample_gates"],
                           "note": "ms_per_step is one pass over this sample, not over the workload above"}
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": max(1, min(args.warmup, 2)),
        "ms_per_step": t_step * 1e3, "step_is_sample": True, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "c64", "data": "synthetic",
        "config": cfg, "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus):
    n_total = args.qubits + int(round(np.log2(n_gpus)))
    return {
        "workload": f"random 1-/2-qubit circuit (Haar U(2) on every qubit + Haar U(4) on random "
                    f"ordered pairs per layer, SURVEY.md 8d C2 recipe), {args.layers} layers, "
                    f"{n_total}-qubit complex64 state" + (f" sharded over {n_gpus} GPUs by the "
                    f"{int(round(np.log2(n_gpus)))} highest-order qubits" if n_gpus > 1 else " on 1 GPU"),
        "qubits": n_total, "qubits_per_gpu": args.qubits, "layers": args.layers,
        "parallelism": "single GPU" if n_gpus == 1 else f"state sharded over {n_gpus} GPUs (global-qubit swaps)",
        "l2_policy": f"inputs larger than L2 ({8 * 2 ** args.qubits / 2 ** 30:.0f} GiB state per GPU vs 126 MB L2)",
    }



class gpu_local_cpus:
    """Run a block on the host cores next to a GPU (sysfs local_cpulist of its PCI device), so
    that pinned buffers allocated inside it are first-touched on the GPU's NUMA node."""

    def __init__(self, index):
        self.index = index
        self.old = None
        self.info = None

    def __enter__(self):
        try:
            bus = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=pci.bus_id",
                                  "--format=csv,noheader"], capture_output=True, text=True, timeout=20).stdout.strip()
            dom, rest = bus.split(":", 1)
            path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/local_cpulist"
            cpus = set()
            for part in open(path).read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            allowed = os.sched_getaffinity(0)
            if cpus & allowed:
                self.old = allowed
                os.sched_setaffinity(0, cpus & allowed)
                self.info = f"{len(cpus & allowed)} cores local to GPU {self.index}"
        except Exception:
            self.old = None
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            os.sched_setaffinity(0, self.old)
        return False


# --------------------------------------------------------------------------- multi-GPU helpers
def sharded_parity_check(world, rank, dev, n_total=24, layers=3, seed=77):
    """Correctness inside the artefact the driver runs: a 24-qubit circuit of the bench recipe on
    the state sharded over all ranks, with BOTH exchange formulations (fused peer-memory stores,
    bit permutation + NCCL send/recv), against the single-GPU engine on rank 0.  Returns a dict;
    `ok` is False when any relative error exceeds the 1e-5 tolerance of BASELINE.json."""
    import torch
    import torch.distributed as dist
    from unitair_b200 import circuit, sharded
    gates_np = random_circuit(n_total, layers, seed)
    # host gate tensors: the shards of this small state then go through the same register-blocked
    # pass kernel as the timed run (device gates on a small state stay on the other pass kernel)
    gates = [(qs, torch.as_tensor(u.astype(np.complex64))) for qs, u in gates_np]
    out = {"qubits": n_total, "layers": layers, "gates": len(gates), "tolerance": 1e-5, "ok": True}
    ref = None
    if rank == 0:
        full = torch.zeros(2 ** n_total, dtype=torch.complex64, device=dev)
        full[0] = 1
        ref = circuit.CompiledCircuit(gates, n_total, torch.complex64).run(full, in_place=True)
    for mode in ("p2p", "nccl"):
        try:
            st = sharded.ShardedState.zero_state(n_total, torch.complex64, dev)
            sc = sharded.ShardedCircuit(gates, n_total, torch.complex64, world, restore=True, exchange=mode)
            sc.run(st)
            sc.run(st)                                   # replayable plan: apply it twice ...
            got = st.gather_logical()
            nrm = float(st.norm_squared().item())
            if rank == 0:
                twice = circuit.CompiledCircuit(gates, n_total, torch.complex64).run(ref.clone(), in_place=True)
                err = float((torch.linalg.vector_norm(got - twice) / torch.linalg.vector_norm(twice)).item())
                out[mode] = {"rel_err_vs_single_gpu": err, "norm_squared": nrm, "swaps": sc.num_swaps,
                             "fused_swaps": sc.num_fused_swaps, "peer_stores": bool(sc.p2p)}
                if not (err < 1e-5 and abs(nrm - 1) < 1e-4):
                    out["ok"] = False
            st.release_peers()
            del st, sc, got
        except Exception as e:  # pragma: no cover
            out[mode] = {"error": repr(e)[:300]}
            out["ok"] = False
    flag = torch.tensor([1.0 if out["ok"] else 0.0], device=dev)
    dist.broadcast(flag, src=0)
    out["ok"] = bool(flag.item() > 0.5)
    torch.cuda.empty_cache()
    return out


def timed_sharded_run(n_total, layers, seed, world, rank, dev, exchange_mode, warm, steps, joint=True):
    """The bench circuit on an n_total-qubit state over `world` ranks (world == 1: single-GPU
    engine): device time per step (max over ranks), norm check after the timed steps.
    joint: the warm-up steps and the timed steps are each planned as ONE gate list (as the main
    measurement does); False: one plan per step, chained through the qubit layout."""
    import torch
    import torch.distributed as dist
    from unitair_b200 import circuit, sharded, _lib
    gates_np = random_circuit(n_total, layers, seed)
    gates = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in gates_np]
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    if world == 1:
        state = torch.zeros(2 ** n_total, dtype=torch.complex64, device=dev)
        state[0] = 1
        cc = circuit.CompiledCircuit(gates, n_total, torch.complex64)
        for _ in range(warm):
            cc.run(state, in_place=True)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            cc.run(state, in_place=True)
        e1.record()
        torch.cuda.synchronize()
        from unitair_b200.states import norm_squared
        nrm = float(norm_squared(state).item())
        elapsed = e0.elapsed_time(e1) / 1e3
        info = {"passes_per_step": cc.num_passes}
        launches = _lib.launch_count() - l0
        del state
    else:
        sstate = sharded.ShardedState.zero_state(n_total, torch.complex64, dev)
        plans, layout = [], sharded.identity_layout(n_total)
        if joint:
            for reps in (warm, steps):
                pl_ = sharded.ShardedCircuit(gates * reps, n_total, torch.complex64, world, layout=layout,
                                             restore=False, exchange=exchange_mode)
                plans.append(pl_)
                layout = pl_.end_layout
            n_warm_plans = 1
        else:
            for _ in range(warm + steps):
                pl_ = sharded.ShardedCircuit(gates, n_total, torch.complex64, world, layout=layout, restore=False,
                                             exchange=exchange_mode)
                plans.append(pl_)
                layout = pl_.end_layout
            n_warm_plans = warm
        for pl_ in plans[:n_warm_plans]:
            pl_.run(sstate)
        dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0.record()
        for pl_ in plans[n_warm_plans:]:
            pl_.run(sstate)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - l0
        t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
        nrm = float(sstate.norm_squared().item())
        timed = plans[n_warm_plans:]
        info = {"passes_per_step": sum(p_.num_passes for p_ in timed) / steps,
                "swaps_per_step": sum(p_.num_swaps for p_ in timed) / steps,
                "swap_bytes_per_gpu_per_step": sum(p_.swap_bytes_per_step for p_ in timed) / steps,
                "planning": "timed steps planned as one gate list" if joint else "one plan per step"}
        sstate.release_peers()
        del sstate, plans
    torch.cuda.empty_cache()
    updates = len(gates_np) * float(2 ** n_total)
    info.update({"qubits": n_total, "n_gpus": world, "steps": steps, "warmup": warm,
                 "ms_per_step": elapsed / steps * 1e3, "value": updates * steps / elapsed, "unit": UNIT,
                 "norm_squared_after": nrm, "norm_ok": bool(abs(nrm - 1) < 1e-3), "gpu_launches": int(launches)})
    return info


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import unitair_b200 as ua
    from unitair_b200 import _lib, circuit

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    n_local = args.qubits
    g_bits = int(round(np.log2(world)))
    assert 2 ** g_bits == world, "--gpus must be a power of two"
    n_total = n_local + g_bits
    gates_np = random_circuit(n_total, args.layers, args.seed)
    gates_dev = [(qs, torch.as_tensor(u.astype(np.complex64)).to(dev)) for qs, u in gates_np]
    num_gates = len(gates_np)
    updates_per_step = num_gates * float(2 ** n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if world == 1:
        state = torch.zeros(2 ** n_total, dtype=torch.complex64, device=dev)
        state[0] = 1
        cc = circuit.CompiledCircuit(gates_dev, n_total, torch.complex64)
        step = lambda: cc.run(state, in_place=True)   # noqa: E731
        launches_per_step = cc.num_passes
        kernel_name = "cluster_ring_kernel" if getattr(cc, "mats_host", None) is not None else "fused_pass_kernel"
        extra = {"passes_per_step": cc.num_passes, "gates_per_step": num_gates}
    else:
        from unitair_b200 import sharded
        sstate = sharded.ShardedState.zero_state(n_total, torch.complex64, dev)
        # Every step applies the same 10-layer circuit to the state the previous step left, i.e.
        # the steps together are one deep circuit, and that is how they are planned: the W warm-up
        # steps as one gate list, the K timed steps as another (pre-built outside the timed region,
        # exactly as the single-GPU arm pre-compiles its circuit).  The qubit layout is carried
        # over (no swap-back), and the few gates at the end of a step that wait for a global qubit
        # share an exchange with the next step instead of paying one of their own.  The same K
        # steps with one plan PER STEP are timed as well and reported under `per_step_plans`.
        n_warm = max(args.warmup, 3)

        def build_plans(exchange):
            warm_ = sharded.ShardedCircuit(gates_dev * n_warm, n_total, torch.complex64, world,
                                           restore=False, exchange=exchange)
            timed_ = sharded.ShardedCircuit(gates_dev * args.steps, n_total, torch.complex64, world,
                                            layout=warm_.end_layout, restore=False, exchange=exchange)
            return warm_, timed_
        exchange_mode = None
        warm_plan, plan = build_plans(exchange_mode)
        if plan.p2p:
            # peer mapping (CUDA IPC) is set up collectively on first use: if this box refuses it,
            # every rank fails the same way and falls back to the send/recv exchange -- loudly
            try:
                sstate.peer_pointers()
            except Exception as e:  # pragma: no cover
                if rank == 0:
                    print(f"bench: peer-memory mapping failed ({e!r}); using the NCCL send/recv exchange",
                          file=sys.stderr, flush=True)
                exchange_mode = "nccl"
                warm_plan, plan = build_plans(exchange_mode)
        launches_per_step = plan.num_passes / args.steps
        kernel_name = "cluster_ring_kernel"
        extra = {"passes_per_step": launches_per_step, "gates_per_step": num_gates,
                 "swaps_per_step": plan.num_swaps / args.steps,
                 "fused_swaps_per_step": plan.num_fused_swaps / args.steps,
                 "exchange": "p2p (last pass of an epoch stores into the peers' memory)" if plan.p2p
                             else "nccl (bit permutation pass + grouped send/recv)",
                 "swap_bytes_per_gpu_per_step": plan.swap_bytes_per_step / args.steps,
                 "planning": f"the {args.steps} timed steps are planned as one gate list"}

    check = None
    if world > 1 and not args.no_check:
        check = sharded_parity_check(world, rank, dev)
    if world == 1:
        for _ in range(max(args.warmup, 3)):
            step()
    else:
        warm_plan.run(sstate)
    barrier()
    sampler.start()
    l0 = _lib.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if world == 1:
        for _ in range(args.steps):
            step()
    else:
        plan.run(sstate)                  # args.steps steps
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count() - l0
    elapsed = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = updates_per_step * args.steps / elapsed

    # roofline of the dominant kernel: algorithmic bytes per launch / average launch time
    peak, peak_src = measured_hbm_peak()
    bytes_per_launch = 16.0 * 2 ** n_local                     # read + write of the local state
    avg_launch_s = elapsed / max(1, launches)
    achieved = bytes_per_launch / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "traffic": recorded_traffic(kernel_name),
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "avg_launch_ms": avg_launch_s * 1e3,
                "note": "launch time = timed region / native launches" +
                        ("" if world == 1 else " (includes NVLink swap time)")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "c64",
        "data": "synthetic", "config": workload_config(args, world), "roofline": roofline,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    line.update(extra)
    # correctness of what was just timed: the state is still normalised (unitary circuit) ...
    if world == 1:
        nrm = float(ua.norm_squared(state).item())
    else:
        nrm = float(sstate.norm_squared().item())
    norm_ok = bool(abs(nrm - 1) < 1e-3)
    if world > 1:
        # ... and the sharded engine agrees with the single-GPU engine (both exchange modes)
        if check is None:
            check = {"ok": True, "skipped": "--no-check"}
        check["norm_squared_after_timed_steps"] = nrm
        check["ok"] = bool(check["ok"] and norm_ok)
        line["check"] = check
    else:
        line["check"] = {"norm_squared_after_timed_steps": nrm, "ok": norm_ok}
    big = n_local > 31          # 32+ qubits per GPU: one state copy only, no pinned host mirror
    if big:
        line["skipped"] = "per_gate / e2e sections need extra state copies (>= 64 GiB each): skipped at this size"
    if world > 1 and not (big or args.no_e2e):
        # ---- end to end, sharded: every rank moves its shard from/to pinned host memory and
        # plans + runs the circuit through the public sharded API inside the timed region
        try:
            shard_elems = 2 ** n_local
            with gpu_local_cpus(local_rank) as numa:
                h_in = torch.zeros(shard_elems, dtype=torch.complex64).pin_memory()
                h_out = torch.zeros(shard_elems, dtype=torch.complex64).pin_memory()
            if rank == 0:
                h_in[0] = 1
            h_gates = [(qs, torch.as_tensor(u.astype(np.complex64)).pin_memory()) for qs, u in gates_np]
            gate_bytes = sum(u.numel() * 8 for _, u in h_gates)

            # the public host-memory API for shards: every job uploads this rank's shard from pinned
            # host memory, plans + packs + runs the sharded circuit and downloads the result; the
            # three stages of consecutive jobs overlap (ShardedHostStream: 3 streams, staging
            # buffers next to the IPC-mapped state/spare pair)
            hs = ua.ShardedHostStream(n_total, torch.complex64, dev, exchange=exchange_mode, state=sstate)

            def run_jobs(k, serial):
                barrier()
                e0.record()
                for _ in range(k):
                    hs.submit(h_gates, h_in, h_out)
                    if serial:
                        hs.join()
                hs.join()
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1) / 1e3], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            run_jobs(1, True)
            t_serial = run_jobs(2, True) / 2
            k_e2e = max(8, args.steps)
            t_e2e = run_jobs(k_e2e, False)
            line["e2e"] = {"value": updates_per_step * k_e2e / t_e2e, "unit": UNIT,
                           "h2d_bytes_per_step": int(8 * shard_elems * world),
                           "d2h_bytes_per_step": int(8 * shard_elems * world), "steps": k_e2e,
                           "ms_per_step": t_e2e / k_e2e * 1e3,
                           "one_job_at_a_time_ms_per_step": t_serial * 1e3,
                           "host_GBs_per_gpu_per_direction": 8 * shard_elems * k_e2e / t_e2e / 1e9,
                           "pinned_buffers": numa.info or "default NUMA placement",
                           "what": "ShardedHostStream.submit per step and rank: pinned host shard -> device, "
                                   "ShardedCircuit (planning, merging, packing; host gate matrices travel as "
                                   "kernel parameters) + run with fused NVLink exchanges, shard -> pinned host; "
                                   "upload of step k+1, circuit of step k, download of step k-1 overlap; "
                                   "results in the plan's end layout"}
            del hs
            sstate.layout = sharded.identity_layout(n_total)
            del h_in, h_out
        except Exception as e:  # pragma: no cover
            line["e2e"] = {"value": None, "unit": UNIT, "error": repr(e)[:200]}
    if world > 1 and not (args.no_extra or big):
        # the same K steps with one plan per step (each step's leftover gates pay their own exchange)
        try:
            ps = timed_sharded_run(n_total, args.layers, args.seed, world, rank, dev, exchange_mode,
                                   max(args.warmup, 3), args.steps, joint=False)
            line["per_step_plans"] = {k: ps[k] for k in ("ms_per_step", "value", "passes_per_step", "swaps_per_step",
                                                         "swap_bytes_per_gpu_per_step", "norm_ok", "planning")}
        except Exception as e:  # pragma: no cover
            line["per_step_plans"] = {"error": repr(e)[:200]}
    if world > 1:
        # untimed extra step with per-phase device timing (permute / exchange / gates), rank 0
        tm = {}
        tplan = sharded.ShardedCircuit(gates_dev, n_total, torch.complex64, world, layout=sstate.layout,
                                       restore=False, exchange=exchange_mode)
        tplan.run(sstate, timing=tm)
        line["phase_ms_per_step_rank0"] = {k: round(v, 2) for k, v in tm.items()}
        nv = None
        if tplan.p2p and tm.get("scatter_pass"):
            # the passes whose stores cross NVLink: bytes leaving this GPU / their device time
            nv = tplan.swap_bytes_per_step / (tm["scatter_pass"] / 1e3) / 1e9
            what = "peer stores of the fused scatter passes (compute included in the pass time)"
        elif tm.get("exchange") and not tplan.p2p:
            nv = tplan.swap_bytes_per_step / (tm["exchange"] / 1e3) / 1e9
            what = "NCCL send/recv exchange"
        if nv is not None:
            line["nvlink"] = {"GBs_per_gpu_per_direction": round(nv, 1), "frac_of_900": round(nv / 900.0, 3),
                              "frac_of_measured_770": round(nv / 770.0, 3), "what": what,
                              "bytes_per_gpu_per_step": tplan.swap_bytes_per_step}

    if world == 1 and not big:
        # ---- per-gate path (the reference's call pattern: one apply_operator per gate) ----
        psi = state
        one_layer = gates_dev[: num_gates // args.layers]
        for qs, u in one_layer[:4]:
            psi = ua.simulation.apply_operator(u, qs, psi)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0.record()
        for qs, u in one_layer:
            psi = ua.simulation.apply_operator(u, qs, psi)
        e1.record()
        torch.cuda.synchronize()
        t_pg = e0.elapsed_time(e1) / 1e3
        n_pg = _lib.launch_count() - l0
        ach = bytes_per_launch * n_pg / t_pg / 1e9
        line["per_gate"] = {
            "value": len(one_layer) * float(2 ** n_total) / t_pg, "unit": UNIT,
            "what": "simulation.apply_operator once per gate, one layer, out-of-place",
            "roofline": {"bound": "hbm", "kernel": "gate_direct_kernel", "achieved": ach, "peak": peak,
                         "unit": "GB/s", "frac": ach / peak, "frac_of_8TBps": ach / 8000.0,
                         "traffic": recorded_traffic("gate_direct_kernel"),
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": t_pg / n_pg * 1e3}}
        del psi

    if world == 1 and not (big or args.no_e2e):
        # ---- end to end through the public API with host buffers -------------------------
        try:
            with gpu_local_cpus(local_rank) as numa:
                h_state = torch.zeros(2 ** n_total, dtype=torch.complex64).pin_memory()
                h_out = torch.zeros(2 ** n_total, dtype=torch.complex64).pin_memory()
            h_state[0] = 1
            h_gates = [(qs, torch.as_tensor(u.astype(np.complex64)).pin_memory()) for qs, u in gates_np]
            gate_bytes = sum(u.numel() * 8 for _, u in h_gates)
            del state
            torch.cuda.empty_cache()

            # the public host-memory API: every job uploads the state and the gates from pinned
            # host memory, plans + packs + runs the circuit and downloads the result; the three
            # stages of consecutive jobs overlap on three CUDA streams (HostCircuitStream)
            hs = ua.HostCircuitStream(n_total, torch.complex64, dev)

            def run_jobs(k, serial):
                e0.record()
                for _ in range(k):
                    hs.submit(h_gates, h_state, h_out)
                    if serial:
                        hs.join()
                hs.join()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / 1e3
            run_jobs(1, True)
            k_e2e = max(8, args.steps)
            t_serial = run_jobs(2, True) / 2
            t_e2e = run_jobs(k_e2e, False)
            line["e2e"] = {"value": updates_per_step * k_e2e / t_e2e, "unit": UNIT,
                           "h2d_bytes_per_step": int(8 * 2 ** n_total + gate_bytes),
                           "d2h_bytes_per_step": int(8 * 2 ** n_total), "steps": k_e2e,
                           "ms_per_step": t_e2e / k_e2e * 1e3,
                           "one_job_at_a_time_ms_per_step": t_serial * 1e3,
                           "pinned_buffers": numa.info or "default NUMA placement",
                           "what": "HostCircuitStream.submit per step: pinned host state + gates -> device, "
                                   "planning + packing + fused passes, final state -> pinned host; upload of "
                                   "step k+1, circuit of step k and download of step k-1 overlap (3 streams, "
                                   "3 device buffers); one_job_at_a_time = the same call with a join after "
                                   "every job (upload, circuit, download in sequence)"}
            del h_state, h_out, hs
        except Exception as e:  # pragma: no cover
            line["e2e"] = {"value": None, "unit": UNIT, "error": repr(e)[:200]}

    if world == 1:
        # ---- CPU baseline (reference's own path on this box's host cores) -----------------
        if not args.no_cpu_baseline:
            try:
                base, _ = cpu_layer_rate(args.cpu_qubits, args.cpu_layers, args.seed, reps=3, warm=1)
                line["cpu_baseline"] = base
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "error": repr(e)[:200]}

    # ---- extra sections (BASELINE config 5): the 33-qubit strong-scaling point of this GPU count
    # and, on 8 GPUs, the 36-qubit run -- driver-visible, each with its own norm check
    if not args.no_extra and args.total_qubits is None and args.qubits == 30:
        try:
            if world > 1:
                sstate.release_peers()
                dist.barrier()
                sstate.local = sstate.spare = None
                warm_plan = plan = tplan = None      # noqa: F841  (drop references before the big states)
            else:
                state = None
            torch.cuda.empty_cache()
            if world == 8:
                line["strong_scaling_33q"] = {"qubits": 33, "n_gpus": 8, "ms_per_step": line["ms_per_step"],
                                              "value": line["value"], "unit": UNIT,
                                              "note": "identical to this run (33 qubits = 30 per GPU on 8 GPUs)"}
                line["config5_36q"] = timed_sharded_run(36, args.layers, args.seed + 5, world, rank, dev,
                                                        exchange_mode, warm=2, steps=2)
            else:
                line["strong_scaling_33q"] = timed_sharded_run(33, args.layers, args.seed, world, rank, dev,
                                                               exchange_mode if world > 1 else None, warm=2, steps=3)
        except Exception as e:  # pragma: no cover
            line["extra_error"] = repr(e)[:300]
    # BASELINE.json's other single-GPU configs at full size (C1: 10 qubits x batch 1024, C2: 24 qubits
    # x 200 layers, C3: 16 qubits x batch 4096 forward + gradient, C4: 30 qubits complex128 with
    # 5-qubit blocks on the FP64 tensor cores + phase layers): timings only, parity is tests/
    if world == 1 and not args.no_extra and args.total_qubits is None and args.qubits == 30:
        try:
            state = None
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
            import bench_configs as bc
            cfgs = {}
            for name in ("c1", "c2", "c3", "c4"):
                try:
                    cfgs[name] = getattr(bc, name)()
                except Exception as e:  # pragma: no cover
                    cfgs[name] = {"error": repr(e)[:200]}
                torch.cuda.empty_cache()
            line["configs"] = cfgs
        except Exception as e:  # pragma: no cover
            line["configs"] = {"error": repr(e)[:300]}
    failed = not line["check"]["ok"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if failed:
        raise SystemExit("bench: correctness check failed (see the `check` key)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=int(os.environ.get("UA_BENCH_QUBITS", 30)),
                    help="qubits per GPU (the state has qubits + log2(gpus) qubits)")
    ap.add_argument("--total-qubits", type=int, default=None,
                    help="strong scaling: fixed state size, qubits per GPU = total - log2(gpus) "
                         "(BASELINE config 5: 33 qubits on 1/2/4/8 GPUs)")
    ap.add_argument("--layers", type=int, default=10)
    ap.add_argument("--seed", type=int, default=202)
    ap.add_argument("--cpu-qubits", type=int, default=24,
                    help="size of the bounded CPU sample for the reference arm")
    ap.add_argument("--cpu-layers", type=int, default=2, help="layers of the bounded CPU sample")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra sections (33-qubit strong-scaling point, 36-qubit run at 8 GPUs)")
    ap.add_argument("--no-check", action="store_true", help="skip the sharded-vs-single-GPU parity check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.total_qubits is not None:
        g_bits = int(round(np.log2(args.gpus)))
        args.qubits = args.total_qubits - g_bits
        args.scaling = "strong"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
