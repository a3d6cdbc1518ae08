"""Circuits on host-resident states: a three-stage pipeline over PCIe.

The reference's API takes and returns torch tensors; a user whose states live in host memory
pays `state.to('cuda')`, the circuit, and `.cpu()` one after the other.  For a 30-qubit
complex64 state that is 8 GiB each way at PCIe speed (~55 GB/s): the copies cost twice the
circuit.  `HostCircuitStream` keeps three device buffers in flight so that the upload of job
k+1, the circuit of job k and the download of job k-1 run at the same time on three CUDA
streams; in steady state a job costs max(upload, circuit, download) instead of their sum.

    stream = HostCircuitStream(num_qubits, torch.complex64, device)
    for gates, h_in, h_out in jobs:            # pinned host tensors
        stream.submit(gates, h_in, h_out)      # returns at once
    stream.drain()                              # all h_out are valid

Every job is planned (gate merging, pass planning, packing) inside `submit`, like
circuit.apply_gates does; the plan of an identical gate list can be reused with `compiled=`.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import circuit


class HostCircuitStream:
    def __init__(self, num_qubits: int, dtype: torch.dtype, device, depth: int = 3):
        if depth < 1:
            raise ValueError("depth must be at least 1")
        self.n = num_qubits
        self.dtype = dtype
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("unitair_b200 runs only on CUDA devices; there is no CPU fallback")
        L.lib()
        self.depth = depth
        self.buffers = [torch.empty(1 << num_qubits, dtype=dtype, device=self.device) for _ in range(depth)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.free = [None] * depth          # event: the download that last used the slot has finished
        self.jobs = 0
        self._started = False

    def _start(self):
        # everything queued here runs after the work already queued on the caller's stream
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        self._started = True

    def submit(self, gates: Sequence[Tuple[Sequence[int], torch.Tensor]], h_in: torch.Tensor,
               h_out: torch.Tensor, compiled: Optional[circuit.CompiledCircuit] = None):
        """Queue one job: h_in (host, ideally pinned) -> device, the gate list (host or device
        operators), result -> h_out (host, ideally pinned).  Returns the CompiledCircuit used."""
        if h_in.dtype != self.dtype or h_in.numel() != 1 << self.n or h_out.numel() != 1 << self.n:
            raise RuntimeError("host state does not match the stream's size / dtype")
        if not self._started:
            self._start()
        slot = self.jobs % self.depth
        buf = self.buffers[slot]
        with torch.cuda.stream(self.s_in):
            if self.free[slot] is not None:
                self.s_in.wait_event(self.free[slot])
            buf.copy_(h_in, non_blocking=True)
            d_gates = None
            if compiled is None:
                if all(u.device.type == "cpu" for _, u in gates):
                    # host operators: merged on the host, their values feed the register-blocked
                    # pass kernel as launch parameters -- nothing to upload, no synchronisation
                    d_gates = [(qs, u) for qs, u in gates]
                else:
                    d_gates = [(qs, u.to(self.device, non_blocking=True)) for qs, u in gates]
                    for _, u in d_gates:
                        if u.device.type == "cuda":
                            u.record_stream(self.s_run)      # allocated on s_in, consumed on s_run
            ev_in = torch.cuda.Event()
            ev_in.record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(ev_in)
            if compiled is None:
                compiled = circuit.CompiledCircuit(d_gates, self.n, self.dtype)
            compiled.run(buf, in_place=True)
            ev_run = torch.cuda.Event()
            ev_run.record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_run)
            h_out.copy_(buf, non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self.s_out)
        self.free[slot] = ev_out
        self.jobs += 1
        return compiled

    def join(self):
        """Make the caller's current stream wait for every queued job (no host blocking)."""
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)
        self._started = False

    def drain(self):
        """Block the host until every queued job has delivered its result."""
        self.join()
        torch.cuda.current_stream(self.device).synchronize()


class ShardedHostStream:
    """`HostCircuitStream` for a state sharded over the GPUs of a `torchrun` job: every rank
    streams its own shard.  Upload of job k+1 (pinned host shard -> a staging buffer), the
    sharded circuit of job k (staging -> state buffer, epochs with their fused exchanges over
    NVLink peer memory, state -> staging) and the download of job k-1 overlap on three streams.

    The state/spare pair that the peers have mapped (CUDA IPC) is allocated once and stays put;
    only the staging buffers rotate, so no peer mapping is ever re-opened.  Collective: all ranks
    submit the same jobs in the same order.  Results are delivered in the qubit layout the plan
    ends in (`plan.end_layout`; pass `restore=True` for the identity layout at the price of the
    restoring exchange).
    """

    def __init__(self, num_qubits: int, dtype: torch.dtype, device, group=None, depth: int = 2,
                 exchange: Optional[str] = None, state=None):
        from . import sharded
        self._sharded = sharded
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("unitair_b200 runs only on CUDA devices; there is no CPU fallback")
        if depth < 1:
            raise ValueError("depth must be at least 1")
        L.lib()
        self.n = num_qubits
        self.dtype = dtype
        self.exchange = exchange
        self.depth = depth
        self.state = state if state is not None else sharded.ShardedState.zero_state(num_qubits, dtype, device, group)
        if self.state.n != num_qubits or self.state.local.dtype != dtype:
            raise RuntimeError("state does not match the stream's size / dtype")
        self.world = self.state.world
        shard = self.state.local.numel()
        self.stage_in = [torch.empty(shard, dtype=dtype, device=self.device) for _ in range(depth)]
        self.stage_out = [torch.empty(shard, dtype=dtype, device=self.device) for _ in range(depth)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.in_free = [None] * depth       # the circuit stage has taken the upload out of the slot
        self.out_free = [None] * depth      # the download that last used the slot has finished
        self.jobs = 0
        self._started = False

    def _start(self):
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        self._started = True

    def submit(self, gates, h_in: torch.Tensor, h_out: torch.Tensor, plan=None, restore: bool = False):
        """Queue one job on this rank's shard: h_in / h_out are host tensors (ideally pinned) of
        2^(n - log2 world) amplitudes; gate operators may live on the host (merged there, never
        uploaded) or on the device.  Returns the ShardedCircuit used (reusable via `plan=`)."""
        shard = self.state.local.numel()
        if h_in.dtype != self.dtype or h_in.numel() != shard or h_out.numel() != shard:
            raise RuntimeError("host shard does not match the stream's size / dtype")
        if not self._started:
            self._start()
        slot = self.jobs % self.depth
        with torch.cuda.stream(self.s_in):
            if self.in_free[slot] is not None:
                self.s_in.wait_event(self.in_free[slot])
            self.stage_in[slot].copy_(h_in, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.s_in)
        if plan is None:
            plan = self._sharded.ShardedCircuit(gates, self.n, self.dtype, self.world, restore=restore,
                                                exchange=self.exchange)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(ev_in)
            st = self.state
            st.local.copy_(self.stage_in[slot], non_blocking=True)
            st.layout = self._sharded.identity_layout(self.n)
            taken = torch.cuda.Event()
            taken.record(self.s_run)
            self.in_free[slot] = taken
            plan.run(st)
            if self.out_free[slot] is not None:
                self.s_run.wait_event(self.out_free[slot])
            self.stage_out[slot].copy_(st.local, non_blocking=True)
            ev_run = torch.cuda.Event()
            ev_run.record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_run)
            h_out.copy_(self.stage_out[slot], non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self.s_out)
        self.out_free[slot] = ev_out
        self.jobs += 1
        return plan

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_in, self.s_run, self.s_out):
            cur.wait_stream(s)
        self._started = False

    def drain(self):
        self.join()
        torch.cuda.current_stream(self.device).synchronize()
