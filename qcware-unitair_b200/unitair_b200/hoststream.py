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
