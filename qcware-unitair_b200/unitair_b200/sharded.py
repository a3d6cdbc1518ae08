"""States too large for one GPU: shard by the highest-order index bits, one process per GPU.

The reference is single-device (no distributed code at all); this module is the
multi-GPU extension BASELINE.json asks for.  An n-qubit state is split over world = 2^g ranks:
rank r owns the amplitudes whose top g index bits equal r, i.e. a contiguous slice of the
vector layout.  With the identity layout unitair qubits 0..g-1 (the most significant bits,
src/unitair/states/conversions.py:43-45) are "global".

* Gates on local qubits run on the local shard with the single-GPU engine (fused passes),
  no communication.
* A gate on a global qubit first makes that qubit local: m global qubits are swapped with m
  local "victim" qubits in ONE exchange step.  The victims are moved to the top m local bits
  (one bit-permutation pass, skipped when they are already there); then the shard is 2^m
  contiguous blocks and rank(a) sends block b to rank(b) and receives that rank's block a
  (NCCL send/recv over NVLink, all 2^m - 1 peers in one group: (1 - 2^-m) of the shard leaves
  and enters every GPU).  Afterwards only the logical->physical qubit map changes; nothing is
  swapped back eagerly.
* Victims are chosen by farthest next use (Belady), so a layer of a random circuit needs one
  exchange per layer instead of one per global-qubit gate.

The planning code is pure host logic and is exercised on CPU with the gloo backend
(tests/test_sharded_gloo.py); the local work there is done by an injected test engine.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


# --------------------------------------------------------------------------- #
# planning (host only)
# --------------------------------------------------------------------------- #
@dataclass
class Epoch:
    """One exchange (possibly empty) followed by a run of local gates."""
    incoming: List[int] = field(default_factory=list)   # logical qubits becoming local
    victims: List[int] = field(default_factory=list)    # logical qubits becoming global
    rank_bits: List[int] = field(default_factory=list)  # rank bit index each pair swaps
    perm_src: Optional[List[int]] = None                # local bit permutation before the exchange
    gates: List[int] = field(default_factory=list)      # gate indices executed after it
    local_bits: List[List[int]] = field(default_factory=list)  # physical bit positions per gate


def identity_layout(n: int) -> List[int]:
    """phys[q] = physical index bit of logical qubit q: qubit 0 is the most significant bit."""
    return [n - 1 - q for q in range(n)]


def plan_epochs(gate_qubits: Sequence[Sequence[int]], n: int, g: int,
                layout: Optional[List[int]] = None, restore: bool = True,
                lookahead: int = 4096) -> Tuple[List[Epoch], List[int]]:
    """Cut a gate list into epochs for a state sharded over 2^g ranks.

    layout[q] is the physical bit of logical qubit q; bits >= n-g are rank bits.  Returns the
    epochs and the final layout.  With restore=True a last exchange puts the layout back to
    the one the plan started from, so the same plan can be replayed step after step.
    """
    n_local = n - g
    phys = list(layout) if layout is not None else identity_layout(n)
    start_layout = list(phys)
    remaining = list(range(len(gate_qubits)))
    epochs: List[Epoch] = []
    cur = Epoch()

    def run_local_gates():
        nonlocal remaining
        blocked = set()
        keep = []
        for idx, gi in enumerate(remaining):
            qs = gate_qubits[gi]
            if idx < lookahead and not any(q in blocked for q in qs) and all(phys[q] < n_local for q in qs):
                cur.gates.append(gi)
                cur.local_bits.append([phys[q] for q in qs])
            else:
                blocked.update(qs)
                keep.append(gi)
        remaining = keep

    def make_exchange(incoming: List[int], victims: List[int]) -> Epoch:
        """Swap `incoming` (global) with `victims` (local); updates phys."""
        m = len(incoming)
        ep = Epoch(incoming=list(incoming), victims=list(victims))
        # victims must sit on the top m local bits: victim j -> local bit n_local - m + j
        want = {v: n_local - m + j for j, v in enumerate(victims)}
        qubit_at = {phys[q]: q for q in range(n) if phys[q] < n_local}
        if any(phys[v] != want[v] for v in victims):
            # new_bit_of[q] for every local qubit: victims to the top, the others keep their
            # relative order below
            others = sorted((p for p in range(n_local) if qubit_at[p] not in want), )
            new_phys = {}
            for slot, p in enumerate(others):
                new_phys[qubit_at[p]] = slot
            for v in victims:
                new_phys[v] = want[v]
            src = [0] * n_local            # output bit p takes input bit src[p]
            for q, newp in new_phys.items():
                src[newp] = phys[q]
            ep.perm_src = src
            for q, newp in new_phys.items():
                phys[q] = newp
        for j, (qin, v) in enumerate(zip(incoming, victims)):
            ep.rank_bits.append(phys[qin] - n_local)
            phys[qin], phys[v] = phys[v], phys[qin]
        return ep

    run_local_gates()
    while remaining:
        epochs.append(cur)
        # global qubits with pending gates, in order of first need
        first_use = {}
        for pos, gi in enumerate(remaining):
            for q in gate_qubits[gi]:
                first_use.setdefault(q, pos)
        globals_needed = sorted((q for q in first_use if phys[q] >= n_local), key=lambda q: first_use[q])
        incoming = globals_needed[:g]
        # Belady: evict the local qubits whose next use is farthest away
        never = len(remaining) + 1
        local_qubits = [q for q in range(n) if phys[q] < n_local]
        local_qubits.sort(key=lambda q: (-first_use.get(q, never), -phys[q]))
        victims = local_qubits[:len(incoming)]
        victims.sort(key=lambda q: phys[q])
        cur = make_exchange(incoming, victims)
        before = len(remaining)
        run_local_gates()
        if len(remaining) == before:
            raise RuntimeError("sharded planner made no progress (gate with more qubits than fit locally?)")
    epochs.append(cur)
    if restore:
        # Put every qubit back where the plan started so the plan can be replayed.  A rank bit r
        # held by the wrong qubit is fixed by swapping that qubit with the qubit whose home is r
        # (one exchange fixes every such bit at once); cycles among global qubits are broken by
        # parking one of them on a local bit first; a final local permutation sorts the rest.
        home_of_bit = {start_layout[q]: q for q in range(n)}
        guard = 0
        while phys != start_layout:
            guard += 1
            if guard > 4 * n + 8:
                raise RuntimeError("layout restore did not converge")
            bad = [q for q in range(n) if phys[q] >= n_local and phys[q] != start_layout[q]]
            if bad:
                pairs = [(q, home_of_bit[phys[q]]) for q in bad if phys[home_of_bit[phys[q]]] < n_local]
                if not pairs:
                    # only cycles among global qubits are left: park one on a local bit
                    park = next(q for q in range(n) if phys[q] < n_local and start_layout[q] < n_local)
                    pairs = [(bad[0], park)]
                pairs.sort(key=lambda t: phys[t[1]])
                epochs.append(make_exchange([t[0] for t in pairs], [t[1] for t in pairs]))
                continue
            ep = Epoch()
            src = [0] * n_local
            for q in range(n):
                if phys[q] < n_local:
                    src[start_layout[q]] = phys[q]
                    phys[q] = start_layout[q]
            ep.perm_src = src
            epochs.append(ep)
    return epochs, phys


# --------------------------------------------------------------------------- #
# local engines
# --------------------------------------------------------------------------- #
class CudaEngine:
    """Local work on the shard with the native single-GPU engine."""

    def __init__(self, dtype):
        self.dtype = dtype

    def compile(self, gates_local, n_local):
        from . import circuit
        return circuit.CompiledCircuit(gates_local, n_local, self.dtype) if gates_local else None

    def run(self, compiled, shard):
        if compiled is not None:
            compiled.run(shard, in_place=True)
        return shard

    def num_passes(self, compiled):
        return compiled.num_passes if compiled is not None else 0

    def permute(self, src_bits, shard, out):
        from . import _lib as L
        n_local = len(src_bits)
        dev = shard.device
        with L.on_device(dev):
            L.check(L.lib().ua_permute_bits(L.dtype_code(shard.dtype), out.data_ptr(), shard.data_ptr(),
                                            n_local, 1, L.int_array(src_bits), L.stream_ptr(dev)))
        return out


# --------------------------------------------------------------------------- #
# state + execution
# --------------------------------------------------------------------------- #
class ShardedState:
    """The local slice of an n-qubit state plus the logical->physical qubit map."""

    def __init__(self, local: torch.Tensor, num_qubits: int, group=None, layout: Optional[List[int]] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        g = self.world.bit_length() - 1
        if 1 << g != self.world:
            raise ValueError("world size must be a power of two")
        self.g = g
        self.n = num_qubits
        self.n_local = num_qubits - g
        if local.numel() != 1 << self.n_local:
            raise ValueError(f"local shard must hold 2^{self.n_local} amplitudes")
        self.local = local
        self.spare = None            # second buffer for exchanges / permutations (lazy)
        self.layout = list(layout) if layout is not None else identity_layout(num_qubits)

    @classmethod
    def zero_state(cls, num_qubits: int, dtype, device, group=None):
        """|0...0>: rank 0 holds the single 1."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        g = world.bit_length() - 1
        local = torch.zeros(1 << (num_qubits - g), dtype=dtype, device=device)
        if rank == 0:
            local[0] = 1
        return cls(local, num_qubits, group)

    def _spare(self):
        if self.spare is None:
            self.spare = torch.empty_like(self.local)
        return self.spare

    # -- reductions: local fused reduce + all_reduce of scalars ------------------------------
    def norm_squared(self) -> torch.Tensor:
        if self.local.is_cuda:
            from .states import norm_squared
            v = norm_squared(self.local).to(torch.float64)
        else:
            v = (self.local.real.double() ** 2 + self.local.imag.double() ** 2).sum()
        if self.world > 1:
            v = v.clone()
            dist.all_reduce(v, group=self.group)
        return v

    def _require_identity(self, what):
        if self.layout != identity_layout(self.n):
            raise RuntimeError(f"{what} needs the identity qubit layout (run plans built with "
                               f"restore=True end in it)")

    def local_index_offset(self) -> int:
        """Index of the first amplitude of this rank's shard in the full state (identity layout)."""
        self._require_identity("local_index_offset")
        return self.rank << self.n_local

    def apply_phase(self, angles_local: torch.Tensor) -> "ShardedState":
        """psi_k <- exp(-i angle_k) psi_k with this rank's slice of the angles (diagonal ops need no
        communication: every rank owns a contiguous slice of the index space)."""
        self._require_identity("apply_phase")
        from .simulation import apply_phase
        self.local = apply_phase(angles_local, self.local)
        self.spare = None
        return self

    def diag_expectation_value(self, diag_local: torch.Tensor) -> torch.Tensor:
        """sum_k d_k |psi_k|^2: local fused reduction + all_reduce of one scalar."""
        self._require_identity("diag_expectation_value")
        from .states import diag_expectation_value
        v = diag_expectation_value(diag_local, self.local).to(torch.float64)
        if self.world > 1:
            v = v.clone()
            dist.all_reduce(v, group=self.group)
        return v

    def gather_logical(self) -> Optional[torch.Tensor]:
        """Full state in the logical (unitair) qubit order on rank 0, None elsewhere.
        Test/debug helper: needs the whole state to fit on one device."""
        mine = torch.view_as_real(self.local)
        parts = [torch.empty_like(mine) for _ in range(self.world)] if self.rank == 0 else None
        if self.world > 1:
            dist.gather(mine, parts, dst=0, group=self.group)
        else:
            parts = [mine]
        if self.rank != 0:
            return None
        full = torch.view_as_complex(torch.cat(parts))   # physical order: rank bits on top
        n = self.n
        t = full.reshape((2,) * n)         # axis i <-> physical bit n-1-i
        axes = [n - 1 - self.layout[q] for q in range(n)]
        return t.permute(axes).reshape(-1).contiguous()


class ShardedCircuit:
    """A gate list planned once for a state sharded over `world` ranks and replayable
    (the plan ends in the layout it started from)."""

    def __init__(self, gates, num_qubits: int, dtype, world: int, engine=None,
                 layout: Optional[List[int]] = None, restore: bool = True):
        g = world.bit_length() - 1
        if 1 << g != world:
            raise ValueError("world size must be a power of two")
        self.n, self.g, self.world = num_qubits, g, world
        self.n_local = num_qubits - g
        self.dtype = dtype
        self.engine = engine if engine is not None else CudaEngine(dtype)
        self.gates = [([int(q) for q in qs], m) for qs, m in gates]
        for qs, _ in self.gates:
            if len(set(qs)) != len(qs) or not set(qs).issubset(range(num_qubits)):
                raise ValueError(f"qubits={qs} is not a valid target list on {num_qubits} qubits")
            if len(qs) > self.n_local:
                raise ValueError("a gate cannot act on more qubits than are local to a rank")
        self.start_layout = list(layout) if layout is not None else identity_layout(num_qubits)
        self.epochs, self.end_layout = plan_epochs([qs for qs, _ in self.gates], num_qubits, g,
                                                   self.start_layout, restore=restore)
        nl = self.n_local
        self.compiled = []
        for ep in self.epochs:
            local_gates = [([nl - 1 - p for p in bits], self.gates[gi][1])
                           for gi, bits in zip(ep.gates, ep.local_bits)]
            self.compiled.append(self.engine.compile(local_gates, nl))
        self.num_swaps = sum(1 for ep in self.epochs if ep.incoming)
        self.num_passes = (sum(self.engine.num_passes(c) for c in self.compiled)
                           + sum(1 for ep in self.epochs if ep.perm_src is not None))
        esz = 8 if dtype == torch.complex64 else 16
        self.swap_bytes_per_step = sum(
            (esz << nl) - ((esz << nl) >> len(ep.incoming)) for ep in self.epochs if ep.incoming)

    # ------------------------------------------------------------------------------------
    def _exchange(self, st: ShardedState, ep: Epoch):
        m = len(ep.incoming)
        nl = self.n_local
        block = 1 << (nl - m)
        src = st.local
        dst = st._spare()
        a = 0
        for j, rb in enumerate(ep.rank_bits):
            a |= ((st.rank >> rb) & 1) << j
        ops = []
        for b in range(1 << m):
            if b == a:
                continue
            peer = st.rank
            for j, rb in enumerate(ep.rank_bits):
                peer = (peer & ~(1 << rb)) | (((b >> j) & 1) << rb)
            # real views: complex dtypes are not supported by every backend (gloo)
            send = torch.view_as_real(src[b * block:(b + 1) * block])
            recv = torch.view_as_real(dst[b * block:(b + 1) * block])
            ops.append(dist.P2POp(dist.isend, send, peer, group=st.group))
            ops.append(dist.P2POp(dist.irecv, recv, peer, group=st.group))
        dst[a * block:(a + 1) * block].copy_(src[a * block:(a + 1) * block])
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        st.local, st.spare = dst, src

    def run(self, st: ShardedState, timing: Optional[dict] = None) -> ShardedState:
        """Execute the plan.  With `timing` (a dict) and a CUDA shard, per-phase device times
        (ms, summed over epochs) are accumulated into it: permute / exchange / gates."""
        if st.layout != self.start_layout:
            raise RuntimeError("state layout does not match the layout this circuit was planned for")
        if st.world != self.world or st.n != self.n:
            raise RuntimeError("state does not match the circuit's size / world")
        marks = []

        def mark(tag):
            if timing is not None and st.local.is_cuda:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((tag, ev))

        mark("start")
        for ep, comp in zip(self.epochs, self.compiled):
            if ep.perm_src is not None:
                out = self.engine.permute(ep.perm_src, st.local, st._spare())
                st.local, st.spare = out, st.local
                mark("permute")
            if ep.incoming:
                self._exchange(st, ep)
                mark("exchange")
            self.engine.run(comp, st.local)
            mark("gates")
        st.layout = list(self.end_layout)
        if marks:
            torch.cuda.synchronize()
            for (_, e0), (tag, e1) in zip(marks[:-1], marks[1:]):
                timing[tag] = timing.get(tag, 0.0) + e0.elapsed_time(e1)
        return st
