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

import os
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
    victim_bits: List[int] = field(default_factory=list)  # physical bits of the victims BEFORE perm_src
    perm_src: Optional[List[int]] = None                # local bit permutation before the exchange
    gates: List[int] = field(default_factory=list)      # gate indices executed after it
    local_bits: List[List[int]] = field(default_factory=list)  # physical bit positions per gate


def identity_layout(n: int) -> List[int]:
    """phys[q] = physical index bit of logical qubit q: qubit 0 is the most significant bit."""
    return [n - 1 - q for q in range(n)]


def plan_epochs(gate_qubits: Sequence[Sequence[int]], n: int, g: int,
                layout: Optional[List[int]] = None, restore: bool = True,
                lookahead: int = 4096, min_victim_bit: int = 0) -> Tuple[List[Epoch], List[int]]:
    """Cut a gate list into epochs for a state sharded over 2^g ranks.

    layout[q] is the physical bit of logical qubit q; bits >= n-g are rank bits.  Returns the
    epochs and the final layout.  With restore=True a last exchange puts the layout back to
    the one the plan started from, so the same plan can be replayed step after step.
    min_victim_bit: victims are only taken from physical bits >= this (the fused scatter
    exchange keeps the low bits of every tile together; restore exchanges may still have to
    evict a lower bit, those fall back to permute + send/recv).
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
        ep = Epoch(incoming=list(incoming), victims=list(victims),
                   victim_bits=[phys[v] for v in victims])
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
        # bring in as many pending global qubits as there are local qubits that may leave: the
        # partners of an incoming qubit in the gate that needs it first must stay
        all_local = [q for q in range(n) if phys[q] < n_local]
        take = min(g, len(globals_needed))
        while take > 1:
            partners = set()
            for q in globals_needed[:take]:
                partners.update(gate_qubits[remaining[first_use[q]]])
            if sum(1 for q in all_local if q not in partners) >= take:
                break
            take -= 1
        incoming = globals_needed[:take]
        # Belady: evict the local qubits whose next use is farthest away
        never = len(remaining) + 1
        # never evict the partners of an incoming qubit in the gate that needs it first (the gate
        # would stay non-local); among the rest prefer bits >= min_victim_bit, then anything
        needed_now = set()
        for q in incoming:
            needed_now.update(gate_qubits[remaining[first_use[q]]])
        local_qubits = [q for q in all_local if phys[q] >= min_victim_bit and q not in needed_now]
        if len(local_qubits) < len(incoming):
            local_qubits = [q for q in all_local if q not in needed_now]
        if len(local_qubits) < len(incoming):
            local_qubits = all_local
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


def exchange_block_id(rank: int, ep: Epoch) -> int:
    """This rank's own value of the rank bits the exchange `ep` swaps (bit j <-> rank_bits[j])."""
    a = 0
    for j, rb in enumerate(ep.rank_bits):
        a |= ((rank >> rb) & 1) << j
    return a


def exchange_peer(rank: int, ep: Epoch, b: int) -> int:
    """The rank that differs from `rank` only in the swapped rank bits, which take the value b."""
    peer = rank
    for j, rb in enumerate(ep.rank_bits):
        peer = (peer & ~(1 << rb)) | (((b >> j) & 1) << rb)
    return peer


# --------------------------------------------------------------------------- #
# local engines
# --------------------------------------------------------------------------- #
class CudaEngine:
    """Local work on the shard with the native single-GPU engine."""

    def __init__(self, dtype):
        self.dtype = dtype

    def compile(self, gates_local, n_local, tail_victims=None):
        from . import circuit
        if not gates_local:
            return None
        return circuit.CompiledCircuit(gates_local, n_local, self.dtype, tail_forbidden=tail_victims)

    def run(self, compiled, shard):
        if compiled is not None:
            compiled.run(shard, in_place=True)
        return shard

    def num_passes(self, compiled):
        return compiled.num_passes if compiled is not None else 0

    # fused scatter exchange (peer-memory stores from the last pass of an epoch)
    supports_scatter = True
    supports_scatter_mark = True

    def min_victim_bit(self, n_local, g):
        from . import circuit
        geo = circuit.default_geometry(n_local, self.dtype)
        return min(geo.low_bits, max(0, n_local - g))

    def scatter_tail(self, compiled, n_local, victim_bits):
        from . import circuit
        return circuit.ScatterTail(compiled, n_local, self.dtype, victim_bits)

    def run_scatter(self, tail, st, ep, before_scatter=None, mark=None):
        """The epoch's passes with the last one storing into the peers' spare buffers: block b of
        the exchange `ep` goes to block a (this rank's own value of the swapped rank bits) of the
        spare buffer of the rank whose swapped rank bits equal b."""
        peers = st.peer_pointers()
        m = len(ep.incoming)
        block_bytes = (1 << (st.n_local - m)) * st.local.element_size()
        a = exchange_block_id(st.rank, ep)
        spare_of = peers[st._spare().data_ptr()]
        dst = [spare_of[exchange_peer(st.rank, ep, b)] + a * block_bytes for b in range(1 << m)]
        # visit order rotated by this rank's own block number: at any moment every rank of the
        # exchange group is writing to a different peer
        tail.run(st.local, dst, before_scatter=before_scatter, visit_xor=a, mark=mark)

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

    # -- peer memory: every rank maps the two buffers of every other rank (CUDA IPC) ---------
    def peer_pointers(self):
        """{my buffer address: [address of the same buffer on rank r, as mapped into this
        process]} for the two buffers (state, spare).  Collective: all ranks call it together.
        The mapping is cached until one of the buffers is replaced."""
        from . import _lib as L
        spare = self._spare()
        key = (self.local.data_ptr(), spare.data_ptr())
        cache = getattr(self, "_peers", None)
        if cache is not None and set(cache["key"]) == set(key):
            return cache["map"]
        self.release_peers()
        mine = [L.ipc_export(ptr) for ptr in key]                   # [(handle, offset)] x 2
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        opened = {}                  # (rank, handle) -> base address of the mapping (small
        #                              buffers may share one allocation: map it once)
        table = {key[0]: [0] * self.world, key[1]: [0] * self.world}
        for r, entries in enumerate(everyone):
            for which, (handle, offset) in enumerate(entries):
                if r == self.rank:
                    table[key[which]][r] = key[which]
                else:
                    if (r, handle) not in opened:
                        opened[(r, handle)] = L.ipc_open(handle, 0)
                    table[key[which]][r] = opened[(r, handle)] + offset
        self._peers = {"key": key, "map": table, "opened": list(opened.values())}
        return table

    def release_peers(self):
        from . import _lib as L
        cache = getattr(self, "_peers", None)
        if cache is not None:
            for base in cache["opened"]:
                L.ipc_close(base, 0)
            self._peers = None

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
    """A gate list planned once for a state sharded over `world` ranks.  With restore=True the
    plan ends in the qubit layout it started from and can be replayed; with restore=False it ends
    in `end_layout` (pass that as `layout=` to the next circuit)."""

    def __init__(self, gates, num_qubits: int, dtype, world: int, engine=None,
                 layout: Optional[List[int]] = None, restore: bool = True, exchange: Optional[str] = None):
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
        # exchange mode: "p2p" = the last pass before an exchange stores its tiles straight into
        # the peers' buffers (fused compute + exchange over NVLink peer memory); "nccl" = bit
        # permutation pass + grouped send/recv
        if exchange is None:
            exchange = os.environ.get("UA_EXCHANGE", "p2p")
        self.p2p = (exchange == "p2p" and world > 1 and getattr(self.engine, "supports_scatter", False))
        min_victim = self.engine.min_victim_bit(self.n_local, g) if self.p2p else 0
        self.epochs, self.end_layout = plan_epochs([qs for qs, _ in self.gates], num_qubits, g,
                                                   self.start_layout, restore=restore,
                                                   min_victim_bit=min_victim)
        nl = self.n_local
        # scatter tails: epoch i ends with the exchange that opens epoch i+1
        want_tail = [False] * len(self.epochs)
        if self.p2p:
            for i in range(len(self.epochs) - 1):
                nxt = self.epochs[i + 1]
                want_tail[i] = bool(nxt.incoming) and min(nxt.victim_bits) >= min_victim
        self.compiled = []
        self.tails = [None] * len(self.epochs)
        for i, ep in enumerate(self.epochs):
            local_gates = [([nl - 1 - p for p in bits], self.gates[gi][1])
                           for gi, bits in zip(ep.gates, ep.local_bits)]
            if want_tail[i]:
                victims = self.epochs[i + 1].victim_bits
                try:
                    comp = self.engine.compile(local_gates, nl, tail_victims=victims)
                    self.tails[i] = self.engine.scatter_tail(comp, nl, victims)
                    self.compiled.append(comp)
                    continue
                except ValueError:
                    # no tile avoids the leaving bits (tiny shards): this exchange uses the
                    # permute + send/recv formulation, which the epoch describes as well
                    self.tails[i] = None
            self.compiled.append(self.engine.compile(local_gates, nl))
        self.num_swaps = sum(1 for ep in self.epochs if ep.incoming)
        self.num_fused_swaps = sum(1 for t in self.tails if t is not None)
        self.num_passes = 0
        for i, ep in enumerate(self.epochs):
            fused_in = i > 0 and self.tails[i - 1] is not None
            if ep.perm_src is not None and not fused_in:
                self.num_passes += 1
            self.num_passes += (self.tails[i].num_passes if self.tails[i] is not None
                                else self.engine.num_passes(self.compiled[i]))
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
        a = exchange_block_id(st.rank, ep)
        ops = []
        for b in range(1 << m):
            if b == a:
                continue
            peer = exchange_peer(st.rank, ep, b)
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

    def _fence(self, st: ShardedState):
        """All ranks' scatter passes are complete (stream-ordered: a tiny all-reduce queued behind
        each rank's pass cannot finish before every rank has reached it)."""
        if getattr(st, "_fence_buf", None) is None:
            st._fence_buf = torch.zeros(1, dtype=torch.float32, device=st.local.device)
        dist.all_reduce(st._fence_buf, group=st.group)

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
        # A scatter pass writes into the peers' spare buffers.  That is safe without further
        # synchronisation only if every peer is known to be past its last use of that buffer:
        # true right after a fenced flip (the buffer was the peer's state until its own scatter
        # pass, which completed before the fence), not at the start of a run or after a local
        # permutation / send-recv exchange -- then a fence goes in front of the scatter pass.
        spare_idle = False
        for i, (ep, comp) in enumerate(zip(self.epochs, self.compiled)):
            if i > 0 and self.tails[i - 1] is not None:
                # the previous epoch's last pass already delivered this exchange into every
                # rank's spare buffer: wait until all ranks have finished writing, then flip
                self._fence(st)
                st.local, st.spare = st.spare, st.local
                spare_idle = True
                mark("exchange")
            else:
                if ep.perm_src is not None:
                    out = self.engine.permute(ep.perm_src, st.local, st._spare())
                    st.local, st.spare = out, st.local
                    spare_idle = False
                    mark("permute")
                if ep.incoming:
                    self._exchange(st, ep)
                    spare_idle = False
                    mark("exchange")
            tail = self.tails[i]
            if tail is not None:
                if getattr(self.engine, "supports_scatter_mark", False):
                    self.engine.run_scatter(tail, st, self.epochs[i + 1],
                                            before_scatter=None if spare_idle else (lambda: self._fence(st)),
                                            mark=mark)
                    mark("scatter_pass")          # the pass whose stores cross NVLink
                else:
                    self.engine.run_scatter(tail, st, self.epochs[i + 1],
                                            before_scatter=None if spare_idle else (lambda: self._fence(st)))
                    mark("gates")
            else:
                self.engine.run(comp, st.local)
                mark("gates")
        st.layout = list(self.end_layout)
        if marks:
            torch.cuda.synchronize()
            for (_, e0), (tag, e1) in zip(marks[:-1], marks[1:]):
                timing[tag] = timing.get(tag, 0.0) + e0.elapsed_time(e1)
        return st
