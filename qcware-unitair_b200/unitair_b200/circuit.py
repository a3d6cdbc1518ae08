"""Circuit-level gate application: many gates per pass over the state.

The reference applies one gate per call and pays 3-4 full-state passes for each
(src/unitair/simulation/operations.py:151-186); its only fusion is multiplying 2x2s that
hit the same qubit (apply_to_qubits, operations.py:416-503).  Here a list of gates is cut
into *passes*: a pass is a run of gates whose target bits all fit inside one shared-memory
tile (the low `low_bits` index bits plus up to `max_high` freely chosen higher bits), and
each pass costs one read + one write of the state however many gates it holds
(native entry ua_apply_fused_pass).  Gates that cannot join a pass (k > 3) go through the
single-gate kernel.

The planner is pure host integer work and is tested on CPU; the execution functions need
CUDA tensors.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import torch

from . import _engine
from . import _lib as L

MAX_FUSED_K = 3


# --------------------------------------------------------------------------- #
# planning (host only)
# --------------------------------------------------------------------------- #
@dataclass
class TileGeometry:
    total_bits: int      # index bits of one independent space (= num_qubits)
    tile_bits: int       # T
    low_bits: int        # L: always-resident contiguous low bits
    max_high: int        # H = T - L
    elem_bits: int = 0   # log2(8-byte TMA elements per amplitude): 0 complex64, 1 complex128
    max_windows: int = 5  # TMA tensor rank limit (see count_windows)
    split_low: int = 0   # register-blocked path: the first TMA dimension is exactly the 4 lowest
    #                      element bits (one 128-byte row, 128-byte shared-memory swizzle)


def count_windows(low_bits: int, high, elem_bits: int = 0, split_low: int = 0) -> int:
    """Number of TMA box dimensions needed for the tile {0..low_bits-1} U high.

    A dimension covers a run of consecutive (8-byte element) index bits, at most 8 of them
    (box extent <= 256); with split_low > 0 a new dimension starts at element bit split_low.
    Mirrors setup_tensor_maps() in csrc/ua_tile.cu: with at most 5 dimensions a tile moves with
    ONE cp.async.bulk.tensor instruction, otherwise the kernel falls back to one bulk copy per
    contiguous run (much slower to issue; the register-blocked path refuses the pass).
    """
    pos = list(range(low_bits + elem_bits)) + [h + elem_bits for h in sorted(high)]
    n = 0
    start = length = None
    for b in pos:
        if n and b == start + length and length < 8 and not (split_low and b == split_low):
            length += 1
        else:
            n += 1
            start, length = b, 1
    return n


@dataclass
class Pass:
    high: List[int] = field(default_factory=list)      # ascending bit positions >= low_bits
    gates: List[int] = field(default_factory=list)     # indices into the gate list, in order
    direct: bool = False                                # single big gate -> direct kernel


def default_geometry(num_qubits: int, dtype: torch.dtype, cluster: bool = False) -> TileGeometry:
    """Tile shape used for a state of `num_qubits` qubits.

    complex64: 2^13 amplitudes = 64 KiB per tile (3 tiles resident per SM), complex128:
    2^12 = 64 KiB.  The low 7 (c64) / 6 (c128) bits are always in the tile, which makes
    every global access a coalesced 1 KiB run.
    cluster=True: the register-blocked complex64 pass (ua_apply_fused_pass_hostmats): 2^12
    amplitudes = 32 KiB per tile, seven tile buffers per SM (three being computed on, four in
    flight to or from HBM), 128-byte swizzled rows.
    """
    if cluster and dtype == torch.complex64:
        tile = int(os.environ.get("UA_CLUSTER_TILE_BITS", 12))
        low = 7
        tile = min(tile, num_qubits)
        low = min(low, tile)
        split = 4 if low >= 4 else 0          # smaller tiles fall back to the shared-memory-matrix kernel
        return TileGeometry(num_qubits, tile, low, tile - low, 0, split_low=split)
    if dtype == torch.complex128:
        tile, low, ebits = 12, 6, 1
    else:
        tile, low, ebits = 13, 7, 0
    tile = int(os.environ.get("UA_TILE_BITS", tile))
    low = int(os.environ.get("UA_TILE_LOW_BITS", low))
    tile = min(tile, num_qubits)
    low = min(low, tile)
    return TileGeometry(num_qubits, tile, low, tile - low, ebits)


def backward_geometry(num_qubits: int, dtype: torch.dtype) -> TileGeometry:
    """Tile shape of the fused backward pass: two tiles (psi and grad) share the CTA's shared
    memory, so the tile is one bit smaller than the forward one."""
    geo = default_geometry(num_qubits, dtype)
    tile = min(geo.tile_bits, (11 if dtype == torch.complex128 else 12), num_qubits)
    low = min(geo.low_bits, tile)
    return TileGeometry(num_qubits, tile, low, tile - low, geo.elem_bits)


def _sorted_to_gate_index(bits):
    """The native kernels index a gate's rows/columns in TARGET-BIT order (bit i of the index
    <-> i-th lowest target bit position); unitair's gate index has its most significant bit on
    qubits[0].  Returns idx with gate[gi, gj] = sorted[idx[gi], idx[gj]], or None if identical."""
    k = len(bits)
    order = sorted(range(k), key=lambda j: bits[j])          # ascending bit position
    gbit = [k - 1 - order[i] for i in range(k)]              # gate-index bit of sorted bit i
    idx = [0] * (1 << k)
    for s_ in range(1 << k):
        gi = 0
        for i in range(k):
            gi |= ((s_ >> i) & 1) << gbit[i]
        idx[gi] = s_
    return None if idx == list(range(1 << k)) else idx


def plan_passes_tail_first(gate_bits: Sequence[Sequence[int]], geo: TileGeometry, forbidden_last,
                           **kw) -> List[Pass]:
    """Like plan_passes, but planned from the END of the list: the LAST pass is the greedy, full
    one (leftover gates end up in the first passes) and none of its tile bits is in
    `forbidden_last`.  Used for the passes in front of a global-qubit exchange: the last pass
    then stores straight into the peers' memory (ScatterTail), and the more gates it holds the
    more of the NVLink time they hide."""
    n = len(gate_bits)
    rev = [list(gate_bits[n - 1 - i]) for i in range(n)]
    passes = plan_passes(rev, geo, forbidden_first=forbidden_last, **kw)
    out = []
    for p in reversed(passes):
        out.append(Pass(high=p.high, gates=[n - 1 - g for g in reversed(p.gates)], direct=p.direct))
    return out


def plan_passes(gate_bits: Sequence[Sequence[int]], geo: TileGeometry,
                max_gates: int = L.MAX_FUSED_GATES, max_mat_elems: int = 2048,
                lookahead: int = 512, max_fused_k: int = MAX_FUSED_K,
                forbidden_first=None) -> List[Pass]:
    """Cut an ordered gate list into passes.

    gate_bits[g] are the index-bit positions gate g acts on.  Gates are only reordered
    across gates they share no bit with (a skipped gate blocks its bits for the rest of
    the pass), so the product of the passes equals the original circuit.
    forbidden_first: bit positions that must not be tile bits of the FIRST pass.
    """
    forbidden = set(forbidden_first or ())
    n_gates = len(gate_bits)
    done = [False] * n_gates
    passes: List[Pass] = []
    first = 0
    while first < n_gates:
        if done[first]:
            first += 1
            continue
        k0 = len(gate_bits[first])
        banned = forbidden if not passes else set()
        if k0 > max_fused_k:
            passes.append(Pass(high=[], gates=[first], direct=True))
            done[first] = True
            continue
        cur = Pass()
        high = set()
        blocked = set(banned)
        mat_elems = 0
        scanned = 0
        for g in range(first, n_gates):
            if done[g]:
                continue
            scanned += 1
            if scanned > lookahead or len(cur.gates) >= max_gates:
                break
            bits = gate_bits[g]
            k = len(bits)
            if k > max_fused_k or any(b in blocked for b in bits):
                blocked.update(bits)
                continue
            need = {b for b in bits if b >= geo.low_bits} - high
            if (len(high) + len(need) > geo.max_high or mat_elems + 4 ** k > max_mat_elems
                    or (need and count_windows(geo.low_bits, high | need, geo.elem_bits, geo.split_low) > geo.max_windows)):
                blocked.update(bits)
                continue
            high |= need
            mat_elems += 4 ** k
            cur.gates.append(g)
            done[g] = True
            if len(blocked) >= geo.total_bits:
                break
        if not cur.gates:
            if banned:
                # nothing fits under the ban (the first gate touches a forbidden bit): lift the
                # ban, the caller copes (extra copy pass)
                forbidden = set()
                passes.append(None)
                continue
            # no ban and the first pending gate still does not fit a tile (tiny max_high, too
            # many TMA windows, a matrix budget below one gate): run it through the single-gate
            # kernel instead of spinning
            passes.append(Pass(high=[], gates=[first], direct=True))
            done[first] = True
            continue
        # fill the unused high slots so the tile is full: first positions that keep the number
        # of TMA dimensions (extend an existing run), then anything
        while len(high) < geo.max_high:
            free = [p for p in range(geo.low_bits, geo.total_bits) if p not in high and p not in banned]
            if not free:
                break
            best = min(free, key=lambda p: (count_windows(geo.low_bits, high | {p}, geo.elem_bits, geo.split_low), p))
            high.add(best)
        cur.high = sorted(high)
        passes.append(cur)
    return [p for p in passes if p is not None]


# --------------------------------------------------------------------------- #
# gate merging (host bookkeeping + tiny device matmuls, no synchronisation)
# --------------------------------------------------------------------------- #
def _lift(qs, m, target):
    """Matrix of the gate (qs, m) as a 2^K x 2^K operator on the ordered qubit list `target`
    (a superset of qs): m (x) identity on the other qubits, re-indexed so that the most
    significant index bit is target[0] (the reference's convention, operations.py:82-86)."""
    qs, target = list(qs), list(target)
    if qs == target:
        return m
    K, k = len(target), len(qs)
    rest = [q for q in target if q not in qs]
    full = m if not rest else _bkron(m, torch.eye(1 << len(rest), dtype=m.dtype, device=m.device))
    order = qs + rest                      # qubit order of `full`, most significant first
    if order == target:
        return full
    pos = {q: K - 1 - i for i, q in enumerate(order)}     # bit of q in full's index
    idx = []
    for i in range(1 << K):
        j = 0
        for t, q in enumerate(target):
            j |= ((i >> (K - 1 - t)) & 1) << pos[q]
        idx.append(j)
    it = torch.tensor(idx, device=m.device)
    return full.index_select(-2, it).index_select(-1, it)


def _bkron(a, b):
    """Kronecker product over the last two dims with broadcasting batch dims."""
    a4 = a.unsqueeze(-1).unsqueeze(-3)          # (..., i, 1, j, 1)
    b4 = b.unsqueeze(-2).unsqueeze(-4)          # (..., 1, k, 1, l)
    out = a4 * b4
    return out.reshape(out.shape[:-4] + (out.shape[-4] * out.shape[-3], out.shape[-2] * out.shape[-1]))


SMALL_STATE_AMPS = 1 << 21     # below this, device-resident gates stay on the device (no host copy)


def use_cluster_path() -> bool:
    """The register-blocked pass kernel (matrix values as kernel parameters) is the default for
    complex64 circuits of shared 1-/2-qubit gates; UA_CLUSTER=0 keeps every pass on the
    shared-memory-matrix kernel (A/B measurements)."""
    return os.environ.get("UA_CLUSTER", "1") != "0"


def default_merge_k() -> int:
    return int(os.environ.get("UA_MERGE_MAX_K", "2"))


def merge_gates(gates, max_k: int = None):
    """Merge neighbouring 1-/2-qubit gates into blocks of at most `max_k` qubits.

    * a gate whose qubits are all covered by the most recent block on those qubits is
      multiplied into that block where it stands;
    * otherwise the gate is fused with the most recent blocks on its qubits when those blocks
      can be moved next to it (each is the LAST block on every one of its qubits) and the union
      has at most max_k qubits; blocks that cannot join stay where they are.
    Fusing two 2-qubit blocks that share a qubit into one 3-qubit block costs the same flops
    (8 complex MACs per amplitude either way) and halves the shared-memory round trips of the
    fused pass.  Gates on more than max_k qubits are kept as they are.  Only gates acting on
    disjoint qubits are commuted, so the product is unchanged.  Returns a new
    [(qubits, matrix)] list.
    """
    if max_k is None:
        max_k = default_merge_k()
    max_k = max(2, int(max_k))
    blocks = []          # [qubits, matrix] or None when absorbed
    last = {}            # qubit -> index of the most recent block touching it
    for qs, m in gates:
        qs = list(qs)
        k = len(qs)
        if k > max_k:
            blocks.append([qs, m])
            for q in qs:
                last[q] = len(blocks) - 1
            continue
        owners = []
        for q in qs:
            bi = last.get(q)
            if bi is not None and bi not in owners:
                owners.append(bi)
        # (1) covered by one block that is the most recent on all of the gate's qubits
        if len(owners) == 1 and all(last.get(q) == owners[0] for q in qs):
            bq, bm = blocks[owners[0]]
            if len(bq) <= max_k and set(qs) <= set(bq):
                blocks[owners[0]][1] = torch.matmul(_lift(qs, m, bq), bm)
                continue
        # (2) pull movable owner blocks into a new block at the end, smallest first
        movable = [bi for bi in owners
                   if len(blocks[bi][0]) <= max_k and all(last.get(q) == bi for q in blocks[bi][0])]
        movable.sort(key=lambda bi: (len(set(blocks[bi][0]) - set(qs)), bi))
        union = list(qs)
        taken = []
        for bi in movable:
            extra = [q for q in blocks[bi][0] if q not in union]
            if len(union) + len(extra) <= max_k:
                union += extra
                taken.append(bi)
        mat = _lift(qs, m, union)
        for bi in taken:         # the taken blocks act on disjoint qubit sets: any order
            mat = torch.matmul(mat, _lift(blocks[bi][0], blocks[bi][1], union))
            blocks[bi] = None
        blocks.append([union, mat])
        for q in union:
            last[q] = len(blocks) - 1
    return [(b[0], b[1]) for b in blocks if b is not None]


# --------------------------------------------------------------------------- #
# execution
# --------------------------------------------------------------------------- #
class _PassLaunch:
    """Pre-marshalled arguments of one native call (everything except the state pointers)."""

    __slots__ = ("direct", "gate", "low", "nhigh", "high", "ngates", "ks", "bits", "offs", "host_ok")

    def __init__(self, p: Pass, geo: TileGeometry, gate_bits, offsets):
        self.direct = p.direct
        self.gate = p.gates[0] if p.direct else -1
        self.host_ok = False
        if p.direct:
            return
        self.low = geo.tile_bits - len(p.high)
        self.nhigh = len(p.high)
        self.high = L.int_array(p.high) if p.high else None
        self.ngates = len(p.gates)
        if not p.gates:                      # pure copy pass (scatter tail without gates)
            self.ks = self.bits = self.offs = None
            return
        self.ks = L.int_array([len(gate_bits[g]) for g in p.gates])
        flat = []
        for g in p.gates:
            b = list(gate_bits[g])
            flat += b + [0] * (3 - len(b))
        self.bits = L.int_array(flat)
        self.offs = L.ll_array([offsets[g] for g in p.gates])
        # register-blocked path (matrices as kernel parameters): 1-/2-qubit gates, tile >= 4 bits
        self.host_ok = geo.tile_bits >= 4 and all(len(gate_bits[g]) <= 2 for g in p.gates)


def _pack_gates(mats_list, batch_shape):
    """One flat device buffer holding every gate matrix; returns (buffer, offsets, row_stride)."""
    offsets = []
    off = 0
    for m in mats_list:
        offsets.append(off)
        off += m.shape[-1] * m.shape[-2]
    if not any(m.dim() > 2 for m in mats_list):
        return torch.cat([m.reshape(-1) for m in mats_list]).contiguous(), offsets, 0
    rows = _engine._prod(batch_shape)
    parts = []
    for m in mats_list:
        if m.dim() == 2:
            parts.append(m.reshape(1, -1).expand(rows, -1))
        else:
            parts.append(m.expand(tuple(batch_shape) + m.shape[-2:]).reshape(rows, -1))
    return torch.cat(parts, dim=1).contiguous(), offsets, off


class CompiledCircuit:
    """An ordered gate list, planned into passes once and replayable on any state of the
    same shape: `run` only enqueues the native pass kernels (no host planning, no packing).

    gates: sequence of (qubits, operator) with operator (2^k, 2^k) or batch_shape + (2^k, 2^k),
    complex CUDA tensors of `dtype`.  Gate values are captured at compile time.
    """

    def __init__(self, gates, num_qubits: int, dtype: torch.dtype, batch_shape=(),
                 geometry: TileGeometry = None, merge: bool = True, tail_forbidden=None):
        from . import states
        n = num_qubits
        self.n = n
        self.dtype = dtype
        self.batch_shape = tuple(batch_shape)
        self.batch = _engine._prod(self.batch_shape)
        self.gates = [([int(q) for q in qs], m) for qs, m in gates]
        self.num_source_gates = len(self.gates)
        for qs, m in self.gates:
            k = states.count_qubits_gate_matrix(m)
            if len(qs) != k or len(set(qs)) != k or not set(qs).issubset(range(n)):
                raise ValueError(f"qubits={qs} is not a valid target list for a {k}-qubit "
                                 f"operator on {n} qubits")
            if m.dtype != dtype:
                raise RuntimeError(f"expected scalar type {dtype} but found {m.dtype}")
            if m.dim() != 2 and tuple(m.shape[:-2]) != self.batch_shape:
                raise RuntimeError(f"operator batch dims {tuple(m.shape[:-2])} do not match the "
                                   f"state batch dims {self.batch_shape}")
        # gate tensors may all live on the host (merged there, no device synchronisation at all)
        # or all on one CUDA device (merged there; the register-blocked path then needs ONE
        # device-to-host copy of the merged matrices at compile time)
        self.gates_on_host = bool(self.gates) and all(m.device.type == "cpu" for _, m in self.gates)
        if self.gates and not self.gates_on_host:
            L.require_cuda(*[m for _, m in self.gates])
        # merging costs a handful of tiny launches per gate when the gates live on the device: more
        # than the gate phase of a small state saves (a pass holds up to 64 gates either way)
        small_dev = bool(self.gates) and not self.gates_on_host and (self.batch << n) < SMALL_STATE_AMPS
        if merge and not small_dev and os.environ.get("UA_MERGE_GATES", "1") != "0":
            with torch.no_grad():
                self.gates = merge_gates(self.gates)
        # register-blocked path: complex64, shared (un-batched) gates, matrix values on the host
        self.cluster = (dtype == torch.complex64 and use_cluster_path() and bool(self.gates)
                        and all(m.dim() == 2 for _, m in self.gates)
                        and any(len(qs) <= 2 for qs, _ in self.gates))
        self.geo = geometry or default_geometry(n, dtype, cluster=self.cluster)
        self.gate_bits = [[n - 1 - q for q in qs] for qs, _ in self.gates]
        if not self.gates:
            self.passes = []
        elif tail_forbidden:
            # index bits that leave the shard right after this circuit (sharded.py): plan from
            # the end so that the last pass is full and avoids them
            self.passes = plan_passes_tail_first(self.gate_bits, self.geo, list(tail_forbidden))
        else:
            self.passes = plan_passes(self.gate_bits, self.geo)
        if self.gates:
            self.mats, self.offsets, self.row_stride = _pack_gates([m for _, m in self.gates],
                                                                  self.batch_shape)
        self.launches = [_PassLaunch(p, self.geo, self.gate_bits, self.offsets) for p in self.passes]
        self._direct = {}
        self._dev_mats = {}         # device -> packed matrices there (fallback kernel, direct launches)
        # matrix VALUES on the host feed the register-blocked pass kernel (they travel as kernel
        # parameters): complex64, shared gates only
        self.mats_host = None
        if self.cluster and self.row_stride == 0 and any(pl.host_ok for pl in self.launches):
            # device gates cost one device-to-host copy (a synchronisation): not worth it for a
            # small state, whose passes take microseconds on either kernel
            if self.gates_on_host or (self.batch << n) >= SMALL_STATE_AMPS:
                self.mats_host = self.mats if self.gates_on_host else self.mats.cpu()
                self._mats_host_ptr = self.mats_host.data_ptr()

    def _device_mats(self, dev):
        """Packed matrices on `dev` (uploaded once when the gates were given on the host)."""
        m = self._dev_mats.get(dev)
        if m is None:
            m = self.mats if self.mats.device == dev else self.mats.to(dev)
            self._dev_mats[dev] = m
        return m

    def _direct_args(self, gate, dev):
        key = (gate, dev)
        d = self._direct.get(key)
        if d is None:
            qs, m = self.gates[gate]
            d = (_engine._aligned(m.to(dev)), L.int_array(qs), len(qs), 0 if m.dim() == 2 else 4 ** len(qs))
            self._direct[key] = d
        return d

    def capture_graph(self, state: torch.Tensor):
        """Capture `run(state, in_place=True)` into a CUDA graph bound to `state`'s buffer.

        Small states make a deep circuit launch-bound (a pass over 2^16 amplitudes takes a few
        microseconds, a launch from Python ~10); replaying the graph removes the host from the
        loop.  Returns a torch.cuda.CUDAGraph; every `.replay()` applies the circuit once more
        to the same buffer.  The state must be contiguous and 16-byte aligned.
        """
        if _engine._aligned(state).data_ptr() != state.data_ptr():
            raise RuntimeError("capture_graph needs a contiguous, 16-byte aligned state")
        # warm up outside the capture (lazy kernel attributes, tensor-map encoder lookup)
        self.run(state, in_place=True)
        torch.cuda.synchronize(state.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.run(state, in_place=True)
        return graph

    @property
    def num_gates(self):
        return len(self.gates)

    @property
    def num_passes(self):
        return len(self.launches)

    def run(self, state: torch.Tensor, in_place: bool = False) -> torch.Tensor:
        """Apply the circuit.  Out-of-place by default (the reference's semantics); with
        in_place=True the (contiguous, aligned) state buffer is overwritten."""
        n = self.n
        if state.dtype != self.dtype or tuple(state.shape) != self.batch_shape + (1 << n,):
            raise RuntimeError(f"state of size {tuple(state.shape)} / {state.dtype} does not match "
                               f"the compiled circuit ({self.batch_shape + (1 << n,)}, {self.dtype})")
        L.require_cuda(state)
        cur = _engine._aligned(state)
        if in_place and cur.data_ptr() != state.data_ptr():
            raise RuntimeError("in_place=True needs a contiguous, 16-byte aligned state")
        if not self.launches or self.batch == 0:
            return cur if in_place else cur.clone()
        out = cur if in_place else torch.empty_like(cur)
        self._run_launches(self.launches, cur, out)
        return out

    def _run_launches(self, launches, src, out):
        """Enqueue `launches`: the first reads `src`, all write (and later ones read) `out`."""
        n = self.n
        dev = out.device
        lib = L.lib()
        code = L.dtype_code(self.dtype)
        total = self.batch << n
        with L.on_device(dev):
            stream = L.stream_ptr(dev)
            mats_ptr = None
            for pl in launches:
                if pl.direct:
                    m, qarr, k, gstride = self._direct_args(pl.gate, dev)
                    dst = out
                    if k > L.MAX_GATE_QUBITS and src.data_ptr() == out.data_ptr():
                        dst = torch.empty_like(out)      # the generic 6..10-qubit kernel is out-of-place only
                    L.check(lib.ua_apply_gate(code, dst.data_ptr(), src.data_ptr(), m.data_ptr(), n, k,
                                              qarr, self.batch, 1 << n, gstride, 0, stream))
                    if dst is not out:
                        out.copy_(dst)
                    src = out
                    continue
                if pl.host_ok and self.mats_host is not None:
                    rc = lib.ua_apply_fused_pass_hostmats(
                        code, out.data_ptr(), src.data_ptr(), total, n, pl.low, pl.nhigh, pl.high,
                        pl.ngates, pl.ks, pl.bits, pl.offs, self._mats_host_ptr, 0, stream)
                    if rc == 0:
                        src = out
                        continue
                    if rc != L.UA_ERR_UNSUPPORTED:
                        L.check(rc)
                    pl.host_ok = False           # e.g. the tile needs > 5 TMA dimensions
                if mats_ptr is None:
                    mats_ptr = self._device_mats(dev).data_ptr()
                L.check(lib.ua_apply_fused_pass(
                    code, out.data_ptr(), src.data_ptr(), total, n, pl.low, pl.nhigh, pl.high,
                    pl.ngates, pl.ks, pl.bits, pl.offs, mats_ptr, self.row_stride, 0, stream))
                src = out


def _fill_high(high, geo: TileGeometry, forbidden=()):
    """Complete `high` to geo.max_high positions, preferring positions that keep the number of
    TMA windows low; positions in `forbidden` are never used.  Returns None if impossible."""
    high = set(high)
    while len(high) < geo.max_high:
        free = [p for p in range(geo.low_bits, geo.total_bits) if p not in high and p not in forbidden]
        if not free:
            return None
        best = min(free, key=lambda p: (count_windows(geo.low_bits, high | {p}, geo.elem_bits, geo.split_low), p))
        high.add(best)
    return sorted(high)


class ScatterTail:
    """The last pass of a compiled circuit re-planned so that its tiles avoid the index bits
    that are about to leave the shard (`victim_bits`): its output can then be stored straight
    into the peers' buffers (ua_apply_fused_pass_scatter).  If the circuit's own last pass
    cannot be used (it touches a victim bit, is a direct big-gate launch, or there is no
    circuit) the tail is a pure copy pass."""

    def __init__(self, compiled, n: int, dtype, victim_bits):
        vs = sorted(int(v) for v in victim_bits)
        m = len(vs)
        base = default_geometry(n, dtype)
        self.victims = L.int_array(vs)
        self.m = m
        self.n = n
        self.dtype = dtype
        self.compiled = compiled
        self.inplace_launches = list(compiled.launches) if compiled is not None else []
        self.reused = False
        launch = None
        if compiled is not None and compiled.launches and not compiled.passes[-1].direct \
                and compiled.geo.tile_bits <= n - m and compiled.batch == 1 and compiled.row_stride == 0:
            last = compiled.passes[-1]
            used = {b for g in last.gates for b in compiled.gate_bits[g]}
            if not (used & set(vs)) and min(vs) >= compiled.geo.low_bits:
                high = _fill_high({b for b in used if b >= compiled.geo.low_bits}, compiled.geo, vs)
                if high is not None:
                    launch = _PassLaunch(Pass(high=high, gates=list(last.gates)), compiled.geo,
                                         compiled.gate_bits, compiled.offsets)
                    self.inplace_launches = self.inplace_launches[:-1]
                    self.reused = True
        if launch is None:
            tile = min(base.tile_bits, n - m)
            low = min(base.low_bits, tile)
            if min(vs) < low:
                raise ValueError(f"scatter bits {vs} must not be among the low {low} index bits")
            geo = TileGeometry(n, tile, low, tile - low, base.elem_bits)
            high = _fill_high(set(), geo, vs)
            if high is None:
                raise ValueError("no room for a scatter tile")
            launch = _PassLaunch(Pass(high=high, gates=[]), geo, [], [])
        self.launch = launch

    @property
    def num_passes(self):
        return len(self.inplace_launches) + 1

    def run(self, state: torch.Tensor, dst_ptrs, before_scatter=None, visit_xor: int = 0, mark=None):
        """In-place passes on `state`, then the tail pass from `state` into the 2^m destination
        blocks `dst_ptrs` (device addresses, possibly of peer GPUs).  `before_scatter` is called
        right before the scatter pass is enqueued (cross-rank fence when the destinations may
        still be in use).  visit_xor: see ua_apply_fused_pass_scatter (rank-dependent tile order).
        mark: optional callback(tag) for phase timing, called after the in-place passes."""
        cc = self.compiled
        if self.inplace_launches:
            cc._run_launches(self.inplace_launches, state, state)
        if before_scatter is not None:
            before_scatter()
        if mark is not None:
            mark("gates")
        pl = self.launch
        dev = state.device
        with L.on_device(dev):
            if pl.ngates and pl.host_ok and cc.mats_host is not None:
                rc = L.lib().ua_apply_fused_pass_scatter_hostmats(
                    L.dtype_code(self.dtype), state.data_ptr(), 1 << self.n, self.n, pl.low, pl.nhigh, pl.high,
                    pl.ngates, pl.ks, pl.bits, pl.offs, cc._mats_host_ptr,
                    self.m, self.victims, L.ptr_array(list(dst_ptrs)), int(visit_xor), L.stream_ptr(dev))
                if rc == 0:
                    return
                if rc != L.UA_ERR_UNSUPPORTED:
                    L.check(rc)
                pl.host_ok = False
            L.check(L.lib().ua_apply_fused_pass_scatter(
                L.dtype_code(self.dtype), state.data_ptr(), 1 << self.n, self.n, pl.low, pl.nhigh, pl.high,
                pl.ngates, pl.ks, pl.bits, pl.offs, cc._device_mats(dev).data_ptr() if pl.ngates else None,
                self.m, self.victims, L.ptr_array(list(dst_ptrs)), int(visit_xor), L.stream_ptr(dev)))


class _AdjointCircuit(torch.autograd.Function):
    """Whole circuit as ONE autograd node with the adjoint method (unitary gates only).

    forward: fused passes, nothing saved but the output state and the gate tensors.
    backward: walk the gates in reverse; psi <- U^H psi recomputes the input of each gate,
    grad_U = g psi^H (ua_gate_grad), g <- U^H g.  Memory: 2 states instead of one saved state
    per parameterised gate (torch's tape through the reference needs 1.25 TiB for config C3,
    SURVEY.md 7 hard part 4); cost: 3 passes per gate.
    """

    @staticmethod
    def forward(ctx, state, n, qubit_lists, *mats):
        gates = list(zip(qubit_lists, mats))
        batch_shape = tuple(state.shape[:-1])
        with torch.no_grad():
            out = CompiledCircuit(gates, n, state.dtype, batch_shape).run(state)
        ctx.n = n
        ctx.qubit_lists = qubit_lists
        ctx.save_for_backward(out, *mats)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        out, *mats = ctx.saved_tensors
        n = ctx.n
        dim = 1 << n
        batch = out.numel() >> n
        batch_shape = tuple(out.shape[:-1])
        dev = out.device
        psi = out.clone()
        g = _engine._aligned(grad_out).clone()
        grads = [None] * len(mats)
        qls = ctx.qubit_lists
        need = [bool(ctx.needs_input_grad[3 + i]) for i in range(len(mats))]

        def one_gate(i):
            qs, m = qls[i], _engine._aligned(mats[i])
            k = len(qs)
            gstride = 0 if m.dim() == 2 else 4 ** k
            _engine.launch_gate(psi, psi, m, n, k, qs, batch, dim, gstride, True)     # psi_in
            if need[i]:
                grads[i] = _engine.launch_gate_grad(g, psi, n, k, qs, batch, dim, gstride, m.shape)
            _engine.launch_gate(g, g, m, n, k, qs, batch, dim, gstride, True)         # grad wrt psi_in

        if os.environ.get("UA_FUSED_BACKWARD", "1") == "0":
            for i in range(len(mats) - 1, -1, -1):
                one_gate(i)
        else:
            # fused backward passes: psi and g tiles staged together, all 1-/2-qubit gates of a
            # pass handled in shared memory (ua_fused_backward_pass)
            geo = backward_geometry(n, out.dtype)
            gate_bits = [[n - 1 - q for q in qs] for qs in qls]
            passes = plan_passes(gate_bits, geo, max_fused_k=2)
            packed, offsets, row_stride = _pack_gates([_engine._aligned(m) for m in mats], batch_shape)
            lib = L.lib()
            code = L.dtype_code(out.dtype)
            rows = batch if row_stride else 1
            with L.on_device(dev):
                stream = L.stream_ptr(dev)
                for p in reversed(passes):
                    if p.direct:
                        one_gate(p.gates[0])
                        continue
                    pl = _PassLaunch(p, geo, gate_bits, offsets)
                    ks = [len(gate_bits[gi]) for gi in p.gates]
                    elems = sum(4 ** k for k in ks)
                    acc = torch.zeros((rows, elems, 2), dtype=torch.float64, device=dev)
                    flags = L.int_array([1 if need[gi] else 0 for gi in p.gates])
                    L.check(lib.ua_fused_backward_pass(
                        code, psi.data_ptr(), g.data_ptr(), batch << n, n, pl.low, pl.nhigh, pl.high,
                        pl.ngates, pl.ks, pl.bits, pl.offs, packed.data_ptr(), row_stride, flags,
                        acc.data_ptr(), stream))
                    if any(need[gi] for gi in p.gates):
                        accc = torch.view_as_complex(acc)            # (rows, elems) complex128
                        off = 0
                        for gi, k in zip(p.gates, ks):
                            d = 1 << k
                            if need[gi]:
                                blk = accc[:, off:off + d * d].reshape(rows, d, d)
                                idx = _sorted_to_gate_index(gate_bits[gi])
                                if idx is not None:
                                    it = torch.tensor(idx, device=dev)
                                    blk = blk.index_select(-2, it).index_select(-1, it)
                                gm = mats[gi]
                                if gm.dim() == 2 and rows > 1:
                                    blk = blk.sum(0)      # a shared gate next to per-entry gates: sum over the batch
                                grads[gi] = blk.to(gm.dtype).reshape(gm.shape)
                            off += d * d
        grad_state = g if ctx.needs_input_grad[0] else None
        return (grad_state, None, None, *grads)


def apply_gates(gates: Sequence[Tuple[Sequence[int], torch.Tensor]], state: torch.Tensor,
                in_place: bool = False, assume_unitary: bool = False) -> torch.Tensor:
    """Apply an ordered list of (qubits, operator) to a state in vector layout.

    Equivalent to calling simulation.apply_operator for each gate in turn.  Operators are
    (2^k, 2^k) or share the state's batch dims; they live on the state's device or (all of them,
    no autograd) on the host.  When no gradient is needed the list is
    executed as fused shared-memory passes.  With autograd it runs one native gate kernel and
    one autograd node per gate (one saved state per gate that needs a gradient, like torch's
    tape through the reference) -- unless assume_unitary=True, which differentiates the whole
    list with the adjoint method: fused forward, no saved intermediate states.
    """
    from . import states
    from .simulation import operations as ops
    n = states.count_qubits(state)
    gates = [([int(q) for q in qs], m) for qs, m in gates]
    needs_grad = torch.is_grad_enabled() and (
        state.requires_grad or any(m.requires_grad for _, m in gates))
    batch_shape = tuple(state.shape[:-1])
    simple = all(m.dim() == 2 or tuple(m.shape[:-2]) == batch_shape for _, m in gates)
    uniform = simple and state.is_complex() and all(m.dtype == state.dtype for _, m in gates)
    if needs_grad and assume_unitary and uniform and not in_place and gates and all(len(qs) <= L.MAX_GATE_QUBITS for qs, _ in gates):
        L.require_cuda(state, *[m for _, m in gates])
        for qs, m in gates:
            k = states.count_qubits_gate_matrix(m)
            if len(qs) != k or len(set(qs)) != k or not set(qs).issubset(range(n)):
                raise ValueError(f"qubits={qs} is not a valid target list for a {k}-qubit operator "
                                 f"on {n} qubits")
        st = _engine._aligned(state)
        return _AdjointCircuit.apply(st, n, [qs for qs, _ in gates], *[m for _, m in gates])
    if needs_grad or not uniform:
        if in_place:
            raise RuntimeError("in_place=True is not available with autograd or broadcasting gates")
        out = state
        for qs, m in gates:
            out = ops.apply_operator(m, qs, out)
        return out
    if gates and all(m.device.type == "cpu" for _, m in gates):
        # host operators (an extension of the reference's contract): merged on the host, their
        # values go to the pass kernel as launch parameters -- no upload, no synchronisation
        L.require_cuda(state)
    else:
        L.require_cuda(state, *[m for _, m in gates])
    return CompiledCircuit(gates, n, state.dtype, batch_shape).run(state, in_place=in_place)


_ALL_QUBITS_PLANS = {}


def _all_qubits_plan(n: int, dtype: torch.dtype):
    """Passes for 'one 1-qubit gate on every qubit' depend only on (n, dtype, tile settings):
    plan once, replay with whatever matrix the call brings (all gates share offset 0)."""
    geo = default_geometry(n, dtype)
    key = (n, dtype, geo.tile_bits, geo.low_bits)
    plan = _ALL_QUBITS_PLANS.get(key)
    if plan is None:
        gate_bits = [[n - 1 - q] for q in range(n)]
        passes = plan_passes(gate_bits, geo)
        plan = [_PassLaunch(p, geo, gate_bits, [0] * n) for p in passes]
        _ALL_QUBITS_PLANS[key] = plan
    return plan


def apply_same_gate_all_qubits(operator: torch.Tensor, state: torch.Tensor, n: int) -> torch.Tensor:
    """The same 2x2 (shared or per batch entry) on every qubit, qubit 0 first
    (src/unitair/simulation/operations.py:369-413)."""
    op_batch = tuple(operator.shape[:-2])
    st_batch = tuple(state.shape[:-1])
    no_grad = not (torch.is_grad_enabled() and (operator.requires_grad or state.requires_grad))
    if no_grad and state.is_complex() and operator.dtype == state.dtype and (not op_batch or op_batch == st_batch) \
            and state.numel() > 0:
        # fast path: cached plan, the operator itself is the matrix buffer (no packing, no merging)
        cur = _engine._aligned(state)
        mats = _engine._aligned(operator)
        out = torch.empty_like(cur)
        batch = _engine._prod(st_batch)
        dev = cur.device
        lib = L.lib()
        code = L.dtype_code(cur.dtype)
        src = cur
        with L.on_device(dev):
            stream = L.stream_ptr(dev)
            for pl in _all_qubits_plan(n, cur.dtype):
                L.check(lib.ua_apply_fused_pass(
                    code, out.data_ptr(), src.data_ptr(), batch << n, n, pl.low, pl.nhigh, pl.high,
                    pl.ngates, pl.ks, pl.bits, pl.offs, mats.data_ptr(), 4 if op_batch else 0, 0, stream))
                src = out
        return out
    if op_batch and not st_batch:
        state = state.expand(op_batch + state.shape[-1:])
    elif op_batch and op_batch != st_batch:
        out_batch = tuple(torch.broadcast_shapes(op_batch, st_batch))
        if len(op_batch) > len(st_batch):
            raise RuntimeError(f"operator batch dims {op_batch} cannot be broadcast to state "
                               f"batch dims {st_batch}")
        operator = operator.expand(out_batch + (2, 2))
        state = state.expand(out_batch + state.shape[-1:])
    return apply_gates([([q], operator) for q in range(n)], state)


class _PermuteQubits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, permutation, n):
        ctx.inverse = [0] * n
        for slot, item in enumerate(permutation):
            ctx.inverse[item] = slot
        ctx.n = n
        return _permute_launch(state, permutation, n)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        return _permute_launch(_engine._aligned(grad), ctx.inverse, ctx.n), None, None


def _permute_launch(state, permutation, n):
    dev = state.device
    out = torch.empty_like(state)
    batch = _engine._prod(state.shape[:-1])
    if batch == 0:
        return out
    # output qubit i takes input qubit permutation[i]; qubit q is index bit n-1-q
    src = [0] * n
    for i, p in enumerate(permutation):
        src[n - 1 - i] = n - 1 - p
    with L.on_device(dev):
        L.check(L.lib().ua_permute_bits(L.dtype_code(state.dtype), out.data_ptr(), state.data_ptr(),
                                        n, batch, L.int_array(src), L.stream_ptr(dev)))
    return out


def permute_qubits_native(permutation, state, n):
    if not state.is_complex():
        # pure data movement: view real data as complex pairs is not possible for odd
        # layouts, so use the stock permute (same device)
        from . import states
        t = states.to_tensor_layout(state)
        nb = t.dim() - n
        return states.to_vector_layout(
            t.permute(list(range(nb)) + [nb + p for p in permutation]), n)
    st = _engine._aligned(state)
    if n == 0:
        return st.clone()
    if torch.is_grad_enabled() and st.requires_grad:
        return _PermuteQubits.apply(st, list(permutation), n)
    return _permute_launch(st, list(permutation), n)
