"""Sampling measurement outcomes (mirrors src/unitair/simulation/measurement.py).

`measure` keeps the reference's signature and return types (a MeasurementHistogram sorted by
count, or a plain {int: count} dict with raw_output=True).  The reference materialises
abs_squared(state), builds a torch.distributions.Categorical over it and pulls every sample
to the host (`samples.tolist()`, measurement.py:40-47); here the state is read once by
ua_sample_block_sums, the draws are located by ua_sample_locate (inverse CDF, one 32 KiB block
per sample) and only the histogram of distinct outcomes crosses to the host.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch

from .. import _engine
from .. import _lib as L
from ..states import count_qubits

BLOCK_LOG2 = 12          # amplitudes per CDF block (2^12 = 32 KiB complex64)


def sample_indices(state: torch.Tensor, num_samples: int, generator: Optional[torch.Generator] = None,
                   uniforms: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`num_samples` basis-state indices (int64, on the state's device) drawn with probability
    |state[i]|^2 / sum |state|^2.  `uniforms` (float64 in [0, 1), one per sample) replaces the
    internal random draws -- the parity tests use it to compare with the oracle's inverse CDF."""
    if state.dim() != 1:
        raise ValueError("measure expects one state in vector layout (no batch dimensions)")
    L.require_cuda(state)
    if not state.is_complex():
        state = state.to(torch.complex128 if state.dtype == torch.float64 else torch.complex64)
    st = _engine._aligned(state.detach())
    dev = st.device
    elems = st.numel()
    blog = BLOCK_LOG2
    num_blocks = (elems + (1 << blog) - 1) >> blog
    sums = torch.empty(num_blocks, dtype=torch.float64, device=dev)
    code = L.dtype_code(st.dtype)
    lib = L.lib()
    with L.on_device(dev):
        stream = L.stream_ptr(dev)
        L.check(lib.ua_sample_block_sums(code, sums.data_ptr(), st.data_ptr(), elems, blog, stream))
        cdf = torch.cumsum(sums, 0)
        # the reference's Categorical refuses a probability vector that is all zero or not finite
        # (src/unitair/simulation/measurement.py:43); sampling returns host data anyway, so the
        # synchronisation costs nothing extra
        total = float(cdf[-1])
        if not (total > 0.0) or total == float("inf"):
            raise ValueError("measure: the state has no finite, non-zero norm (sum of |amplitude|^2 = "
                             f"{total}); probabilities cannot be formed")
        if uniforms is None:
            uniforms = torch.rand(num_samples, dtype=torch.float64, device=dev, generator=generator)
        else:
            uniforms = uniforms.to(device=dev, dtype=torch.float64)
            if uniforms.numel() != num_samples:
                raise ValueError("uniforms must hold one number per sample")
        targets = (uniforms * cdf[-1]).contiguous()
        out = torch.empty(num_samples, dtype=torch.int64, device=dev)
        L.check(lib.ua_sample_locate(code, out.data_ptr(), st.data_ptr(), elems, blog, cdf.data_ptr(),
                                     targets.data_ptr(), num_samples, stream))
    return out


def measure(state: torch.Tensor, num_samples: int, raw_output: bool = False):
    """Draw samples from the probability distribution of a state in vector layout
    (src/unitair/simulation/measurement.py:9-70).  The state itself is not changed.

    Returns a MeasurementHistogram (bit strings, most frequent first) or, with raw_output=True,
    a dict {basis-state index: count}.
    """
    num_qubits = count_qubits(state)
    samples = sample_indices(state, num_samples)
    values, counts = torch.unique(samples, return_counts=True)
    if raw_output:
        return dict(zip(values.tolist(), counts.tolist()))
    order = torch.argsort(counts, descending=True, stable=True)
    histogram = MeasurementHistogram(num_qubits=num_qubits)
    for v, c in zip(values[order].tolist(), counts[order].tolist()):
        histogram[format(v, f"0{num_qubits}b")] = c
    return histogram


class MeasurementHistogram:
    """Counts per observed bit string, in the order they were inserted (most frequent first when
    produced by `measure`).  Same interface as the reference's class (measurement.py:73-140)."""

    def __init__(self, num_qubits: int, histogram: Optional[OrderedDict] = None):
        self.num_qubits = num_qubits
        self.histogram = OrderedDict() if histogram is None else histogram
        self._int_key_histogram = None

    def __getitem__(self, item: str):
        if item in self.histogram:
            return self.histogram[item]
        if len(item) == self.num_qubits and set(item) <= {"0", "1"}:
            return 0
        raise KeyError(f"Given key {item} is not a valid binary string for {self.num_qubits} bits.")

    def __setitem__(self, key, value):
        self.histogram[key] = value
        self._int_key_histogram = None

    @property
    def num_distinct_samples(self):
        return len(self.histogram)

    @property
    def observed_samples(self):
        return set(self.histogram)

    def int_key_histogram(self):
        if self._int_key_histogram is None:
            self._int_key_histogram = OrderedDict((int(k, 2), c) for k, c in self.histogram.items())
        return self._int_key_histogram

    def __repr__(self):
        items = list(self.histogram.items())
        if len(items) < 50:
            return "\n".join(f"{k}: {c}" for k, c in items)
        head = "".join(f"{k}: {c}\n" for k, c in items[:15])
        return head + f" ... ({len(items) - 15} lines omitted)"
