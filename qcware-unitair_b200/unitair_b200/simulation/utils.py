"""Gate batch-structure helpers (mirror of src/unitair/simulation/utils.py:4-31)."""
import torch


def count_gate_batch_dims(gate: torch.Tensor) -> int:
    out = gate.dim() - 2
    if out < 0:
        raise RuntimeError(
            f"Gate with size {gate.size()} is incorrectly shaped for an operator batch. "
            "Expected size is\n  (*optional_batch_dims, 2^k, 2^k)\n"
            "with k the number of qubits on which the gate acts.")
    return out


def gate_batch_size(gate: torch.Tensor):
    return gate.size()[:count_gate_batch_dims(gate)]
