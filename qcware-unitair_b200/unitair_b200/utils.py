"""Host-side list/permutation helpers (mirror of src/unitair/utils.py:5-49)."""
from typing import List, Sequence


def permutation_to_front(n: int, entries: Sequence[int]) -> List[int]:
    """Permutation of range(n) with `entries` first, in the given order.

    Same contract as the reference (src/unitair/utils.py:5-27), including the
    ValueError on an entry that repeats or is out of range.
    """
    entries = list(entries)
    rest = list(range(n))
    for q in entries:
        if q not in rest:
            raise ValueError(f"{q} is not in list")
        rest.remove(q)
    return entries + rest


def inverse_list_permutation(perm: Sequence[int]) -> List[int]:
    """Inverse of a permutation given as a list (src/unitair/utils.py:30-49)."""
    inv = [None] * len(perm)
    for slot, item in enumerate(perm):
        inv[item] = slot
    return inv
