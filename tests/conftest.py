"""pytest config: registers the `gpu` marker and puts the product package and the
oracle (test infrastructure) on sys.path."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "qcware-unitair_b200")
for p in (PKG_ROOT, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy access to tests/golden/*.npz + manifest.json (made by make_golden.py)."""

    def __init__(self):
        with open(os.path.join(GOLDEN, "manifest.json")) as f:
            self.manifest = json.load(f)
        self._files = {}

    def arrays(self, name):
        if name not in self._files:
            self._files[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return self._files[name]


@pytest.fixture(scope="session")
def golden():
    return Golden()


# tolerances from BASELINE.json north_star: 1e-5 relative (complex64), 1e-12 (complex128)
RTOL = {"c64": 1e-5, "c128": 1e-12,
        np.dtype("complex64"): 1e-5, np.dtype("complex128"): 1e-12,
        np.dtype("float32"): 1e-5, np.dtype("float64"): 1e-12}


def rel_err(a, b):
    """norm-wise relative error max over batch entries: |a-b|_2 / |b|_2."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    num = np.linalg.norm((a - b).reshape(-1).astype(np.complex128))
    den = np.linalg.norm(b.reshape(-1).astype(np.complex128))
    return float(num / den) if den > 0 else float(num)


def assert_close(a, b, dtype_key, factor=1.0, what=""):
    tol = RTOL[dtype_key] * factor
    err = rel_err(a, b)
    assert err <= tol, f"{what}: rel err {err:.3e} > {tol:.1e}"
