"""ctypes binding of libunitair_b200.so (the C ABI declared in include/unitair_b200.h).

There is NO CPU path and no fallback: if the shared library is missing or a tensor is
not on a CUDA device the calls below raise.  Outputs and workspaces are allocated by
the caller through torch's caching allocator; kernels are enqueued on torch's current
CUDA stream of the tensor's device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_longlong, c_size_t, c_void_p, c_char_p, c_ulonglong, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# UA_LIB_PATH: a differently built library for A/B measurements (tools/); the default is the in-tree build
LIB_PATH = os.environ.get("UA_LIB_PATH") or os.path.join(_HERE, "lib", "libunitair_b200.so")

UA_C64, UA_C128 = 0, 1
UA_ERR_UNSUPPORTED = 2
MAX_GATE_QUBITS = 5
MAX_GENERIC_GATE_QUBITS = 10
MAX_FUSED_GATES = 64

_lib = None


class EngineError(RuntimeError):
    """The native engine rejected a call (bad argument, unsupported size, CUDA error)."""


def lib():
    """Load the native library once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"unitair_b200: native library not found at {LIB_PATH}. Build it with "
                "`qcware-unitair_b200/csrc/build.sh` (or `python -c 'import __graft_entry__ as g; "
                "g.build()'`). There is no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    p_int = POINTER(c_int)
    p_ll = POINTER(c_longlong)
    L.ua_version.restype = c_int
    L.ua_last_error.restype = c_char_p
    L.ua_launch_count.restype = c_ulonglong
    L.ua_apply_gate.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, p_int,
                                c_longlong, c_longlong, c_longlong, c_int, c_void_p]
    L.ua_gate_grad_workspace_bytes.restype = c_size_t
    L.ua_gate_grad_workspace_bytes.argtypes = [c_int, c_int, c_int, c_longlong, c_longlong]
    L.ua_gate_grad.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, p_int,
                               c_longlong, c_longlong, c_longlong, c_void_p, c_size_t, c_void_p]
    L.ua_apply_phase.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong,
                                 c_longlong, c_longlong, c_longlong, c_int, c_void_p]
    L.ua_phase_backward.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_longlong, c_longlong, c_longlong, c_longlong, c_longlong,
                                    c_void_p]
    L.ua_reduce_workspace_bytes.restype = c_size_t
    L.ua_reduce_workspace_bytes.argtypes = [c_longlong, c_longlong]
    L.ua_abs_squared.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_void_p]
    L.ua_norm_squared.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_longlong,
                                  c_void_p, c_size_t, c_void_p]
    L.ua_diag_expectation.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong,
                                      c_longlong, c_longlong, c_void_p, c_size_t, c_void_p]
    L.ua_inner_product.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong,
                                   c_longlong, c_longlong, c_void_p, c_size_t, c_void_p]
    L.ua_real_scale.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_longlong, c_longlong,
                                c_longlong, c_longlong, c_longlong, ctypes.c_double, c_void_p]
    L.ua_fused_limits.argtypes = [c_int, p_int, p_int]
    L.ua_apply_fused_pass.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                      p_int, c_int, p_int, p_int, p_ll, c_void_p, c_longlong,
                                      c_int, c_void_p]
    L.ua_apply_fused_pass_hostmats.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                               p_int, c_int, p_int, p_int, p_ll, c_void_p, c_int, c_void_p]
    L.ua_apply_fused_pass_scatter_hostmats.argtypes = [c_int, c_void_p, c_longlong, c_int, c_int, c_int, p_int,
                                                       c_int, p_int, p_int, p_ll, c_void_p, c_int, p_int,
                                                       POINTER(c_void_p), c_int, c_void_p]
    L.ua_fused_backward_pass.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                         p_int, c_int, p_int, p_int, p_ll, c_void_p, c_longlong,
                                         p_int, c_void_p, c_void_p]
    L.ua_apply_fused_pass_scatter.argtypes = [c_int, c_void_p, c_longlong, c_int, c_int, c_int, p_int,
                                              c_int, p_int, p_int, p_ll, c_void_p, c_int, p_int,
                                              POINTER(c_void_p), c_int, c_void_p]
    L.ua_ipc_export.argtypes = [c_void_p, c_void_p, p_ll]
    L.ua_ipc_open.argtypes = [c_void_p, c_longlong, POINTER(c_void_p)]
    L.ua_ipc_close.argtypes = [c_void_p, c_longlong]
    L.ua_sample_block_sums.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_int, c_void_p]
    L.ua_sample_locate.argtypes = [c_int, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p,
                                   c_longlong, c_void_p]
    L.ua_permute_bits.argtypes = [c_int, c_void_p, c_void_p, c_int, c_longlong, p_int, c_void_p]
    L.ua_apply_sign_masks.argtypes = [c_int, c_void_p, c_void_p, c_int, c_longlong, c_int,
                                      POINTER(c_ulonglong), c_void_p]
    for name in ("ua_apply_sign_masks", "ua_apply_gate", "ua_gate_grad", "ua_apply_phase", "ua_phase_backward",
                 "ua_abs_squared", "ua_norm_squared", "ua_diag_expectation", "ua_inner_product",
                 "ua_fused_limits", "ua_real_scale", "ua_apply_fused_pass", "ua_fused_backward_pass", "ua_permute_bits",
                 "ua_apply_fused_pass_scatter", "ua_apply_fused_pass_hostmats",
                 "ua_apply_fused_pass_scatter_hostmats", "ua_ipc_export", "ua_ipc_open", "ua_ipc_close",
                 "ua_sample_block_sums", "ua_sample_locate"):
        getattr(L, name).restype = c_int


def check(rc: int):
    if rc != 0:
        msg = lib().ua_last_error().decode("utf-8", "replace")
        raise EngineError(f"unitair_b200 native call failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().ua_launch_count())


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.complex64:
        return UA_C64
    if dtype == torch.complex128:
        return UA_C128
    raise EngineError(f"unitair_b200: unsupported dtype {dtype}")


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t.device.type != "cuda":
            raise RuntimeError(
                "unitair_b200 runs only on CUDA tensors (B200 / sm_100a): got a tensor on "
                f"'{t.device}'. There is no CPU fallback; use the reference unitair on CPU.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(
                f"Expected all tensors to be on the same device, but found {dev} and {t.device}")
    return dev


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """Make `device` current for the duration of a launch if it is not already."""

    __slots__ = ("dev", "prev")

    def __init__(self, device):
        self.dev = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def int_array(values):
    return (c_int * len(values))(*values)


def ll_array(values):
    return (c_longlong * len(values))(*values)


def ptr_array(values):
    return (c_void_p * len(values))(*values)


def ipc_export(ptr: int):
    """(64-byte handle, offset) naming the device allocation that contains `ptr`."""
    buf = ctypes.create_string_buffer(64)
    off = c_longlong(0)
    check(lib().ua_ipc_export(c_void_p(ptr), buf, ctypes.byref(off)))
    return bytes(buf.raw), int(off.value)


def ipc_open(handle: bytes, offset: int) -> int:
    out = c_void_p(0)
    buf = ctypes.create_string_buffer(handle, 64)
    check(lib().ua_ipc_open(buf, offset, ctypes.byref(out)))
    return int(out.value)


def ipc_close(ptr: int, offset: int):
    check(lib().ua_ipc_close(c_void_p(ptr), offset))


def workspace(nbytes: int, device):
    if nbytes <= 0:
        return None, 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws, ws.data_ptr()
