"""ctypes binding of ``libsplinegrids_b200.so`` (the C ABI declared in ``include/splinegrids_b200.h``).

This is exactly what the Julia ``ccall`` shim marshals: raw device
pointers, explicit sizes, a ``cudaStream_t`` and an ``int`` status.  There is NO fallback: if the
shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Sequence

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsplinegrids_b200.so"
HEADER_PATH = _PKG.parent / "include" / "splinegrids_b200.h"

_lib: Optional[C.CDLL] = None


class SplineGridsB200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile the CUDA library in-tree (nvcc, sm_100a).  Used by ``__graft_entry__.build()``."""
    import subprocess
    cmd = ["make", "-C", str(_PKG / "csrc"), "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SplineGridsB200Error("building libsplinegrids_b200.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SplineGridsB200Error(
                f"{LIB_PATH} not found: build it with `make -C {_PKG / 'csrc'}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "splinegrids.jl_b200 has no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.sg_version.restype = C.c_int
        _lib.sg_status_string.restype = C.c_char_p
        _lib.sg_status_string.argtypes = [C.c_int]
        _lib.sg_launch_count.restype = C.c_int64
        _lib.sg_launch_count_reset.restype = None
        _lib.sg_set_kernel_policy.restype = None
        _lib.sg_set_kernel_policy.argtypes = [C.c_int]
        _lib.sg_last_variant.restype = C.c_char_p
        _lib.sg_evaluate_adjoint_workspace_bytes.restype = C.c_size_t
        _lib.sg_adjoint_plan_destroy.argtypes = [C.c_void_p]
        _lib.sg_adjoint_plan_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.sg_profile_adjoint_main.restype = None
        _lib.sg_profile_adjoint_main.argtypes = [C.c_int]
        _lib.sg_profile_adjoint_main_ms.restype = C.c_float
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().sg_status_string(int(status)).decode()
        raise SplineGridsB200Error(f"{what} failed with status {status}: {msg}")


def suffix(dtype) -> str:
    import torch
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"unsupported float type {dtype}; the library provides Float32 and Float64")


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def ptr_array(ts: Sequence) -> C.Array:
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def i64_array(v: Sequence[int]) -> C.Array:
    return (C.c_int64 * len(v))(*[int(x) for x in v])


def int_array(v: Sequence[int]) -> C.Array:
    return (C.c_int * len(v))(*[int(x) for x in v])


def stream_ptr(device=None) -> C.c_void_p:
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count() -> int:
    return int(lib().sg_launch_count())


def launch_count_reset() -> None:
    lib().sg_launch_count_reset()


def set_kernel_policy(policy: int) -> None:
    """0 = automatic, 1 = generic kernels only, 2 = prefer the tiled fast paths."""
    lib().sg_set_kernel_policy(int(policy))


def last_variant() -> str:
    return lib().sg_last_variant().decode()


def declared_symbols() -> list:
    """Names of every function the public header declares (used by the ABI export test)."""
    import re
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text)))
