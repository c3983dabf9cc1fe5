"""Column-major device arrays.

The reference stores everything in Julia's column-major layout (first index fastest) and the C
ABI consumes exactly that memory.  On the Python side an array with *Julia shape* ``(n1, .., nk)`` is
a torch CUDA tensor whose shape is ``(n1, .., nk)`` and whose strides are ``(1, n1, n1*n2, ..)`` -- a
permuted view of a contiguous tensor of the reversed shape.  Indexing therefore reads like the
reference (``eval[j1, j2, o]``, 0-based) while ``data_ptr()`` is the Julia buffer byte for byte.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

FLOAT_TYPES = {"Float32": torch.float32, "Float64": torch.float64,
               np.float32: torch.float32, np.float64: torch.float64,
               np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
               torch.float32: torch.float32, torch.float64: torch.float64}

NP_OF = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}


def float_type(ft) -> torch.dtype:
    try:
        return FLOAT_TYPES[ft]
    except (KeyError, TypeError):
        raise TypeError(f"float_type must be Float32 or Float64, got {ft!r}")


def require_cuda(device=None) -> torch.device:
    """The only backend is CUDA; fail loudly otherwise (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise RuntimeError("splinegrids.jl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"splinegrids.jl_b200 arrays live on CUDA devices, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _rev(n: int):
    return tuple(range(n - 1, -1, -1))


def jl_empty(shape: Sequence[int], dtype, device) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    t = torch.empty(tuple(reversed(shape)), dtype=dtype, device=device)
    return t.permute(_rev(len(shape))) if len(shape) > 1 else t


def jl_zeros(shape: Sequence[int], dtype, device) -> torch.Tensor:
    t = jl_empty(shape, dtype, device)
    t.zero_()
    return t


def jl_ones(shape: Sequence[int], dtype, device) -> torch.Tensor:
    t = jl_empty(shape, dtype, device)
    t.fill_(1)
    return t


def is_colmajor(t: torch.Tensor) -> bool:
    """True if ``t`` is dense in column-major order (what the C ABI expects)."""
    if t.dim() <= 1:
        return t.is_contiguous()
    return t.permute(_rev(t.dim())).is_contiguous()


def as_colmajor(t: torch.Tensor) -> torch.Tensor:
    """Return ``t`` itself if already column-major, else a column-major copy."""
    if is_colmajor(t):
        return t
    out = jl_empty(t.shape, t.dtype, t.device)
    out.copy_(t)
    return out


def to_device(a, dtype=None, device=None) -> torch.Tensor:
    """numpy array (any order) or tensor -> column-major device tensor with the same Julia shape
    (the ``adapt(backend, x)`` of the reference, src/utils.jl:118-185)."""
    device = require_cuda(device)
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=dtype if dtype is not None else a.dtype)
        return as_colmajor(t)
    a = np.asarray(a)
    host = torch.from_numpy(np.ascontiguousarray(a.T))          # reversed shape, C order == F order of a
    if dtype is not None:
        host = host.to(dtype)
    dev = host.to(device)
    return dev.permute(_rev(a.ndim)) if a.ndim > 1 else dev


def to_numpy(t: torch.Tensor) -> np.ndarray:
    """Device tensor -> numpy array of the same Julia shape, Fortran order (``adapt(CPU(), x)``)."""
    if t.dim() <= 1:
        return t.detach().cpu().numpy()
    c = t.detach().permute(_rev(t.dim())).contiguous().cpu().numpy()
    return c.T


def reshape_colmajor(t: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """Julia ``reshape`` of a column-major array (no copy)."""
    assert is_colmajor(t), "reshape_colmajor needs a column-major array"
    shape = tuple(int(s) for s in shape)
    flat = t.permute(_rev(t.dim())).reshape(-1) if t.dim() > 1 else t
    assert flat.numel() == int(np.prod(shape)), "reshape: size mismatch"
    r = flat.view(tuple(reversed(shape)))
    return r.permute(_rev(len(shape))) if len(shape) > 1 else r
