"""``RefinementMatrix`` and its application ``mult!`` / ``mult_adjoint!`` -- mirrors
src/refinement_matrix.jl and src/adjoint.jl:85-152 of the reference.

The application (K5/K6) AND the set-up algebra -- validation (K14), the sparse product (K16 + scan + K15) and
``collect`` (K17) -- run on the device through the C ABI (``device_setup.py``); the arrays live in device memory and
the ``*_host`` numpy mirrors are fetched lazily, only when host code (printing, indexing, tests) asks for them.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .arrays import NP_OF, float_type, is_colmajor, require_cuda
from .config import after_launch
from .validation import SplineGridsError, validate_mult_input

C = _lib.C


class RefinementMatrix:
    """Banded-row sparse matrix: row ``i`` holds consecutive non-zeros
    ``nzval[row_pointer[i] : row_pointer[i+1]-1]`` starting at column ``column_start[i]`` (all 1-based, as
    stored by the reference, src/refinement_matrix.jl:19-57).

    ``row_pointer`` / ``column_start`` / ``nzval`` are device tensors (Int32 / Tv); ``*_host`` are the
    numpy mirrors used by the set-up algebra."""

    def __init__(self, m: int, n: int, row_pointer, column_start, nzval, device=None, validate: bool = True):
        self.m, self.n = int(m), int(n)
        self._rp_h = self._cs_h = self._nz_h = None
        on_dev = all(isinstance(t, torch.Tensor) and t.is_cuda for t in (row_pointer, column_start, nzval))
        if on_dev:                                    # built on the device (K13 / K15): no host round trip
            self.row_pointer = row_pointer.to(torch.int32).contiguous()
            self.column_start = column_start.to(torch.int32).contiguous()
            self.nzval = nzval.contiguous()
        else:
            rp = _host(row_pointer).astype(np.int32)
            cs = _host(column_start).astype(np.int32)
            nz = _host(nzval)
            if nz.dtype not in (np.float32, np.float64):
                nz = nz.astype(np.float64)
            assert np.array_equal(rp, np.sort(rp))
            self._rp_h, self._cs_h, self._nz_h = rp, cs, nz
            dev = require_cuda(device)
            self.row_pointer = torch.from_numpy(rp).to(dev)
            self.column_start = torch.from_numpy(cs).to(dev)
            self.nzval = torch.from_numpy(nz).to(dev)
        assert self.row_pointer.numel() == self.column_start.numel() == self.m
        if validate:
            bad = self.invalid_rows()
            if bad:
                raise SplineGridsError(f"Invalid rows: [{', '.join(str(b) for b in bad)}].")

    # -- lazily fetched host mirrors --------------------------------------------------------------
    @property
    def row_pointer_host(self) -> np.ndarray:
        if self._rp_h is None:
            self._rp_h = self.row_pointer.cpu().numpy()
        return self._rp_h

    @property
    def column_start_host(self) -> np.ndarray:
        if self._cs_h is None:
            self._cs_h = self.column_start.cpu().numpy()
        return self._cs_h

    @property
    def nzval_host(self) -> np.ndarray:
        if self._nz_h is None:
            self._nz_h = self.nzval.cpu().numpy()
        return self._nz_h

    # -- structure helpers (src/refinement_matrix.jl:103-125) -----------------------------------
    def _row_lengths(self) -> np.ndarray:
        nxt = np.append(self.row_pointer_host[1:], len(self.nzval_host) + 1).astype(np.int64)
        return nxt - self.row_pointer_host.astype(np.int64)

    def column_ranges(self) -> Tuple[np.ndarray, np.ndarray]:
        """(column_start, column_end) per row, 1-based inclusive (host arrays)."""
        cs = self.column_start_host.astype(np.int64)
        return cs, cs + self._row_lengths() - 1

    def invalid_rows(self):
        """``validate_refinement_matrix_kernel`` (K14, on the device) -- src/refinement_matrix.jl:134-181: 1-based numbers
        of the invalid rows (``findall(.!valid_row)``)."""
        from . import device_setup as D
        valid = D.refmat_valid_rows(self.row_pointer, self.column_start, self.nzval.numel(), self.n)
        bad = D.compact_flags(valid, invert=True)
        return [int(i) + 1 for i in bad.cpu().tolist()] if bad.numel() else []

    @property
    def shape(self):
        return (self.m, self.n)

    def __len__(self):
        return self.m * self.n

    @property
    def dtype(self) -> torch.dtype:
        return self.nzval.dtype

    @property
    def device(self):
        return self.nzval.device

    def __eq__(self, other):   # src/refinement_matrix.jl:62-70
        return (isinstance(other, RefinementMatrix) and self.shape == other.shape
                and np.array_equal(self.row_pointer_host, other.row_pointer_host)
                and np.array_equal(self.column_start_host, other.column_start_host)
                and np.array_equal(self.nzval_host, other.nzval_host))

    def __getitem__(self, ij):  # src/refinement_matrix.jl:72-86 (1-based i, j)
        i, j = ij
        if not (1 <= i <= self.m and 1 <= j <= self.n):
            raise SplineGridsError(f"Index ({i}, {j}) out of bounds for refinement matrix of size ({self.m}, {self.n}).")
        cs, ce = self.column_ranges()
        if cs[i - 1] <= j <= ce[i - 1]:
            return self.nzval_host[self.row_pointer_host[i - 1] - 1 + j - cs[i - 1]]
        return self.nzval_host.dtype.type(0)

    def collect(self) -> np.ndarray:
        """``collect(A)`` (K17, on the device) -- dense matrix, src/refinement_matrix.jl:329-363 (returned as numpy)."""
        from . import device_setup as D
        return D.refmat_collect(self).cpu().numpy()

    def collect_device(self) -> torch.Tensor:
        """``collect(A)`` as a column-major device array."""
        from . import device_setup as D
        return D.refmat_collect(self)

    def __matmul__(self, other: "RefinementMatrix") -> "RefinementMatrix":
        """``A * B`` -- src/refinement_matrix.jl:273-327 on the device: K16 (non-zeros and first column per row), exclusive
        scan for the row pointers, K15 (values, terms accumulated in ascending k like the reference kernel)."""
        from . import device_setup as D
        A, B = self, other
        if A.n != B.m:
            raise ValueError("DimensionMismatch: Inner dimensions must match")
        rp, cs, nz = D.refmat_matmul(A, B)
        return RefinementMatrix(A.m, B.n, rp, cs, nz, device=A.device)

    def __repr__(self):
        return f"RefinementMatrix({self.m}x{self.n}, nnz={self.nzval.numel()}, {self.dtype})"


def _host(a) -> np.ndarray:
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def rmeye(n: int, device=None, float_type_="Float32") -> RefinementMatrix:
    """``rmeye`` -- identity refinement matrix, src/refinement_matrix.jl:463-474."""
    r = np.arange(1, n + 1, dtype=np.int32)
    return RefinementMatrix(n, n, r, r.copy(), np.ones(n, dtype=NP_OF[float_type(float_type_)]), device=device)


def refinement_matrix_from_dense(A: np.ndarray, device=None) -> RefinementMatrix:
    """``RefinementMatrix(A::Matrix)`` -- src/refinement_matrix.jl:476-504 (raises "Invalid rows: [...]")."""
    A = np.asarray(A)
    rp, cs, nz = [], [], []
    pointer = 1
    for row in A:
        nzi = np.flatnonzero(row)
        rp.append(pointer)
        if len(nzi) == 0:
            cs.append(0)
        else:
            cs.append(int(nzi[0]) + 1)
            nz.append(row[nzi[0]:nzi[-1] + 1])
            pointer += int(nzi[-1] - nzi[0]) + 1
    nzv = np.concatenate(nz) if nz else np.zeros(0, dtype=A.dtype)
    return RefinementMatrix(A.shape[0], A.shape[1], rp, cs, nzv, device=device)


def _refmat_call(name: str, out: torch.Tensor, inp: torch.Tensor, Y: torch.Tensor, B: torch.Tensor,
                 As: Sequence[RefinementMatrix], dims_refinement: Sequence[int]) -> None:
    assert is_colmajor(Y) and is_colmajor(B), "mult!/mult_adjoint! need dense column-major arrays"
    assert Y.dtype == B.dtype and all(A.dtype == Y.dtype for A in As)
    n_ref = len(As)
    with torch.cuda.device(Y.device):
        fn = getattr(_lib.lib(), name + _lib.suffix(Y.dtype))
        _lib.check(fn(_lib.ptr(out), _lib.ptr(inp), C.c_int(Y.dim()), _lib.i64_array(Y.shape), _lib.i64_array(B.shape),
                      C.c_int(n_ref), _lib.int_array(dims_refinement),
                      _lib.ptr_array([A.row_pointer for A in As]), _lib.ptr_array([A.column_start for A in As]),
                      _lib.ptr_array([A.nzval for A in As]), _lib.i64_array([A.nzval.numel() for A in As]),
                      _lib.stream_ptr(Y.device)), name.rstrip("_"))
    after_launch(Y.device)


def mult_(Y: torch.Tensor, As: Sequence[RefinementMatrix], B: torch.Tensor, dims_refinement: Sequence[int]) -> None:
    """``mult!(Y, As, B, dims_refinement)`` (K5) -- src/refinement_matrix.jl:421-445: left-multiply ``B`` by
    every refinement matrix along its (1-based) dimension, into ``Y``."""
    validate_mult_input(tuple(Y.shape), As, tuple(B.shape), dims_refinement)
    _refmat_call("sg_refmat_mul_", Y, B, Y, B, As, dims_refinement)


def mult_adjoint_(B: torch.Tensor, As: Sequence[RefinementMatrix], Y: torch.Tensor,
                  dims_refinement: Sequence[int]) -> None:
    """``mult_adjoint!(B, As, Y, dims_refinement)`` (K6) -- src/adjoint.jl:127-152 (B is overwritten)."""
    validate_mult_input(tuple(Y.shape), As, tuple(B.shape), dims_refinement)
    _refmat_call("sg_refmat_mul_adjoint_", B, Y, Y, B, As, dims_refinement)
