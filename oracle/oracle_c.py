"""ctypes front end of the C oracle (``oracle/liboracle.so``; build with ``make -C oracle``).

TEST INFRASTRUCTURE ONLY -- same rules as ``oracle_np`` (see its header).  Same array
conventions: numpy arrays in Julia shapes and Fortran order, 1-based index arrays.
The C code is the reference algorithm (one work item per sample, full window loop, atomics in
the adjoint) under OpenMP; it doubles as the timed CPU baseline of ``bench.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    srcs = [_HERE / "oracle.c", _HERE / "oracle_impl.h"]
    if force or not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        _LIB = C.CDLL(str(so))
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


def _suf(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(dtype)


def _p(a: np.ndarray):
    assert a.flags.f_contiguous or a.ndim <= 1 and a.flags.c_contiguous, "need Fortran-contiguous array"
    return C.c_void_p(a.ctypes.data)


def _i64(v: Sequence[int]):
    return (C.c_int64 * len(v))(*[int(x) for x in v])


def _i32(v: Sequence[int]):
    return (C.c_int * len(v))(*[int(x) for x in v])


def _ptrs(arrs: Sequence[np.ndarray]):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def span_indices(sample_points: np.ndarray, knots_all: np.ndarray, degree: int) -> np.ndarray:
    out = np.empty(len(sample_points), dtype=np.int32)
    getattr(lib(), "orc_span_indices_" + _suf(knots_all.dtype))(
        _p(sample_points), C.c_int64(len(sample_points)), _p(knots_all), C.c_int64(len(knots_all)),
        C.c_int(degree), _p(out))
    return out


def basis_tables(knots_all, sample_points, sample_indices, degree: int, mdo: int = 0) -> np.ndarray:
    n = len(sample_points)
    out = np.zeros((n, degree + 1, mdo + 1), dtype=knots_all.dtype, order="F")
    getattr(lib(), "orc_basis_tables_" + _suf(knots_all.dtype))(
        _p(out), _p(knots_all), _p(sample_points), _p(sample_indices), C.c_int64(n),
        C.c_int(degree), C.c_int(mdo))
    return out


def evaluate(tables, sample_indices, degrees, derivative_order, control_points: np.ndarray,
             weights: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
    nin = len(tables)
    n = [len(ix) for ix in sample_indices]
    nout = control_points.shape[-1]
    if out is None:
        out = np.empty(tuple(n) + (nout,), dtype=control_points.dtype, order="F")
    getattr(lib(), "orc_evaluate_" + _suf(control_points.dtype))(
        _p(out), C.c_int(nin), _i64(n), _i64(control_points.shape[:-1]), C.c_int(nout),
        _ptrs(tables), _ptrs(sample_indices), _i32(degrees), _i32(derivative_order),
        _p(control_points), _p(weights) if weights is not None else None)
    return out


def evaluate_adjoint(tables, sample_indices, degrees, derivative_order, eval_in: np.ndarray,
                     cp_shape: Sequence[int], weights: Optional[np.ndarray] = None,
                     out: Optional[np.ndarray] = None) -> np.ndarray:
    nin = len(tables)
    n = [len(ix) for ix in sample_indices]
    nout = eval_in.shape[-1]
    if out is None:
        out = np.empty(tuple(cp_shape), dtype=eval_in.dtype, order="F")
    getattr(lib(), "orc_evaluate_adjoint_" + _suf(eval_in.dtype))(
        _p(out), C.c_int(nin), _i64(n), _i64(cp_shape[:-1]), C.c_int(nout),
        _ptrs(tables), _ptrs(sample_indices), _i32(degrees), _i32(derivative_order),
        _p(eval_in), _p(weights) if weights is not None else None)
    return out


def _refmat_args(ndims, As, dims_refinement):
    of_dim = [-1] * ndims
    for a, d in enumerate(dims_refinement):
        of_dim[d - 1] = a
    return (_i32(of_dim), _ptrs([A.row_pointer for A in As]), _ptrs([A.column_start for A in As]),
            _ptrs([A.nzval for A in As]), _i64([len(A.nzval) for A in As]))


def mult(Y: np.ndarray, As, B: np.ndarray, dims_refinement) -> None:
    """K5 through the C oracle.  ``As`` are oracle_np.RefMat with int32 index arrays."""
    getattr(lib(), "orc_refmat_mul_" + _suf(Y.dtype))(
        _p(Y), _p(B), C.c_int(Y.ndim), _i64(Y.shape), _i64(B.shape), *_refmat_args(Y.ndim, As, dims_refinement))


def mult_adjoint(B: np.ndarray, As, Y: np.ndarray, dims_refinement) -> None:
    getattr(lib(), "orc_refmat_mul_adjoint_" + _suf(Y.dtype))(
        _p(B), _p(Y), C.c_int(Y.ndim), _i64(Y.shape), _i64(B.shape), *_refmat_args(Y.ndim, As, dims_refinement))


def scatter_active(cp: np.ndarray, idx: np.ndarray, vals: np.ndarray) -> None:
    idx = np.asfortranarray(idx, dtype=np.int32)
    vals = np.asfortranarray(vals)
    getattr(lib(), "orc_scatter_active_" + _suf(cp.dtype))(
        _p(cp), C.c_int(cp.ndim - 1), _i64(cp.shape[:-1]), C.c_int(cp.shape[-1]), _p(idx), _p(vals),
        C.c_int64(idx.shape[0]))


def gather_zero_active(vals: np.ndarray, cp: np.ndarray, idx: np.ndarray) -> None:
    idx = np.asfortranarray(idx, dtype=np.int32)
    assert vals.flags.f_contiguous
    getattr(lib(), "orc_gather_zero_active_" + _suf(cp.dtype))(
        _p(vals), _p(cp), C.c_int(cp.ndim - 1), _i64(cp.shape[:-1]), C.c_int(cp.shape[-1]), _p(idx),
        C.c_int64(idx.shape[0]))
