"""Device side of the refinement set-up (scope row f3): thin ctypes wrappers over the K13-K18 / Flag / compaction entry
points of ``csrc/sg_refine_setup.cu``.  Everything stays in device memory; the only host traffic is the handful of
counts the reference also brings back (``sum(n_nonzero_C)``, the length of a ``findall``)."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import _lib
from .arrays import jl_empty
from .config import after_launch

C = _lib.C


def _st(dev):
    return _lib.stream_ptr(dev)


def boehm_matrix(knots_all_old: torch.Tensor, degree: int, knot_span_index: int, knot_new) -> Tuple[torch.Tensor, ...]:
    """K13 -- src/refinement.jl:3-36: ``(row_pointer, column_start, nzval)`` of the (n+1) x n insertion matrix."""
    dev, dt = knots_all_old.device, knots_all_old.dtype
    n = knots_all_old.numel() - degree - 1
    rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
    cs = torch.empty(n + 1, dtype=torch.int32, device=dev)
    nz = torch.empty(n + degree + 1, dtype=dt, device=dev)
    fn = getattr(_lib.lib(), "sg_boehm_matrix_" + _lib.suffix(dt))
    kn = C.c_float(float(knot_new)) if dt == torch.float32 else C.c_double(float(knot_new))
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.ptr(rp), _lib.ptr(cs), _lib.ptr(nz), _lib.ptr(knots_all_old), C.c_int64(n), C.c_int64(knot_span_index),
                      kn, C.c_int(degree), _st(dev)), "sg_boehm_matrix")
    after_launch(dev)
    return rp, cs, nz


def refmat_valid_rows(rp: torch.Tensor, cs: torch.Tensor, nnz: int, n_columns: int) -> torch.Tensor:
    """K14 -- src/refinement_matrix.jl:134-181: uint8 validity flag per row."""
    m = rp.numel()
    valid = torch.empty(m, dtype=torch.uint8, device=rp.device)
    with torch.cuda.device(rp.device):
        _lib.check(_lib.lib().sg_refmat_validate_i32(_lib.ptr(valid), _lib.ptr(rp), _lib.ptr(cs), C.c_int64(m), C.c_int64(nnz),
                                                     C.c_int64(n_columns), _st(rp.device)), "sg_refmat_validate")
    after_launch(rp.device)
    return valid


def refmat_matmul(A, B):
    """K16 + scan + K15 -- src/refinement_matrix.jl:273-327: ``(row_pointer, column_start, nzval)`` of ``A * B``."""
    dev, dt = A.nzval.device, A.nzval.dtype
    lib = _lib.lib()
    counts = torch.empty(A.m, dtype=torch.int32, device=dev)
    cs_c = torch.empty(A.m, dtype=torch.int32, device=dev)
    rp_c = torch.empty(A.m + 1, dtype=torch.int32, device=dev)
    total = C.c_int64(0)
    with torch.cuda.device(dev):
        _lib.check(lib.sg_refmat_mul_nonzeros_i32(_lib.ptr(counts), _lib.ptr(cs_c), _lib.ptr(A.row_pointer), _lib.ptr(A.column_start),
                                                  C.c_int64(A.m), C.c_int64(A.nzval.numel()), _lib.ptr(B.row_pointer),
                                                  _lib.ptr(B.column_start), C.c_int64(B.m), C.c_int64(B.nzval.numel()), C.c_int64(B.n),
                                                  _st(dev)), "sg_refmat_mul_nonzeros")
        _lib.check(lib.sg_row_pointer_from_counts(_lib.ptr(rp_c), C.byref(total), _lib.ptr(counts), C.c_int64(A.m), _st(dev)),
                   "sg_row_pointer_from_counts")
        nz_c = torch.empty(int(total.value), dtype=dt, device=dev)
        fn = getattr(lib, "sg_refmat_mul_values_" + _lib.suffix(dt))
        _lib.check(fn(_lib.ptr(nz_c), C.c_int64(nz_c.numel()), _lib.ptr(rp_c), _lib.ptr(cs_c),
                      _lib.ptr(A.row_pointer), _lib.ptr(A.column_start), _lib.ptr(A.nzval), C.c_int64(A.m), C.c_int64(A.nzval.numel()),
                      _lib.ptr(B.row_pointer), _lib.ptr(B.column_start), _lib.ptr(B.nzval), C.c_int64(B.m), C.c_int64(B.nzval.numel()),
                      _st(dev)), "sg_refmat_mul_values")
    after_launch(dev)
    return rp_c[:A.m], cs_c, nz_c


def refmat_collect(A) -> torch.Tensor:
    """K17 -- src/refinement_matrix.jl:329-363: dense column-major ``(m, n)`` device array."""
    out = jl_empty((A.m, A.n), A.nzval.dtype, A.nzval.device)
    fn = getattr(_lib.lib(), "sg_refmat_collect_" + _lib.suffix(A.nzval.dtype))
    with torch.cuda.device(A.nzval.device):
        _lib.check(fn(_lib.ptr(out), _lib.ptr(A.row_pointer), _lib.ptr(A.column_start), _lib.ptr(A.nzval), C.c_int64(A.m), C.c_int64(A.n),
                      C.c_int64(A.nzval.numel()), _st(A.nzval.device)), "sg_refmat_collect")
    after_launch(A.nzval.device)
    return out


def compact_flags(flags: torch.Tensor, invert: bool = False) -> torch.Tensor:
    """``findall(flags)`` (or ``findall(.!flags)``): ascending 0-based Int32 positions (device)."""
    n = flags.numel()
    dev = flags.device
    out = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    scratch = torch.empty(n + 1, dtype=torch.int32, device=dev)
    count = C.c_int64(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sg_compact_flags(_lib.ptr(out), C.byref(count), _lib.ptr(flags), C.c_int64(n), C.c_int(1 if invert else 0), _lib.ptr(scratch), _st(dev)),
                   "sg_compact_flags")
    return out[:int(count.value)]


def gather_rows(mat: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """``mat[rows, :]`` for a column-major (n, ncols) device matrix and 0-based Int32 row numbers."""
    n_in, ncols = int(mat.shape[0]), int(mat.shape[1])
    out = jl_empty((rows.numel(), ncols), mat.dtype, mat.device)
    suf = "i32" if mat.dtype == torch.int32 else _lib.suffix(mat.dtype)
    with torch.cuda.device(mat.device):
        _lib.check(getattr(_lib.lib(), "sg_gather_rows_" + suf)(_lib.ptr(out), _lib.ptr(mat), _lib.ptr(rows), C.c_int64(n_in),
                                                                C.c_int64(rows.numel()), C.c_int(ncols), _st(mat.device)), "sg_gather_rows")
    after_launch(mat.device)
    return out


def refinement_values_new(values_old: torch.Tensor, cp_refined: torch.Tensor, idx_new: torch.Tensor) -> torch.Tensor:
    """K18 -- src/control_points.jl:427-456."""
    n_new, nin = int(idx_new.shape[0]), int(idx_new.shape[1])
    nout = int(cp_refined.shape[-1])
    out = jl_empty((n_new, nout), cp_refined.dtype, cp_refined.device)
    fn = getattr(_lib.lib(), "sg_refinement_values_new_" + _lib.suffix(cp_refined.dtype))
    with torch.cuda.device(cp_refined.device):
        _lib.check(fn(_lib.ptr(out), _lib.ptr(values_old), C.c_int64(values_old.shape[0]), _lib.ptr(cp_refined), C.c_int(nin),
                      _lib.i64_array(cp_refined.shape[:-1]), C.c_int(nout), _lib.ptr(idx_new), C.c_int64(n_new), _st(cp_refined.device)),
                   "sg_refinement_values_new")
    after_launch(cp_refined.device)
    return out


def unique_new_rows(old_idx: torch.Tensor, new_idx: torch.Tensor, n_cp: Sequence[int]) -> torch.Tensor:
    """uint8 flags: which rows of ``new_idx`` survive ``unique(vcat(old, new); dims = 1)`` (first occurrences kept)."""
    dev = new_idx.device
    total = 1
    for c in n_cp:
        total *= int(c)
    keep = torch.empty(new_idx.shape[0], dtype=torch.uint8, device=dev)
    scratch = torch.empty(total, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sg_unique_new_rows_i32(_lib.ptr(keep), _lib.ptr(old_idx), C.c_int64(old_idx.shape[0]), _lib.ptr(new_idx),
                                                     C.c_int64(new_idx.shape[0]), C.c_int(len(n_cp)), _lib.i64_array(n_cp),
                                                     _lib.ptr(scratch), _st(dev)), "sg_unique_new_rows")
    after_launch(dev)
    return keep


def scatter_active_flag(flags: torch.Tensor, n_cp: Sequence[int], idx: torch.Tensor, value: bool) -> None:
    with torch.cuda.device(flags.device):
        _lib.check(_lib.lib().sg_scatter_active_flag(_lib.ptr(flags), C.c_int(len(n_cp)), _lib.i64_array(n_cp), _lib.ptr(idx),
                                                     C.c_int64(idx.shape[0]), C.c_int(1 if value else 0), _st(flags.device)),
                   "sg_scatter_active_flag")
    after_launch(flags.device)


def gather_active_flag(flags: torch.Tensor, n_cp: Sequence[int], idx: torch.Tensor) -> torch.Tensor:
    out = torch.empty(idx.shape[0], dtype=torch.uint8, device=flags.device)
    with torch.cuda.device(flags.device):
        _lib.check(_lib.lib().sg_gather_active_flag(_lib.ptr(out), _lib.ptr(flags), C.c_int(len(n_cp)), _lib.i64_array(n_cp), _lib.ptr(idx),
                                                    C.c_int64(idx.shape[0]), _st(flags.device)), "sg_gather_active_flag")
    after_launch(flags.device)
    return out


def mult_adjoint_flag(size_b: Sequence[int], As, Y: torch.Tensor, size_y: Sequence[int], dims_refinement: Sequence[int]) -> torch.Tensor:
    """K6 on Flag arrays -- src/adjoint.jl:117-121: OR over the structural windows."""
    total = 1
    for c in size_b:
        total *= int(c)
    B = torch.empty(total, dtype=torch.uint8, device=Y.device)
    with torch.cuda.device(Y.device):
        _lib.check(_lib.lib().sg_refmat_mul_adjoint_flag(_lib.ptr(B), _lib.ptr(Y), C.c_int(len(size_y)), _lib.i64_array(size_y),
                                                         _lib.i64_array(size_b), C.c_int(len(As)), _lib.int_array(dims_refinement),
                                                         _lib.ptr_array([A.row_pointer for A in As]),
                                                         _lib.ptr_array([A.column_start for A in As]),
                                                         _lib.i64_array([A.nzval.numel() for A in As]), _st(Y.device)),
                   "sg_refmat_mul_adjoint_flag")
    after_launch(Y.device)
    return B


def error_flags(cp_err: torch.Tensor, threshold_factor: float) -> torch.Tensor:
    """Reduce + threshold of error_informed_local_refinement! -- src/control_points.jl:556-563: uint8 flag per control point."""
    dev, dt = cp_err.device, cp_err.dtype
    nout = int(cp_err.shape[-1])
    total = cp_err.numel() // nout
    flags = torch.empty(total, dtype=torch.uint8, device=dev)
    grid_err = torch.empty(total, dtype=dt, device=dev)
    sums = torch.empty((total + 255) // 256, dtype=dt, device=dev)
    fn = getattr(_lib.lib(), "sg_error_flags_" + _lib.suffix(dt))
    tf = C.c_float(threshold_factor) if dt == torch.float32 else C.c_double(threshold_factor)
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.ptr(flags), _lib.ptr(grid_err), C.c_void_p(0), _lib.ptr(cp_err), C.c_int64(total), C.c_int(nout), tf,
                      _lib.ptr(sums), _st(dev)), "sg_error_flags")
    after_launch(dev)
    return flags


def indices_from_linear(lin: torch.Tensor, n_cp: Sequence[int]) -> torch.Tensor:
    out = jl_empty((lin.numel(), len(n_cp)), torch.int32, lin.device)
    with torch.cuda.device(lin.device):
        _lib.check(_lib.lib().sg_indices_from_linear_i32(_lib.ptr(out), _lib.ptr(lin), C.c_int64(lin.numel()), C.c_int(len(n_cp)),
                                                         _lib.i64_array(n_cp), _st(lin.device)), "sg_indices_from_linear")
    after_launch(lin.device)
    return out
