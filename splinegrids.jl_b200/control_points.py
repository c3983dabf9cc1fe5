"""Control points -- mirrors src/control_points.jl of the reference.

``DefaultControlPoints`` wraps a dense ``(c_1..c_D, Nout)`` array; ``LocallyRefinedControlPoints`` is the
THB hierarchy: per level a dense array plus a ``LocalRefinement`` (refinement matrices + the active
control points that overwrite the refined values).  Applying the hierarchy (``evaluate!`` /
``evaluate_adjoint!`` on the control points, K5-K8) runs on the device; editing it (activation,
deactivation) is set-up work done on the host on the small index arrays.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .arrays import (as_colmajor, is_colmajor, jl_empty, jl_zeros, reshape_colmajor, to_device, to_numpy)
from .config import after_launch
from .refinement_matrix import RefinementMatrix, mult_, mult_adjoint_

C = _lib.C


class AbstractControlPoints:
    def obtain(self) -> torch.Tensor:
        raise NotImplementedError

    # Base.ndims/size/length/eltype/vec overloads, src/control_points.jl:269-275
    @property
    def shape(self):
        return tuple(self.obtain().shape)

    def size(self):
        return self.shape

    @property
    def dtype(self):
        return self.obtain().dtype

    @property
    def device(self):
        return self.obtain().device

    @property
    def Nin(self) -> int:
        return self.obtain().dim() - 1

    @property
    def Nout(self) -> int:
        return self.obtain().shape[-1]


class DefaultControlPoints(AbstractControlPoints):
    """``DefaultControlPoints(control_points)`` -- src/control_points.jl:6-24."""

    def __init__(self, control_points: torch.Tensor):
        self.control_points = as_colmajor(control_points)

    def obtain(self) -> torch.Tensor:
        return self.control_points

    def n_control_points(self) -> int:
        return int(np.prod(self.shape[:-1]))

    def __getitem__(self, inds):
        return self.control_points[inds]

    def __setitem__(self, inds, val):
        self.control_points[inds] = val

    def fill_(self, v):
        self.control_points.fill_(v)
        return self

    def __repr__(self):
        return (f"DefaultControlPoints for grid of size {self.shape[:-1]} in R^{self.Nout} ({self.dtype}).")


class LocalRefinement:
    """``LocalRefinement`` -- src/control_points.jl:70-110.  ``dims_refinement`` is 1-based;
    ``refinement_indices`` is a device Int32 ``(n_active, Nin)`` column-major matrix of 1-based indices;
    ``refinement_values`` is ``(n_active, Nout)`` column-major."""

    def __init__(self, dims_refinement: Sequence[int], refinement_matrices: Sequence[RefinementMatrix],
                 refinement_indices: torch.Tensor, refinement_values: torch.Tensor):
        self.dims_refinement = [int(d) for d in dims_refinement]
        self.refinement_matrices = list(refinement_matrices)
        self.refinement_indices = as_colmajor(refinement_indices)
        self.refinement_values = as_colmajor(refinement_values)

    @property
    def n_active(self) -> int:
        return int(self.refinement_indices.shape[0])


class LocallyRefinedControlPoints(AbstractControlPoints):
    """``LocallyRefinedControlPoints`` -- src/control_points.jl:173-199."""

    def __init__(self, control_points_refined: List[torch.Tensor], local_refinements: List[LocalRefinement]):
        self.control_points_refined = [as_colmajor(c) for c in control_points_refined]
        self.local_refinements = list(local_refinements)

    def obtain(self) -> torch.Tensor:                      # src/control_points.jl:354-356
        return self.control_points_refined[-1]

    def n_control_points(self) -> int:                     # src/control_points.jl:281-288
        return sum(lr.n_active for lr in self.local_refinements)

    def get_values(self) -> torch.Tensor:
        """The flat parameter matrix: vertical concatenation of the per-level ``refinement_values``
        (``get_control_point_view``, src/control_points.jl:201-208)."""
        return torch.cat([lr.refinement_values for lr in self.local_refinements], dim=0)

    def set_values(self, M: torch.Tensor) -> None:
        assert M.shape[0] == self.n_control_points()
        off = 0
        for lr in self.local_refinements:
            k = lr.n_active
            lr.refinement_values.copy_(M[off:off + k])
            off += k

    def __repr__(self):
        lines = [f"LocallyRefinedControlPoints for final grid of size {self.shape[:-1]} in R^{self.Nout} "
                 f"({self.dtype}). Local refinements:"]
        lines.append("  input dim. | # c.p. before | # c.p. after | # activated c.p.")
        lines.append(f"  - | - | - | {self.local_refinements[0].n_active}")
        for lr in self.local_refinements[1:]:
            for i, (d, A) in enumerate(zip(lr.dims_refinement, lr.refinement_matrices)):
                act = lr.n_active if i == len(lr.dims_refinement) - 1 else 0
                lines.append(f"  {d} | {A.n} | {A.m} | {act}")
        return "\n".join(lines)


def obtain(control_points) -> torch.Tensor:
    """``obtain`` -- src/control_points.jl:351-356: the dense array of the finest level (raw arrays pass)."""
    return control_points.obtain() if isinstance(control_points, AbstractControlPoints) else control_points


def get_n_control_points(x) -> int:
    cp = getattr(x, "control_points", x)
    if isinstance(cp, (DefaultControlPoints, LocallyRefinedControlPoints)):
        return cp.n_control_points()
    return int(np.prod(cp.shape[:-1]))


def copyto_(dst, src):
    """``copyto!`` overloads of src/control_points.jl:47-58, 224-233 (control points <-> matrices/arrays)."""
    if isinstance(dst, LocallyRefinedControlPoints):
        dst.set_values(_as_tensor(src, dst))
    elif isinstance(dst, DefaultControlPoints):
        s = _as_tensor(src, dst)
        t = dst.obtain()
        if tuple(s.shape) != tuple(t.shape):
            s = reshape_colmajor(as_colmajor(s), t.shape)
        t.copy_(s)
    elif isinstance(src, LocallyRefinedControlPoints):
        dst.copy_(src.get_values())
    elif isinstance(src, DefaultControlPoints):
        dst.copy_(reshape_colmajor(src.obtain(), dst.shape))
    else:
        dst.copy_(src)
    return dst


def _as_tensor(src, like) -> torch.Tensor:
    if isinstance(src, torch.Tensor):
        return src.to(device=like.device, dtype=like.dtype)
    return to_device(np.asarray(src), dtype=like.dtype, device=like.device)


# ---------------------------------------------------------------------------------------------
# K7 / K8 and the hierarchy loops
# ---------------------------------------------------------------------------------------------


def _scatter_active(cp: torch.Tensor, lr: LocalRefinement) -> None:
    with torch.cuda.device(cp.device):
        fn = getattr(_lib.lib(), "sg_scatter_active_" + _lib.suffix(cp.dtype))
        _lib.check(fn(_lib.ptr(cp), C.c_int(cp.dim() - 1), _lib.i64_array(cp.shape[:-1]), C.c_int(cp.shape[-1]),
                      _lib.ptr(lr.refinement_indices), _lib.ptr(lr.refinement_values), C.c_int64(lr.n_active),
                      _lib.stream_ptr(cp.device)), "sg_scatter_active")
    after_launch(cp.device)


def _gather_zero_active(cp: torch.Tensor, lr: LocalRefinement) -> None:
    with torch.cuda.device(cp.device):
        fn = getattr(_lib.lib(), "sg_gather_zero_active_" + _lib.suffix(cp.dtype))
        _lib.check(fn(_lib.ptr(lr.refinement_values), _lib.ptr(cp), C.c_int(cp.dim() - 1),
                      _lib.i64_array(cp.shape[:-1]), C.c_int(cp.shape[-1]), _lib.ptr(lr.refinement_indices),
                      C.c_int64(lr.n_active), _lib.stream_ptr(cp.device)), "sg_gather_zero_active")
    after_launch(cp.device)


def evaluate_control_points_(control_points) -> None:
    """``evaluate!(::AbstractControlPoints)`` -- src/control_points.jl:313-349: per level, apply the
    refinement matrices (K5) then overwrite the active control points (K7).  No-op for dense points."""
    if not isinstance(control_points, LocallyRefinedControlPoints):
        return None
    for i, lr in enumerate(control_points.local_refinements):
        cp_new = control_points.control_points_refined[i]
        if i > 0:
            mult_(cp_new, lr.refinement_matrices, control_points.control_points_refined[i - 1], lr.dims_refinement)
        if lr.n_active > 0:
            _scatter_active(cp_new, lr)
    return None


def evaluate_adjoint_control_points_(control_points) -> None:
    """``evaluate_adjoint!(::AbstractControlPoints)`` -- src/adjoint.jl:172-205: reverse level loop,
    gather-and-zero the active entries (K8) then ``mult_adjoint!`` to the level below (K6)."""
    if not isinstance(control_points, LocallyRefinedControlPoints):
        return None
    for i in reversed(range(len(control_points.local_refinements))):
        lr = control_points.local_refinements[i]
        cp_new = control_points.control_points_refined[i]
        if lr.n_active > 0:
            _gather_zero_active(cp_new, lr)
        if i > 0:
            mult_adjoint_(control_points.control_points_refined[i - 1], lr.refinement_matrices, cp_new,
                          lr.dims_refinement)
    return None


# ---------------------------------------------------------------------------------------------
# THB bookkeeping (set-up; host side on the index arrays) -- src/control_points.jl:427-680
# ---------------------------------------------------------------------------------------------


def base_local_refinement(cp_base: torch.Tensor) -> LocalRefinement:
    """Base level: every control point active, values = copy (src/control_points.jl:126-154; K12)."""
    shape = tuple(cp_base.shape[:-1])
    n_cp = int(np.prod(shape))
    idx = np.stack(np.unravel_index(np.arange(n_cp), shape, order="F"), axis=1).astype(np.int32) + 1
    vals = reshape_colmajor(cp_base, (n_cp, cp_base.shape[-1])).clone()
    return LocalRefinement([], [], to_device(idx, device=cp_base.device), vals)


def activate_local_refinement_(control_points: LocallyRefinedControlPoints, refinement_indices,
                               refinement_index: Optional[int] = None) -> None:
    """``activate_local_refinement!`` -- src/control_points.jl:470-516.  New active control points take the
    current refined value (geometry unchanged); duplicates are dropped keeping first occurrences
    (``unique(vcat(old, new); dims = 1)``: on the CPU in the reference, :482-494; here a first-row table on the device,
    ``sg_unique_new_rows``), then K18 fills the values.  Nothing but the number of kept rows comes back to the host."""
    from . import device_setup as D
    level = len(control_points.local_refinements) if refinement_index is None else refinement_index
    lr = control_points.local_refinements[level - 1]
    cp = control_points.control_points_refined[level - 1]
    nin = cp.dim() - 1
    if isinstance(refinement_indices, torch.Tensor):
        new = refinement_indices.to(device=cp.device, dtype=torch.int32)
    else:
        new = to_device(np.asarray(refinement_indices, dtype=np.int32).reshape(-1, nin), device=cp.device)
    assert new.dim() == 2 and new.shape[1] == nin, \
        "Number of indices per control point must match the number of input dimensions."
    new = as_colmajor(new)
    old = as_colmajor(lr.refinement_indices.reshape(-1, nin)) if lr.refinement_indices.numel() else \
        jl_empty((0, nin), torch.int32, cp.device)
    n_old = int(old.shape[0])
    if new.shape[0] > 0:
        keep = D.unique_new_rows(old, new, cp.shape[:-1])
        rows = D.compact_flags(keep)
        kept = D.gather_rows(new, rows)
    else:
        kept = new
    idx_new = jl_empty((n_old + kept.shape[0], nin), torch.int32, cp.device)
    idx_new[:n_old] = old
    idx_new[n_old:] = kept
    vals_old = lr.refinement_values if n_old else jl_empty((0, cp.shape[-1]), cp.dtype, cp.device)
    vals = D.refinement_values_new(as_colmajor(vals_old), cp, idx_new)       # K18 refinement_values_new_kernel (:427-456)
    control_points.local_refinements[level - 1] = LocalRefinement(lr.dims_refinement, lr.refinement_matrices,
                                                                  idx_new, vals)


def activate_local_control_point_range_(spline_grid_or_cp, *ranges) -> None:
    """``activate_local_control_point_range!`` -- src/control_points.jl:531-539.  Each range is a Python
    ``range`` of 1-BASED indices (``range(1, 5)`` is Julia's ``1:4``) or an inclusive ``(lo, hi)`` tuple."""
    cp = getattr(spline_grid_or_cp, "control_points", spline_grid_or_cp)
    rs = [r if isinstance(r, range) else range(r[0], r[1] + 1) for r in ranges]
    assert len(rs) == cp.Nin
    # Row order of the reference: ``reduce(vcat, A')`` of the Iterators.product array (:535-537) -- for Nin = 2 the
    # adjoint makes dimension 2 the fastest; Nin = 1 is unchanged; Nin >= 3 errors in the reference (an N-d array has no
    # adjoint), here it keeps the product order (dimension 1 fastest).
    if len(rs) == 2:
        rows = list(itertools.product(*rs))
    else:
        rows = [tuple(reversed(t)) for t in itertools.product(*reversed(rs))]
    activate_local_refinement_(cp, np.array(rows, dtype=np.int32).reshape(len(rows), len(rs)))


def deactivate_overwritten_control_points_(control_points: LocallyRefinedControlPoints,
                                           local_refinement_level: Optional[int] = None) -> None:
    """``deactivate_overwritten_control_points!`` -- src/control_points.jl:584-680, on the device like the reference:
    a Boolean ``Flag`` array of the next level starts all true, the next level's active entries are set false (K7 on
    Flags), ``mult_adjoint!`` on Flags ORs each structural window back to this level (K6, src/adjoint.jl:117-121), K8
    reads the flag of every active control point of this level, ``findall`` + row selection keep the visible ones."""
    from . import device_setup as D
    if local_refinement_level is None:
        for level in range(len(control_points.local_refinements) - 1, 0, -1):
            deactivate_overwritten_control_points_(control_points, level)
        return
    lrs = control_points.local_refinements
    assert 1 <= local_refinement_level <= len(lrs) - 1
    lr, lr_next = lrs[local_refinement_level - 1], lrs[local_refinement_level]
    nin = control_points.Nin
    dev = control_points.device
    shape = tuple(control_points.control_points_refined[local_refinement_level - 1].shape[:-1])
    shape_next = tuple(control_points.control_points_refined[local_refinement_level].shape[:-1])
    flags_next = torch.ones(int(np.prod(shape_next)), dtype=torch.uint8, device=dev)      # ones(Flag, ...) :636
    if lr_next.n_active:
        D.scatter_active_flag(flags_next, shape_next, as_colmajor(lr_next.refinement_indices.reshape(-1, nin)), False)
    flags = D.mult_adjoint_flag(shape, lr_next.refinement_matrices, flags_next, shape_next, lr_next.dims_refinement)
    if lr.n_active:
        idx = as_colmajor(lr.refinement_indices.reshape(-1, nin))
        visible = D.gather_active_flag(flags, shape, idx)
        rows = D.compact_flags(visible)                                # findall(f -> f.flag, refinement_values_) :667
        idx_new, vals_new = D.gather_rows(idx, rows), D.gather_rows(as_colmajor(lr.refinement_values), rows)
    else:
        idx_new, vals_new = lr.refinement_indices, lr.refinement_values
    lrs[local_refinement_level - 1] = LocalRefinement(lr.dims_refinement, lr.refinement_matrices, idx_new, vals_new)
