"""Knot insertion / refinement and the THB set-up helpers -- mirrors src/refinement.jl and
src/control_points.jl:364-425, 541-575 of the reference.

These are set-up-time operations ("next"): the O(n_knots) matrix construction runs on the
host in numpy; everything that touches control-point or sample arrays (``mult!``, basis tables, the
adjoint inside ``error_informed_local_refinement!``) runs on the device through the C ABI.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .arrays import jl_zeros, to_device, to_numpy
from .control_points import (DefaultControlPoints, LocalRefinement, LocallyRefinedControlPoints,
                             activate_local_refinement_, base_local_refinement, obtain)
from .knot_vector import KnotVector
from .refinement_matrix import RefinementMatrix, mult_, rmeye
from .spline_dimension import SplineDimension, build_, evaluate_dimension_, set_sample_indices_
from .spline_grid import SplineGrid, evaluate_adjoint_


def boehm_refinement_matrix(knots_all_old, degree: int, knot_span_index: int, knot_new,
                            device=None) -> RefinementMatrix:
    """``RefinementMatrix(spline_dimension, knot_span_index, knot_new)`` (K13, on the device) -- src/refinement.jl:3-88:
    Boehm's single-knot insertion as an (n+1) x n banded matrix.  ``knots_all_old``: device tensor or host array."""
    from . import device_setup as D
    if not isinstance(knots_all_old, torch.Tensor):
        knots_all_old = to_device(np.asarray(knots_all_old), device=device)
    n = knots_all_old.numel() - int(degree) - 1
    rp, cs, nz = D.boehm_matrix(knots_all_old, int(degree), int(knot_span_index), knot_new)
    return RefinementMatrix(n + 1, n, rp, cs, nz, device=knots_all_old.device)


def insert_knot(obj, *args, **kwargs):
    """``insert_knot`` -- the three methods of src/refinement.jl:107-235 (dispatch on the first argument)."""
    if isinstance(obj, KnotVector):
        return _insert_knot_kv(obj, *args, **kwargs)
    if isinstance(obj, SplineDimension):
        return _insert_knot_sd(obj, *args, **kwargs)
    return _insert_knot_grid(obj, *args, **kwargs)


def _insert_knot_kv(knot_vector: KnotVector, knot_new) -> Tuple[KnotVector, int]:
    """src/refinement.jl:107-133: returns ``(knot_vector_new, knot_span_index)``."""
    kv, mu = knot_vector.knot_values_host, knot_vector.multiplicities_host
    knot_new = kv.dtype.type(knot_new)
    assert not np.any(kv == knot_new)
    pos = int(np.searchsorted(kv, knot_new, side="left"))
    kv_new = np.insert(kv, pos, knot_new)
    mu_new = np.insert(mu, pos, 1)
    return KnotVector(kv_new, mu_new, device=knot_vector.device), int(mu[:pos].sum())


def _insert_knot_sd(sd: SplineDimension, knot_new, recompute_sample_indices: bool = True,
                    evaluate: bool = True) -> Tuple[SplineDimension, RefinementMatrix]:
    """src/refinement.jl:161-186."""
    kv_new, k = _insert_knot_kv(sd.knot_vector, knot_new)
    R = boehm_refinement_matrix(sd.knot_vector.knots_all, sd.degree, k, knot_new, device=sd.device)
    sd_new = sd.with_knot_vector(kv_new)
    if recompute_sample_indices:
        set_sample_indices_(sd_new)
        if evaluate:
            evaluate_dimension_(sd_new)
    return sd_new, R


def _insert_knot_grid(grid: SplineGrid, dim_refinement: int, knot_new, evaluate_spline_dimension: bool = True):
    """src/refinement.jl:210-235 (``dim_refinement`` is 1-based)."""
    assert not grid.is_nurbs(), "Knot insertion not supported for NURBS"
    sd_new, R = _insert_knot_sd(grid.spline_dimensions[dim_refinement - 1], knot_new,
                                evaluate=evaluate_spline_dimension)
    return _refine_grid_with(grid, sd_new, dim_refinement, R)


def refine(obj, *args, **kwargs):
    """``refine`` -- src/refinement.jl:256-386 (dispatch on the first argument)."""
    if isinstance(obj, SplineDimension):
        return _refine_sd(obj, *args, **kwargs)
    return _refine_grid(obj, *args, **kwargs)


def _refine_sd(sd: SplineDimension, knots_new: Optional[Sequence[float]] = None):
    """src/refinement.jl:256-284: insert several knots (default: midpoints of all knot-value intervals),
    accumulating the product of the single-knot matrices.  Sample indices are recomputed; the tables are
    NOT (the caller does ``evaluate!``), as in the reference."""
    kv = sd.knot_vector.knot_values_host
    if knots_new is None:
        knots_new = kv[:-1] + np.diff(kv) / kv.dtype.type(2)
    R = rmeye(sd.n_basis_functions, device=sd.device, float_type_=sd.dtype)
    sd_new = sd
    for knot in np.asarray(knots_new):
        sd_new, R_knot = _insert_knot_sd(sd_new, knot, recompute_sample_indices=False)
        R = R_knot @ R
    set_sample_indices_(sd_new)
    return sd_new, R


def _refine_grid(grid: SplineGrid, dim_refinement: int, knots_new=None):
    """src/refinement.jl:309-332."""
    assert not grid.is_nurbs(), "Knot insertion not supported for NURBS"
    sd_new, R = _refine_sd(grid.spline_dimensions[dim_refinement - 1], knots_new=knots_new)
    evaluate_dimension_(sd_new)
    return _refine_grid_with(grid, sd_new, dim_refinement, R)


def _refine_grid_with(grid: SplineGrid, sd_new: SplineDimension, dim_refinement: int, R: RefinementMatrix):
    """src/refinement.jl:344-386: refine the control points with ``mult!`` and swap the dimension."""
    cp = obtain(grid.control_points)
    shape = list(cp.shape)
    shape[dim_refinement - 1] += R.m - R.n
    cp_new = jl_zeros(shape, cp.dtype, cp.device)
    mult_(cp_new, (R,), cp, (dim_refinement,))
    dims = tuple(sd_new if d == dim_refinement - 1 else sd for d, sd in enumerate(grid.spline_dimensions))
    return grid.replace(spline_dimensions=dims, control_points=DefaultControlPoints(cp_new)), R


def add_default_local_refinement(grid: SplineGrid) -> SplineGrid:
    """``add_default_local_refinement`` -- src/control_points.jl:364-425: bisect every knot span in every
    dimension and append one level (no active control points yet) to the hierarchy."""
    nin = grid.Nin
    dims_new, mats = [], []
    for sd in grid.spline_dimensions:
        sd_new, R = _refine_sd(sd)
        build_(sd_new)                      # evaluate!.(spline_dimensions_new), :388
        dims_new.append(sd_new)
        mats.append(R)
    cp_old = obtain(grid.control_points)
    nout = cp_old.shape[-1]
    cp_new = jl_zeros(tuple(sd.n_basis_functions for sd in dims_new) + (nout,), cp_old.dtype, cp_old.device)
    dims_refinement = tuple(range(1, nin + 1))
    mult_(cp_new, tuple(mats), cp_old, dims_refinement)
    lr = LocalRefinement(dims_refinement, mats, jl_zeros((0, nin), torch.int32, cp_old.device),
                         jl_zeros((0, nout), cp_old.dtype, cp_old.device))
    if isinstance(grid.control_points, LocallyRefinedControlPoints):
        cps = grid.control_points
        cps.control_points_refined.append(cp_new)
        cps.local_refinements.append(lr)
    else:
        cps = LocallyRefinedControlPoints([cp_old, cp_new], [base_local_refinement(cp_old), lr])
    eval_new = grid.eval
    return grid.replace(spline_dimensions=tuple(dims_new), control_points=cps, eval=eval_new)


def error_informed_local_refinement_(grid: SplineGrid, error: torch.Tensor, threshold_factor: float = 1.0) -> None:
    """``error_informed_local_refinement!`` -- src/control_points.jl:541-575: map the error back onto the
    finest control grid with the adjoint (K4), sum over outputs, activate every control point above
    ``threshold_factor * mean``."""
    assert tuple(error.shape) == tuple(grid.eval.shape), "The error array must have the same size as the eval array."
    from . import device_setup as D
    cp_err = torch.zeros_like(obtain(grid.control_points))
    evaluate_adjoint_(grid, eval=error, control_points=cp_err)
    # sum over outputs, threshold = factor * mean, findall(> threshold) in column-major order, CartesianIndex -> (n, Nin)
    # Int32 matrix: all on the device (sg_error_flags, sg_compact_flags, sg_indices_from_linear)
    flags = D.error_flags(cp_err, float(threshold_factor))
    hit = D.compact_flags(flags)
    idx = D.indices_from_linear(hit, cp_err.shape[:-1])
    activate_local_refinement_(grid.control_points, idx)
