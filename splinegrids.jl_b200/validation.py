"""Host-side validation with the reference's error conventions -- src/validation.jl.

Error texts are pinned by the reference's tests (test/test_spline_grid.jl:75-82,
test/test_nurbs_grid.jl:21-22); one ``logging.error`` record is emitted per offending dimension,
like the reference's ``@error`` records.
"""
from __future__ import annotations

import logging
from typing import Sequence

logger = logging.getLogger("splinegrids_b200")


class SplineGridsError(Exception):
    """Stands in for Julia's ``ErrorException`` raised by ``error(...)``."""


def validate_spline_grid(spline_dimensions, control_points_shape, weights_shape, eval_shape) -> None:
    """src/validation.jl:1-50."""
    errors = False
    nin_sd = len(spline_dimensions)
    nin_cp = len(control_points_shape) - 1
    nin_eval = len(eval_shape) - 1
    if weights_shape is None:
        nins = (nin_sd, nin_cp, nin_eval)
        if len(set(nins)) != 1:
            logger.error("The number of input dimensions from the spline dimensions, control_points, and eval "
                         f"must agree, got {nins} respectively.")
            errors = True
    else:
        nins = (nin_sd, nin_cp, nin_eval, len(weights_shape))
        if len(set(nins)) != 1:
            logger.error("The number of input dimensions from the spline dimensions, control_points, eval and "
                         f"weights must agree, got {nins} respectively.")
            errors = True
    nout_cp, nout_eval = control_points_shape[-1], eval_shape[-1]
    if nout_cp != nout_eval:
        logger.error(f"The number of output dimensions from the control points {nout_cp} and eval {nout_eval} "
                     "must agree.")
        errors = True
    cp_sd = tuple(sd.n_basis_functions for sd in spline_dimensions)
    cp_cp = tuple(control_points_shape[:-1])
    if cp_sd != cp_cp:
        logger.error(f"The control point grid sizes from the spline dimensions {cp_sd} and the control points "
                     f"{cp_cp} must agree.")
        errors = True
    if errors:
        raise SplineGridsError("Errors encountered when validating spline grid constructor inputs.")


def validate_partial_derivatives(spline_dimensions, derivative_order: Sequence[int], is_nurbs: bool = False) -> None:
    """src/validation.jl:52-88."""
    if is_nurbs:
        if any(int(d) != 0 for d in derivative_order):
            raise SplineGridsError("Computing derivatives of NURBS is currently not supported.")
        return
    errors = False
    for i, (sd, der) in enumerate(zip(spline_dimensions, derivative_order), start=1):
        mdo = sd.max_derivative_order
        if not (0 <= der <= mdo):
            errors = True
            logger.error(f"The maximum derivative order available for spline dimension {i} is {mdo}, got {der}.")
    if errors:
        raise SplineGridsError(
            "Invalid derivative order(s) supplied. If you want to evaluate (higher order) derivatives, "
            "specify this at construction as SplineDimension(...; max_derivative_order).")


def validate_mult_input(Y_shape, As, B_shape, dims_refinement) -> None:
    """src/validation.jl:90-107."""
    dims_refinement = list(dims_refinement)
    assert len(set(dims_refinement)) == len(dims_refinement), "Refinement dimensions must be unique."
    assert len(As) == len(dims_refinement), \
        "There must be exactly one refinement dimension per refinement matrix."
    for dim, (sb, sy) in enumerate(zip(B_shape, Y_shape), start=1):
        if dim in dims_refinement:
            A = As[dims_refinement.index(dim)]
            if not (A.m == sy and A.n == sb):
                raise SplineGridsError(
                    f"Size of refinement matrix does not match `B` and `Y` along refinement dimension {dim}.")
        elif sb != sy:
            raise SplineGridsError(f"`B` and `Y` don't have the same size along dimension {dim}.")
