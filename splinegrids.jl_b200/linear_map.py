"""``LinearMap(spline_grid)`` -- mirrors ext/SplineGridsLinearMapsExt.jl:5-52: the grid as a linear operator
flat control points -> flat evaluation, with its adjoint, for iterative least-squares solvers."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .arrays import reshape_colmajor
from .control_points import copyto_, get_n_control_points
from .spline_grid import SplineGrid, evaluate_, evaluate_adjoint_
from .validation import validate_partial_derivatives


class SplineGridLinearMap:
    """``A = SplineGridLinearMap(grid)``; ``A.matvec(cp_flat, out=eval_flat)``; ``A.rmatvec(eval_flat, out=cp_flat)``.
    Flat vectors are column-major device vectors, as in the reference (``vec(...)``)."""

    def __init__(self, spline_grid: SplineGrid, derivative_order: Optional[Sequence[int]] = None):
        self.grid = spline_grid
        self.derivative_order = tuple(derivative_order) if derivative_order is not None else (0,) * spline_grid.Nin
        validate_partial_derivatives(spline_grid.spline_dimensions, self.derivative_order, spline_grid.is_nurbs())
        self.n_cp = get_n_control_points(spline_grid)
        self.shape = (spline_grid.eval.numel(), self.n_cp * spline_grid.Nout)

    def matvec(self, control_points_flat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = self.grid
        if out is None:
            out = torch.empty(self.shape[0], dtype=g.dtype, device=g.device)
        copyto_(g.control_points, reshape_colmajor(control_points_flat, (self.n_cp, g.Nout)))
        evaluate_(g.control_points)
        evaluate_(g, derivative_order=self.derivative_order, eval=reshape_colmajor(out, g.eval.shape))
        return out

    def rmatvec(self, evaluation_flat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = self.grid
        if out is None:
            out = torch.empty(self.shape[1], dtype=g.dtype, device=g.device)
        evaluate_adjoint_(g, derivative_order=self.derivative_order,
                          eval=reshape_colmajor(evaluation_flat, g.eval.shape))
        evaluate_adjoint_(g.control_points)
        copyto_(reshape_colmajor(out, (self.n_cp, g.Nout)), g.control_points)
        return out
