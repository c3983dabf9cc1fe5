"""AD rules for the opaque boundary (scope row f2).

The reference is differentiated THROUGH its KernelAbstractions kernels by Enzyme (``ext/SplineGridsEnzymeExt.jl``,
``test/test_EnzymeExt.jl:24-49``, README.md:22).  Behind a C ABI the kernels are opaque to any AD tool, so the rules are
stated explicitly.  ``evaluate!`` is linear in the control points (src/spline_grid.jl:130-182), hence

* forward rule (JVP):   d(eval) = ``evaluate!`` applied to the tangent control points,
* reverse rule (VJP):   d(control_points) = ``evaluate_adjoint!`` applied to the cotangent of ``eval`` (src/adjoint.jl:52-83).

Here they are a ``torch.autograd.Function`` (the host mirror's AD system); the Julia shim carries the same two rules as
``EnzymeRules.forward`` / ``augmented_primal`` + ``reverse`` (``julia/SplineGridsB200.jl``).  ``make_zero_`` mirrors the
reference's only Enzyme method, ``Enzyme.make_zero!(::SplineGrid)`` (ext/SplineGridsEnzymeExt.jl:5-11)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .arrays import as_colmajor, jl_empty, reshape_colmajor
from .spline_grid import SplineGrid, evaluate_, evaluate_adjoint_


class _Evaluate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, control_points_flat: torch.Tensor, grid: SplineGrid, der):
        cp = reshape_colmajor(control_points_flat.detach().contiguous().reshape(-1), grid.control_points.shape)
        out = jl_empty(grid.eval.shape, grid.dtype, grid.device)
        evaluate_(grid, control_points=cp, eval=out, derivative_order=der)
        ctx.grid, ctx.der, ctx.n = grid, der, control_points_flat.numel()
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        grid = ctx.grid
        g = as_colmajor(grad_out.detach())
        cp_grad = jl_empty(grid.control_points.shape, grid.dtype, grid.device)
        evaluate_adjoint_(grid, eval=g, control_points=cp_grad, derivative_order=ctx.der)
        flat = cp_grad.permute(*reversed(range(cp_grad.dim()))).reshape(-1)      # column-major flattening (Julia's vec)
        return flat, None, None

    @staticmethod
    def jvp(ctx, tangent_flat: torch.Tensor, _grid, _der):
        grid = ctx.grid
        t = reshape_colmajor(tangent_flat.detach().contiguous().reshape(-1), grid.control_points.shape)
        out = jl_empty(grid.eval.shape, grid.dtype, grid.device)
        evaluate_(grid, control_points=t, eval=out, derivative_order=ctx.der)
        return out


def evaluate_autograd(grid: SplineGrid, control_points_flat: torch.Tensor,
                      derivative_order: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Differentiable ``evaluate!``: ``control_points_flat`` is the column-major flattening of the control-point array
    (the ``control_points_flat`` of test/test_EnzymeExt.jl:24-29 and ext/SplineGridsLinearMapsExt.jl); returns a NEW
    array of the shape of ``grid.eval``.  Gradients flow to ``control_points_flat`` through ``evaluate_adjoint!``."""
    assert not grid.is_nurbs(), "the reverse rule is evaluate_adjoint!, which the reference does not define for NURBS"
    der = tuple(int(d) for d in derivative_order) if derivative_order is not None else (0,) * grid.Nin
    assert control_points_flat.numel() == grid.control_points.obtain().numel()
    return _Evaluate.apply(control_points_flat, grid, der)


def make_zero_(grid: SplineGrid) -> None:
    """``Enzyme.make_zero!(::SplineGrid)`` -- ext/SplineGridsEnzymeExt.jl:5-11: zero ``eval`` and every dimension's table
    (used on the SHADOW grid of a ``Duplicated`` argument)."""
    grid.eval.zero_()
    for sd in grid.spline_dimensions:
        sd.eval.zero_()
