"""``SplineGrid`` / ``NURBSGrid`` and the hot path ``evaluate!`` / ``evaluate_adjoint!`` -- mirrors
src/spline_grid.jl and src/adjoint.jl:1-83 of the reference.  Julia's ``f!`` is spelled ``f_`` here."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .arrays import is_colmajor, jl_ones, jl_zeros, to_device
from .config import adjoint_plans, after_launch
from .control_points import (AbstractControlPoints, DefaultControlPoints, LocallyRefinedControlPoints,
                             evaluate_adjoint_control_points_, evaluate_control_points_, obtain)
from .spline_dimension import SplineDimension, evaluate_dimension_
from .validation import validate_partial_derivatives, validate_spline_grid

C = _lib.C


def _unit_cp_grid(cp_shape: Sequence[int], np_dtype) -> np.ndarray:
    """``set_unit_cp_grid!`` -- src/utils.jl:1-15: output n = coordinate along input dimension n."""
    cp = np.zeros(tuple(cp_shape), dtype=np_dtype, order="F")
    nin = len(cp_shape) - 1
    for n in range(min(nin, cp_shape[-1])):
        shape = [1] * nin
        shape[n] = cp_shape[n]
        cp[..., n] = np.linspace(0, 1, cp_shape[n]).reshape(shape)
    return cp


class SplineGrid:
    """``SplineGrid(spline_dimensions, Nout)`` -- src/spline_grid.jl:70-94.

    Fields (src/spline_grid.jl:19-58): ``spline_dimensions`` (tuple), ``control_points``
    (``DefaultControlPoints`` or ``LocallyRefinedControlPoints``), ``weights`` (``None`` or an array of the
    control-grid shape: NURBS), ``eval`` of shape ``(n_1..n_D, Nout)``.
    """

    def __init__(self, spline_dimensions, Nout: Optional[int] = None, *, control_points=None, weights=None,
                 eval_=None):
        if isinstance(spline_dimensions, SplineDimension):            # src/spline_grid.jl:111-117
            spline_dimensions = (spline_dimensions,)
        self.spline_dimensions: Tuple[SplineDimension, ...] = tuple(spline_dimensions)
        sd0 = self.spline_dimensions[0]
        dev, dtype = sd0.device, sd0.dtype
        if control_points is None:
            assert Nout is not None
            cp_shape = tuple(sd.n_basis_functions for sd in self.spline_dimensions) + (int(Nout),)
            control_points = DefaultControlPoints(
                to_device(_unit_cp_grid(cp_shape, np.float64), dtype=dtype, device=dev))
        elif isinstance(control_points, torch.Tensor):
            control_points = DefaultControlPoints(control_points)
        self.control_points = control_points
        nout = control_points.Nout
        self.weights = weights
        self.eval = eval_ if eval_ is not None else jl_zeros(self.sample_grid_size + (nout,), dtype, dev)
        validate_spline_grid(self.spline_dimensions, control_points.shape,
                             None if weights is None else tuple(weights.shape), tuple(self.eval.shape))

    # -- size helpers (src/utils.jl:32-56) --------------------------------------------------------
    @property
    def Nin(self) -> int:
        return len(self.spline_dimensions)

    @property
    def Nout(self) -> int:
        return self.control_points.Nout

    @property
    def sample_grid_size(self) -> Tuple[int, ...]:
        return tuple(sd.n_sample_points for sd in self.spline_dimensions)

    @property
    def control_point_grid_size(self) -> Tuple[int, ...]:
        return tuple(sd.n_basis_functions for sd in self.spline_dimensions)

    @property
    def dtype(self):
        return self.eval.dtype

    @property
    def device(self):
        return self.eval.device

    def is_nurbs(self) -> bool:
        return self.weights is not None

    def replace(self, **kw) -> "SplineGrid":
        """``setproperties(spline_grid; ...)``: a new grid sharing all other fields."""
        g = SplineGrid.__new__(SplineGrid)
        g.spline_dimensions = tuple(kw.get("spline_dimensions", self.spline_dimensions))
        g.control_points = kw.get("control_points", self.control_points)
        g.weights = kw.get("weights", self.weights)
        g.eval = kw.get("eval", self.eval)
        return g

    def __repr__(self):
        kind = "NURBSGrid" if self.is_nurbs() else "SplineGrid"
        shape_name = {1: "curve", 2: "surface", 3: "volume"}.get(self.Nin, f"hyper ({self.Nin}) volume")
        return (f"{kind} {shape_name} with outputs in R^{self.Nout} ({self.dtype}); control grid "
                f"{self.control_point_grid_size}, degrees {tuple(sd.degree for sd in self.spline_dimensions)}, "
                f"samples {self.sample_grid_size}")


def NURBSGrid(spline_dimensions, Nout: int) -> SplineGrid:
    """``NURBSGrid(spline_dimensions, Nout)`` -- src/spline_grid.jl:100-108: a SplineGrid with weights of ones."""
    g = SplineGrid(spline_dimensions, Nout)
    g.weights = jl_ones(g.control_point_grid_size, g.dtype, g.device)
    return g


def _grid_call_args(grid: SplineGrid, derivative_order):
    sds = grid.spline_dimensions
    return (C.c_int(len(sds)), _lib.i64_array([sd.n_sample_points for sd in sds]),
            _lib.i64_array([sd.n_basis_functions for sd in sds]), C.c_int(grid.Nout),
            _lib.ptr_array([sd.eval for sd in sds]), _lib.ptr_array([sd.sample_indices for sd in sds]),
            _lib.int_array([sd.degree for sd in sds]), _lib.int_array([sd.max_derivative_order for sd in sds]),
            _lib.int_array(derivative_order))


def _check_arrays(grid: SplineGrid, control_points, eval_):
    cp = obtain(control_points)
    assert tuple(cp.shape) == tuple(grid.control_points.shape)       # src/spline_grid.jl:208
    assert tuple(eval_.shape) == tuple(grid.eval.shape)              # src/spline_grid.jl:209
    assert cp.dtype == grid.dtype and eval_.dtype == grid.dtype
    assert is_colmajor(cp) and is_colmajor(eval_), "control_points and eval must be dense column-major arrays"
    assert cp.device == grid.device and eval_.device == grid.device
    return cp


# Prepared calls: a fitting loop calls evaluate! / evaluate_adjoint! thousands of times with the same arrays, and on
# a thin slab of a sharded grid the kernels take only tens of microseconds, so the host side of a repeated call must
# be a dictionary look-up plus one foreign call.  The key holds everything the marshalled arguments depend on (array
# addresses, shapes are implied by the validated first call); a new array, derivative order or rebuilt table misses.
_PREPARED_MAX = 64


def _prepared_key(kind, grid, der, cp, eval_):
    return (kind, der, cp.data_ptr(), eval_.data_ptr(), cp.shape, cp.stride(), cp.dtype, eval_.shape, eval_.stride(),
            eval_.dtype, 0 if grid.weights is None else grid.weights.data_ptr(),
            tuple(sd.eval.data_ptr() for sd in grid.spline_dimensions),
            tuple(sd.sample_indices.data_ptr() for sd in grid.spline_dimensions),
            tuple(sd.__dict__.get("_version", 0) for sd in grid.spline_dimensions))


def _prepared_store(grid, key, value):
    cache = grid.__dict__.setdefault("_prepared", {})
    if len(cache) >= _PREPARED_MAX:
        cache.clear()
    cache[key] = value


def evaluate_(obj, *, derivative_order: Optional[Sequence[int]] = None, control_points=None, eval=None) -> None:
    """``evaluate!`` -- dispatches like the reference's methods:

    * ``SplineGrid``: the hot path, src/spline_grid.jl:200-230 (K3).  ``control_points`` may be a
      control-point object or a raw/reshaped dense array, ``eval`` any same-size array.
    * ``SplineDimension``: src/spline_dimension.jl:231-242 (K2).
    * control points: src/control_points.jl:313-349 (K5+K7; no-op for dense control points).
    """
    if isinstance(obj, SplineDimension):
        return evaluate_dimension_(obj)
    if isinstance(obj, AbstractControlPoints):
        return evaluate_control_points_(obj)
    grid: SplineGrid = obj
    nin = grid.Nin
    der = tuple(int(d) for d in derivative_order) if derivative_order is not None else (0,) * nin
    control_points = grid.control_points if control_points is None else control_points
    eval_ = grid.eval if eval is None else eval
    cp = obtain(control_points)
    key = _prepared_key("fwd", grid, der, cp, eval_)
    prep = grid.__dict__.get("_prepared", {}).get(key)
    if prep is None:
        assert len(der) == nin
        validate_partial_derivatives(grid.spline_dimensions, der, is_nurbs=grid.is_nurbs())
        cp = _check_arrays(grid, control_points, eval_)
        fn = getattr(_lib.lib(), "sg_evaluate_" + _lib.suffix(grid.dtype))
        # the tuple keeps the marshalled ctypes arrays (and through them nothing else) alive
        prep = (fn, (_lib.ptr(eval_), *_grid_call_args(grid, der), _lib.ptr(cp), _lib.ptr(grid.weights)), grid.device.index)
        _prepared_store(grid, key, prep)
    fn, args, dev_index = prep
    if torch.cuda.current_device() == dev_index:
        _lib.check(fn(*args, _lib.stream_ptr(grid.device)), "sg_evaluate")
    else:
        with torch.cuda.device(grid.device):
            _lib.check(fn(*args, _lib.stream_ptr(grid.device)), "sg_evaluate")
    after_launch(grid.device)
    return None


def evaluate_multi_(grid: SplineGrid, derivative_orders: Sequence[Sequence[int]], evals: Sequence[torch.Tensor], *,
                    control_points=None) -> None:
    """Several ``evaluate!`` calls on the same control points as ONE library call (``sg_evaluate_multi``): what the
    reference's callers write as back-to-back calls with different ``derivative_order`` and ``eval`` arrays
    (docs/src/examples_optics.md:189-191, docs/src/examples_pde.md:69-72).  ``evals[q]`` receives
    ``evaluate!(grid; derivative_order = derivative_orders[q], eval = evals[q])``."""
    nin = grid.Nin
    ders = [tuple(int(d) for d in der) for der in derivative_orders]
    assert len(ders) == len(evals) and len(ders) >= 1
    control_points = grid.control_points if control_points is None else control_points
    for der, ev in zip(ders, evals):
        assert len(der) == nin
        validate_partial_derivatives(grid.spline_dimensions, der, is_nurbs=grid.is_nurbs())
        cp = _check_arrays(grid, control_points, ev)
    fn = getattr(_lib.lib(), "sg_evaluate_multi_" + _lib.suffix(grid.dtype))
    flat = [d for der in ders for d in der]
    args = _grid_call_args(grid, (0,) * nin)[:-1]                     # without the single derivative order
    with torch.cuda.device(grid.device):
        _lib.check(fn(_lib.ptr_array(list(evals)), C.c_int(len(ders)), _lib.int_array(flat), *args, _lib.ptr(cp),
                      _lib.ptr(grid.weights), _lib.stream_ptr(grid.device)), "sg_evaluate_multi")
    after_launch(grid.device)
    return None


def _workspace(grid: SplineGrid) -> torch.Tensor:
    """Adjoint workspace owned by the grid (its raw pointer is baked into prepared calls and captured CUDA graphs, so
    its lifetime is the grid's).  One grid must not run evaluate_adjoint! on two streams at the same time."""
    sds = grid.spline_dimensions
    nbytes = int(_lib.lib().sg_evaluate_adjoint_workspace_bytes(
        C.c_int(len(sds)), _lib.i64_array([sd.n_sample_points for sd in sds]),
        _lib.i64_array([sd.n_basis_functions for sd in sds]), C.c_int(grid.Nout),
        _lib.int_array([sd.degree for sd in sds]), C.c_int(grid.eval.element_size()),
        C.c_int(1 if grid.is_nurbs() else 0)))
    ws = grid.__dict__.get("_adjoint_ws")
    if ws is None or ws.numel() < nbytes or ws.device != grid.device:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=grid.device)
        grid.__dict__["_adjoint_ws"] = ws
    return ws


class _AdjointPlan:
    """Owner of one ``sg_adjoint_plan`` handle (include/splinegrids_b200.h): destroyed with the prepared call."""

    def __init__(self, grid: SplineGrid, der):
        self.handle = C.c_void_p(0)
        fn = getattr(_lib.lib(), "sg_adjoint_plan_create_" + _lib.suffix(grid.dtype))
        with torch.cuda.device(grid.device):
            _lib.check(fn(C.byref(self.handle), *_grid_call_args(grid, der), C.c_int(1 if grid.is_nurbs() else 0),
                          _lib.stream_ptr(grid.device)), "sg_adjoint_plan_create")

    def info(self):
        mono, fused, rows = C.c_int(0), C.c_int(0), C.c_int(0)
        _lib.check(_lib.lib().sg_adjoint_plan_info(self.handle, C.byref(mono), C.byref(fused), C.byref(rows)), "sg_adjoint_plan_info")
        return {"monotone": bool(mono.value), "fused_tables_fit": bool(fused.value), "rows2_max": rows.value}

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().sg_adjoint_plan_destroy(self.handle)
                self.handle = C.c_void_p(0)
        except Exception:
            pass


def evaluate_adjoint_(obj, *, derivative_order: Optional[Sequence[int]] = None, control_points=None,
                      eval=None, allow_nurbs: bool = False, _retry: bool = False, _push=None) -> None:
    """``evaluate_adjoint!`` -- src/adjoint.jl:52-83 (K4) for a grid; src/adjoint.jl:172-205 for control
    points.  Overwrites ``control_points`` (zero fill first, :61) with the adjoint applied to ``eval``.

    The reference defines no method for NURBS grids (``MethodError``).  ``allow_nurbs=True`` opts in to
    this package's extension: the transpose of the fixed-weights rational map (parity unpinned)."""
    if isinstance(obj, AbstractControlPoints):
        return evaluate_adjoint_control_points_(obj)
    grid: SplineGrid = obj
    if grid.is_nurbs() and not allow_nurbs:
        raise TypeError("MethodError: no method matching evaluate_adjoint!(::NURBSGrid): "
                        "Adjoint evaluation not supported for NURBS.")
    nin = grid.Nin
    der = tuple(int(d) for d in derivative_order) if derivative_order is not None else (0,) * nin
    control_points = grid.control_points if control_points is None else control_points
    eval_ = grid.eval if eval is None else eval
    cp = obtain(control_points)
    # _push = (tag, peer_ptrs, world, rank, k0, np, max_planes[, keep_local, multicast_ptr, (k0s, nps)]): fused gradient push (distributed.PeerGradientExchange)
    key = _prepared_key(("adj", adjoint_plans()) if _push is None else ("adj_push", _push[0], adjoint_plans()), grid, der, cp, eval_)
    prep = grid.__dict__.get("_prepared", {}).get(key)
    if prep is None:
        validate_partial_derivatives(grid.spline_dimensions, der, is_nurbs=grid.is_nurbs())
        cp = _check_arrays(grid, control_points, eval_)
        ws = _workspace(grid)
        plan = None
        if adjoint_plans():
            plan = _AdjointPlan(grid, der)
            fn = getattr(_lib.lib(), "sg_evaluate_adjoint_planned_" + _lib.suffix(grid.dtype))
            args = (plan.handle, _lib.ptr(cp), _lib.ptr(eval_), _lib.ptr(grid.weights), _lib.ptr(ws), C.c_size_t(ws.numel()))
            if _push is None:
                args = args + (C.c_void_p(0), C.c_int(0), C.c_int(0), C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int(1),
                               C.c_void_p(0))
        else:
            args = (_lib.ptr(cp), *_grid_call_args(grid, der), _lib.ptr(eval_), _lib.ptr(grid.weights), _lib.ptr(ws),
                    C.c_size_t(ws.numel()))
            fn = getattr(_lib.lib(), "sg_evaluate_adjoint_" + ("" if _push is None else "push_") + _lib.suffix(grid.dtype))
        if _push is not None:
            _, peer_ptrs, world, rank, k0, np_, max_planes = _push[:7]
            keep_local = _push[7] if len(_push) > 7 else 1
            supports = _push[9] if len(_push) > 9 else None
            if plan is not None and supports is not None:
                # support-plane exchange: planes go to the ranks whose slabs read them only (supports = (k0s, nps) of all ranks)
                fn = getattr(_lib.lib(), "sg_evaluate_adjoint_planned_support_" + _lib.suffix(grid.dtype))
                args = args + (peer_ptrs, C.c_int(world), C.c_int(rank), supports[0], supports[1], C.c_int64(max_planes),
                               C.c_int(keep_local))
            else:
                args = args + (peer_ptrs, C.c_int(world), C.c_int(rank), C.c_int64(k0), C.c_int64(np_), C.c_int64(max_planes),
                               C.c_int(keep_local))
                if plan is not None:                   # NVLS multicast address of the staging buffers (0: none)
                    args = args + (C.c_void_p(_push[8] if len(_push) > 8 else 0),)
        prep = (fn, args, grid.device.index, ws, plan)
        _prepared_store(grid, key, prep)
    fn, args, dev_index = prep[0], prep[1], prep[2]
    if torch.cuda.current_device() == dev_index:
        status = fn(*args, _lib.stream_ptr(grid.device))
    else:
        with torch.cuda.device(grid.device):
            status = fn(*args, _lib.stream_ptr(grid.device))
    if status == -3 and not _retry:            # SG_ERR_WORKSPACE: the pipeline choice (tuning environment) changed since
        grid.__dict__.get("_prepared", {}).pop(key, None)   # the call was prepared -> size the workspace again
        return evaluate_adjoint_(obj, derivative_order=derivative_order, control_points=control_points, eval=eval,
                                 allow_nurbs=allow_nurbs, _retry=True, _push=_push)
    _lib.check(status, "sg_evaluate_adjoint")
    after_launch(grid.device)
    return None
