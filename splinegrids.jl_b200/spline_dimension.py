"""``SplineDimension`` -- mirrors src/spline_dimension.jl of the reference: one dimension's basis
functions and how they are sampled.  The tables are built ON DEVICE by the K1/K2 kernels."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .arrays import NP_OF, float_type, jl_zeros, require_cuda
from .config import after_launch
from .knot_vector import KnotVector, julia_range

C = _lib.C


class SplineDimension:
    """Fields as in the reference struct (src/spline_dimension.jl:25-66): ``degree``,
    ``max_derivative_order``, ``knot_vector``, ``sample_points[n]``, ``sample_indices[n]`` (Int32,
    1-based knot span per sample), ``eval[n, degree+1, max_derivative_order+1]``.
    ``eval_prev`` (the reference's scratch array) is not needed by the register-resident K2 kernel
    and is ``None``.

    ``SplineDimension(n_basis_functions, degree, n_sample_points; max_derivative_order, knot_vector,
    float_type, extent, distribution)`` -- src/spline_dimension.jl:97-156.
    """

    def __init__(self, n_basis_functions: int, degree: int, n_sample_points: int, *,
                 max_derivative_order: int = 0, knot_vector: Optional[KnotVector] = None, device=None,
                 float_type="Float32", **kwargs):
        assert 0 <= max_derivative_order <= degree, \
            "The max_degree must be positive and derivatives order higher than `degree` are all 0."
        if knot_vector is None:
            knot_vector = KnotVector.clamped(n_basis_functions, degree, device=device, float_type_=float_type,
                                             **kwargs)
        else:
            assert knot_vector.knots_all.numel() == n_basis_functions + degree + 1, "Incompatible knot vector supplied."
        dev = knot_vector.device
        dtype = knot_vector.dtype
        sp = julia_range(knot_vector.extent[0], knot_vector.extent[1], n_sample_points, NP_OF[dtype])
        self._init_fields(degree, max_derivative_order, knot_vector, torch.from_numpy(sp).to(dev),
                          torch.zeros(n_sample_points, dtype=torch.int32, device=dev),
                          jl_zeros((n_sample_points, degree + 1, max_derivative_order + 1), dtype, dev))
        build_(self)     # fused K1+K2 == set_sample_indices!(s); evaluate!(s)   (:153-154)

    def _init_fields(self, degree, max_derivative_order, knot_vector, sample_points, sample_indices, eval_):
        self.degree = int(degree)
        self.max_derivative_order = int(max_derivative_order)
        self.knot_vector = knot_vector
        self.sample_points = sample_points
        self.sample_indices = sample_indices
        self.eval = eval_
        self.eval_prev = None

    @classmethod
    def from_fields(cls, degree, max_derivative_order, knot_vector, sample_points, sample_indices, eval_) -> "SplineDimension":
        """The reference's inner (struct) constructor, src/spline_dimension.jl:40-65: takes the arrays as
        they are (e.g. user-supplied, possibly unsorted sample points) and computes nothing."""
        self = cls.__new__(cls)
        self._init_fields(degree, max_derivative_order, knot_vector, sample_points, sample_indices, eval_)
        return self

    def with_knot_vector(self, knot_vector: KnotVector) -> "SplineDimension":
        """``setproperties(spline_dimension; knot_vector)`` (src/refinement.jl:172).  The reference shares
        ``sample_indices``/``eval`` between the old and the new object, so refining silently invalidates the old
        one; here the new object gets its own copies (a few KB) and the old dimension stays usable."""
        return SplineDimension.from_fields(self.degree, self.max_derivative_order, knot_vector, self.sample_points,
                                           self.sample_indices.clone(), self.eval.clone())

    # -- small helpers of src/utils.jl:32-56 ------------------------------------------------------
    @property
    def n_basis_functions(self) -> int:
        return self.knot_vector.knots_all.numel() - self.degree - 1

    @property
    def n_sample_points(self) -> int:
        return self.sample_points.numel()

    @property
    def dtype(self) -> torch.dtype:
        return self.eval.dtype

    @property
    def device(self) -> torch.device:
        return self.eval.device

    def __repr__(self):
        return (f"SplineDimension(n_basis_functions={self.n_basis_functions}, degree={self.degree}, "
                f"n_sample_points={self.n_sample_points}, max_derivative_order={self.max_derivative_order}, "
                f"{self.dtype})")


def _common(sd: SplineDimension):
    kn = sd.knot_vector.knots_all
    return kn, C.c_int64(kn.numel()), C.c_int64(sd.n_sample_points)


def set_sample_indices_(sd: SplineDimension) -> None:
    """``set_sample_indices!`` (K1) -- src/utils.jl:19-29."""
    sd.__dict__["_version"] = sd.__dict__.get("_version", 0) + 1   # invalidates adjoint plans (spline_grid.py)
    kn, n_knots, n = _common(sd)
    with torch.cuda.device(sd.device):
        fn = getattr(_lib.lib(), "sg_span_indices_" + _lib.suffix(sd.dtype))
        _lib.check(fn(_lib.ptr(sd.sample_indices), _lib.ptr(sd.sample_points), n, _lib.ptr(kn), n_knots,
                      C.c_int(sd.degree), _lib.stream_ptr(sd.device)), "sg_span_indices")
    after_launch(sd.device)


def evaluate_dimension_(sd: SplineDimension) -> None:
    """``evaluate!(::SplineDimension)`` (K2) -- src/spline_dimension.jl:231-242."""
    sd.__dict__["_version"] = sd.__dict__.get("_version", 0) + 1   # invalidates adjoint plans (spline_grid.py)
    kn, n_knots, n = _common(sd)
    with torch.cuda.device(sd.device):
        fn = getattr(_lib.lib(), "sg_basis_tables_" + _lib.suffix(sd.dtype))
        _lib.check(fn(_lib.ptr(sd.eval), _lib.ptr(kn), n_knots, _lib.ptr(sd.sample_points),
                      _lib.ptr(sd.sample_indices), n, C.c_int(sd.degree), C.c_int(sd.max_derivative_order),
                      _lib.stream_ptr(sd.device)), "sg_basis_tables")
    after_launch(sd.device)


def build_(sd: SplineDimension) -> None:
    """Fused K1+K2 (one launch): span lookup + Cox-de Boor tables."""
    sd.__dict__["_version"] = sd.__dict__.get("_version", 0) + 1   # invalidates adjoint plans (spline_grid.py)
    kn, n_knots, n = _common(sd)
    with torch.cuda.device(sd.device):
        fn = getattr(_lib.lib(), "sg_dimension_build_" + _lib.suffix(sd.dtype))
        _lib.check(fn(_lib.ptr(sd.sample_indices), _lib.ptr(sd.eval), _lib.ptr(kn), n_knots,
                      _lib.ptr(sd.sample_points), n, C.c_int(sd.degree), C.c_int(sd.max_derivative_order),
                      _lib.stream_ptr(sd.device)), "sg_dimension_build")
    after_launch(sd.device)


def decompress(sd: SplineDimension, derivative_order: int = 0) -> torch.Tensor:
    """``decompress`` (K10) -- src/spline_dimension.jl:251-272: dense ``(n_sample_points, n_basis)`` matrix."""
    assert derivative_order <= sd.max_derivative_order
    out = jl_zeros((sd.n_sample_points, sd.n_basis_functions), sd.dtype, sd.device)
    with torch.cuda.device(sd.device):
        fn = getattr(_lib.lib(), "sg_decompress_" + _lib.suffix(sd.dtype))
        _lib.check(fn(_lib.ptr(out), _lib.ptr(sd.eval), _lib.ptr(sd.sample_indices), C.c_int64(sd.n_sample_points),
                      C.c_int64(sd.n_basis_functions), C.c_int(sd.degree), C.c_int(derivative_order),
                      _lib.stream_ptr(sd.device)), "sg_decompress")
    after_launch(sd.device)
    return out
