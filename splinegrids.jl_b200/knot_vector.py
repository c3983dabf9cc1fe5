"""``KnotVector`` -- mirrors src/knot_vector.jl of the reference.

Host-side set-up (O(n_knots)); the expanded vector ``knots_all`` is produced on the device by the
K9 kernel (``sg_expand_knot_vector``), as the reference does at src/knot_vector.jl:29-37.
``knot_values`` / ``multiplicities`` are additionally kept on the host because every refinement
operation of the reference reads them on the CPU anyway (src/refinement.jl:114-131, :264-279).
"""
from __future__ import annotations

from fractions import Fraction
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .arrays import NP_OF, float_type, require_cuda


def julia_range(a, b, n: int, np_dtype) -> np.ndarray:
    """``float_type.(range(a, b; length=n))``: Julia's ranges use twice-precision stepping, i.e.
    (to within the last bit of Float64) the correctly rounded ``a + i*(b-a)/(n-1)``."""
    if n == 1:
        return np.array([a], dtype=np_dtype)
    lo, hi = Fraction(float(a)), Fraction(float(b))
    step = (hi - lo) / (n - 1)
    return np.array([float(lo + i * step) for i in range(n)], dtype=np.float64).astype(np_dtype)


class KnotVector:
    """``KnotVector(knot_values, multiplicities)`` -- src/knot_vector.jl:11-50.

    Fields: ``knot_values``, ``multiplicities`` (device tensors, Int32 multiplicities), ``knots_all``
    (device, built by K9), ``extent``; ``knot_values_host`` / ``multiplicities_host`` numpy copies.
    """

    def __init__(self, knot_values, multiplicities, device=None):
        kv = knot_values.detach().cpu().numpy() if isinstance(knot_values, torch.Tensor) else np.asarray(knot_values)
        mu = multiplicities.detach().cpu().numpy() if isinstance(multiplicities, torch.Tensor) \
            else np.asarray(multiplicities)
        # validation first (no device needed), messages as in src/knot_vector.jl:22-26
        assert len(kv) == len(mu), "knot_values and multiplicities must be of the same length."
        assert np.array_equal(kv, np.sort(kv)), "knot_values must be sorted."
        assert len(np.unique(kv)) == len(kv), "knot_values must be unique."
        if kv.dtype not in (np.float32, np.float64):
            kv = kv.astype(np.float64)
        self.knot_values_host = kv
        self.multiplicities_host = mu.astype(np.int32)
        self.extent: Tuple[float, float] = (kv[0].item(), kv[-1].item())
        dev = require_cuda(device if device is not None else
                           (knot_values.device if isinstance(knot_values, torch.Tensor) and knot_values.is_cuda else None))
        dtype = float_type(kv.dtype)
        self.knot_values = torch.from_numpy(kv).to(dev)
        self.multiplicities = torch.from_numpy(self.multiplicities_host).to(dev)
        self.knots_all = torch.empty(int(self.multiplicities_host.sum()), dtype=dtype, device=dev)
        with torch.cuda.device(dev):
            fn = getattr(_lib.lib(), "sg_expand_knot_vector_" + _lib.suffix(dtype))
            _lib.check(fn(_lib.ptr(self.knots_all), _lib.ptr(self.knot_values), _lib.ptr(self.multiplicities),
                          _lib.C.c_int64(len(kv)), _lib.stream_ptr(dev)), "sg_expand_knot_vector")

    @property
    def dtype(self) -> torch.dtype:
        return self.knots_all.dtype

    @property
    def device(self) -> torch.device:
        return self.knots_all.device

    @classmethod
    def clamped(cls, n_basis_functions: int, degree: int, extent=(0, 1), distribution: str = "equispaced",
                device=None, float_type_="Float32", rng: Optional[np.random.Generator] = None) -> "KnotVector":
        """``KnotVector(n_basis_functions, degree; extent, distribution, backend, float_type)`` --
        src/knot_vector.jl:75-113: clamped knot vector, end multiplicities ``degree+1``."""
        assert n_basis_functions - degree >= 1
        np_dtype = NP_OF[float_type(float_type_)]
        n_knot_values = n_basis_functions - degree + 1
        if distribution == "random":
            rng = np.random.default_rng() if rng is None else rng
            kv = np.cumsum(rng.random(n_knot_values).astype(np_dtype)).astype(np_dtype)
            kv -= kv[0]
            kv /= kv[-1] - kv[0]
            kv *= np_dtype(extent[1] - extent[0])
            kv += np_dtype(extent[0])
        elif distribution == "equispaced":
            kv = julia_range(extent[0], extent[1], n_knot_values, np_dtype)
        else:
            raise ValueError(f"Unsupported knot distribution type {distribution}.")
        mult = np.ones(n_knot_values, dtype=np.int32)
        mult[0] = degree + 1
        mult[-1] = degree + 1
        return cls(kv, mult, device=device)
