"""Multi-GPU: slab sharding of the sample grid + gradient all-reduce (new; the reference is
single-device).

The sample grid is split into contiguous slabs along its SLOWEST input axis (dimension ``Nin``); rank
``r`` owns sample rows ``[lo_r, hi_r)`` of that dimension, the full tables of the other dimensions and
a replica of the control points (+ weights).  In column-major ``(n_1..n_D, Nout)`` a slab is ``Nout``
contiguous chunks, so the local output is simply a smaller dense array.

* ``evaluate!`` needs no communication: every output sample depends only on its own table rows and
  a window of the replicated control points.
* ``evaluate_adjoint!`` produces a per-rank partial gradient; the partials are summed with ONE
  all-reduce (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU unit tests of the host logic).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from .arrays import as_colmajor, is_colmajor
from .control_points import obtain
from .spline_dimension import SplineDimension
from .spline_grid import NURBSGrid, SplineGrid, evaluate_, evaluate_adjoint_


def slab_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``range(n)``: the first ``n % world_size`` ranks get one extra row."""
    assert world_size >= 1 and 0 <= rank < world_size and n >= 0
    q, r = divmod(n, world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def slice_dimension(sd: SplineDimension, lo: int, hi: int) -> SplineDimension:
    """Rows ``[lo, hi)`` of a spline dimension (sample points, span indices, table rows) as an independent
    dense dimension object.  Span indices are slices of the global array, so they stay bit-exact."""
    assert 0 <= lo < hi <= sd.n_sample_points
    return SplineDimension.from_fields(
        sd.degree, sd.max_derivative_order, sd.knot_vector,
        sd.sample_points[lo:hi].clone(), sd.sample_indices[lo:hi].clone(), as_colmajor(sd.eval[lo:hi]))


def _contiguous_base(t: torch.Tensor) -> torch.Tensor:
    """A contiguous view sharing memory with a column-major tensor (what collectives need)."""
    assert is_colmajor(t)
    return t if t.dim() <= 1 else t.permute(tuple(range(t.dim() - 1, -1, -1)))


def allreduce_gradient_(control_points: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the replicated control-point gradient over all ranks
    (``ncclAllReduce`` through ``torch.distributed``).  No-op when not running distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(_contiguous_base(control_points), op=dist.ReduceOp.SUM, group=group)
    return control_points


class SlabShardedGrid:
    """This rank's slab of a global spline grid.

    ``global_dims``: the full per-dimension objects (cheap: 1-D tables).  ``local`` is a ``SplineGrid``
    whose last dimension holds rows ``[lo, hi)`` only; its ``eval`` is the local ``(n_1..n_{D-1}, hi-lo, Nout)``
    slab and its control points (and weights) are full replicas.
    """

    def __init__(self, global_dims: Sequence[SplineDimension], Nout: int, rank: int, world_size: int,
                 nurbs: bool = False, group=None):
        self.global_dims = tuple(global_dims)
        self.rank, self.world_size, self.group = rank, world_size, group
        n_last = self.global_dims[-1].n_sample_points
        assert n_last >= world_size, "fewer sample rows along the slowest axis than ranks"
        self.lo, self.hi = slab_bounds(n_last, world_size, rank)
        dims = self.global_dims[:-1] + (slice_dimension(self.global_dims[-1], self.lo, self.hi),)
        self.local: SplineGrid = NURBSGrid(dims, Nout) if nurbs else SplineGrid(dims, Nout)

    @property
    def global_sample_grid_size(self) -> Tuple[int, ...]:
        return tuple(sd.n_sample_points for sd in self.global_dims)

    def evaluate_(self, **kw) -> None:
        """Local ``evaluate!``: no communication."""
        evaluate_(self.local, **kw)

    def evaluate_adjoint_(self, *, control_points: Optional[torch.Tensor] = None, **kw) -> None:
        """Local ``evaluate_adjoint!`` followed by the gradient all-reduce; afterwards every rank holds the
        full gradient, exactly what the single-device call produces (up to summation order)."""
        evaluate_adjoint_(self.local, control_points=control_points, **kw)
        cp = obtain(self.local.control_points if control_points is None else control_points)
        allreduce_gradient_(cp, self.group)
