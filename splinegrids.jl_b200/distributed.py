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

import os

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


def slab_supports(last: SplineDimension, world_size: int) -> Tuple[list, list]:
    """Control planes of the slowest axis every rank's slab touches: ``(k0s, nps)``, 0-based first plane and number of
    planes per rank.  A slab with sample rows ``[lo, hi)`` reads / contributes to the planes ``idx[lo] - p - 1 ..
    idx[hi - 1] - 1`` (0-based; ``idx`` = the 1-based span indices, src/spline_grid.jl:130-182 window offset)."""
    idx = last.sample_indices.cpu().numpy()
    p = last.degree
    k0s, nps = [], []
    for r in range(world_size):
        lo, hi = slab_bounds(last.n_sample_points, world_size, r)
        k0 = int(idx[lo]) - p - 1
        k0s.append(k0)
        nps.append(int(idx[hi - 1]) - k0)
    return k0s, nps


def owned_planes(k0s: Sequence[int], nps: Sequence[int], rank: int) -> Tuple[int, int]:
    """The part ``[lo, hi)`` of rank's support below the first plane of the next rank's support:
    ``[k0_r, min(k0_{r+1}, k0_r + np_r))`` (the last rank owns its whole support).  The owned ranges are disjoint and their
    union is the union of the supports, so they say who contributes which plane when the replicated array is rebuilt."""
    lo = k0s[rank]
    hi = k0s[rank] + nps[rank]
    if rank + 1 < len(k0s):
        hi = min(hi, max(k0s[rank + 1], lo))
    return lo, hi


def allgather_support_planes_(control_points: torch.Tensor, k0s: Sequence[int], nps: Sequence[int], group=None) -> torch.Tensor:
    """After a fit that used the support-plane exchange every rank holds valid control points on ITS support planes only;
    this rebuilds the replicated array: every rank broadcasts the planes it owns (``owned_planes``).  One call at the END of
    a fit (plumbing through ``torch.distributed``), not part of the step."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return control_points
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    for r in range(world):
        lo, hi = owned_planes(k0s, nps, r)
        if hi <= lo:
            continue
        view = control_points[..., lo:hi, :]
        buf = view.contiguous() if r == me else torch.empty(view.shape, dtype=view.dtype, device=view.device)
        dist.broadcast(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
        if r != me:
            view.copy_(buf)
    return control_points


class PeerGradientExchange:
    """Gradient exchange over NVLink peer memory (replaces the all-reduce).

    A slab's partial gradient is non-zero only on the control planes its samples touch.  Every rank pushes
    those planes into its slot of every peer's staging buffer with peer-to-peer stores (fused into the adjoint's last
    kernel, ``sg_evaluate_adjoint_push``, or ``sg_exchange_push``); then ONE kernel per rank
    (``sg_exchange_wait_reduce``) tells the peers that the push has landed (release-store of an exchange counter into
    their flag arrays), waits on the device for all peers' flags and sums the few slots that cover each plane (rank
    order -> deterministic).  No host-side barrier, no collective call: the whole step is plain kernels and can be
    captured in a CUDA graph.  Buffers are ``torch.distributed._symmetric_memory`` allocations (plumbing: allocation and
    the exchange of the peer pointers only); the kernels and the barrier are ours.  Two staging buffers alternate.

    ``mode="support"`` (the halo variant, ``sg_evaluate_adjoint_planned_support`` / ``sg_exchange_wait_reduce_support``):
    ``evaluate!`` on a slab reads only the control planes of the slab's support, so a fitting loop needs the summed
    gradient on THOSE planes only.  Every finished plane then goes to the ranks whose support contains it (the ``p`` halo
    planes shared with each neighbour), the flag barrier involves those neighbours only, and the reduce writes the planes
    ``support_planes()`` of the gradient (bit-identical to the replicated exchange there) and leaves the others untouched.
    """

    def __init__(self, global_dims: Sequence[SplineDimension], Nout: int, rank: int, world_size: int, group=None,
                 mode: str = "replicated"):
        assert mode in ("replicated", "support")
        self.mode = mode
        import numpy as np
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        last = global_dims[-1]
        self.k0, self.np_ = slab_supports(last, world_size)
        self.rank, self.world = rank, world_size
        self.c_last = last.n_basis_functions
        self.plane_elems = int(np.prod([sd.n_basis_functions for sd in global_dims[:-1]])) if len(global_dims) > 1 else 1
        self.nout = int(Nout)
        self.max_planes = max(self.np_)
        self.dtype, self.device = last.dtype, last.device
        slot = self.nout * self.max_planes * self.plane_elems
        grp = dist.group.WORLD if group is None else group
        self.stage, self.hdl, self.peer_ptrs, self.mc_ptrs = [], [], [], []
        for _ in range(2):
            t = symm_mem.empty(world_size * slot, dtype=self.dtype, device=self.device)
            h = symm_mem.rendezvous(t, grp)
            self.stage.append(t)
            self.hdl.append(h)
            self.peer_ptrs.append((_lib.C.c_void_p * world_size)(*[int(x) for x in h.buffer_ptrs]))
            # NVLS multicast mapping of the same buffers (0 when the fabric / driver has none): one store reaches all ranks
            mc = int(getattr(h, "multicast_ptr", 0) or 0) if os.environ.get("SG_EXCHANGE_MULTICAST", "1") != "0" else 0
            self.mc_ptrs.append(mc)
        # flag array (peer-mapped, one uint64 per rank) and the local exchange counter of the C ABI's own barrier
        self.flags = symm_mem.empty(max(world_size, 16), dtype=torch.int64, device=self.device)
        self.flags.zero_()
        self.flags_hdl = symm_mem.rendezvous(self.flags, grp)
        self.peer_flags = (_lib.C.c_void_p * world_size)(*[int(x) for x in self.flags_hdl.buffer_ptrs])
        self.sync = torch.zeros(8, dtype=torch.int64, device=self.device)          # SG_EXCHANGE_SYNC_BYTES
        torch.cuda.synchronize(self.device)
        dist.barrier(group=grp)                              # every rank's flags are zero before anybody signals
        self._k0s = _lib.i64_array(self.k0)
        self._nps = _lib.i64_array(self.np_)
        self.step = 0
        self._prepared = {}

    def adjoint_and_exchange_(self, grid, *, control_points: torch.Tensor, **kw) -> torch.Tensor:
        """Local ``evaluate_adjoint!`` with the gradient push FUSED into its last kernel
        (``sg_evaluate_adjoint_push``: finished control planes leave for the peers' staging slots while the kernel is still
        running; the local partial gradient is not written at all), then signal + wait + reduce in one kernel.
        Same result as ``evaluate_adjoint_`` + ``exchange_``."""
        from . import _lib
        b = self.step & 1
        self.step += 1
        if self.mode == "support":
            push = ((b, "support"), self.peer_ptrs[b], self.world, self.rank, self.k0[self.rank], self.np_[self.rank],
                    self.max_planes, 0, 0, (self._k0s, self._nps))
        else:
            push = (b, self.peer_ptrs[b], self.world, self.rank, self.k0[self.rank], self.np_[self.rank], self.max_planes, 0,
                    self.mc_ptrs[b])
        evaluate_adjoint_(grid, control_points=control_points, _push=push, **kw)
        return self._wait_reduce_(control_points, b, _lib.stream_ptr(self.device))

    def _wait_reduce_(self, grad: torch.Tensor, b: int, st) -> torch.Tensor:
        from . import _lib
        C = _lib.C
        key = ("wait_reduce", b, grad.data_ptr(), self.mode)
        prep = self._prepared.get(key)
        if prep is None:
            suf = _lib.suffix(self.dtype)
            prep = (getattr(_lib.lib(), "sg_exchange_wait_reduce_" + ("support_" if self.mode == "support" else "") + suf),
                    (_lib.ptr(grad), _lib.ptr(self.stage[b]), _lib.ptr(self.flags), _lib.ptr(self.sync), self.peer_flags,
                     C.c_int(self.world), C.c_int(self.rank), self._k0s, self._nps, C.c_int64(self.plane_elems),
                     C.c_int64(self.c_last), C.c_int(self.nout), C.c_int64(self.max_planes)))
            if len(self._prepared) >= 16:
                self._prepared.clear()
            self._prepared[key] = prep
        _lib.check(prep[0](*prep[1], st), "sg_exchange_wait_reduce")
        return grad

    def support_planes(self, rank: Optional[int] = None) -> Tuple[int, int]:
        """0-based control planes ``[lo, hi)`` of the slowest axis that rank's slab touches (= reads in ``evaluate!``)."""
        r = self.rank if rank is None else rank
        return self.k0[r], self.k0[r] + self.np_[r]

    def allgather_support_planes_(self, control_points: torch.Tensor, group=None) -> torch.Tensor:
        """Rebuild the replicated control-point array after a fit with the support-plane exchange (see the module function)."""
        return allgather_support_planes_(control_points, self.k0, self.np_, group)

    def exchange_(self, grad: torch.Tensor) -> torch.Tensor:
        """Push kernel + (signal, wait, reduce) kernel on an already computed local partial gradient."""
        from . import _lib
        C = _lib.C
        b = self.step & 1
        self.step += 1
        st = _lib.stream_ptr(self.device)
        key = ("push", b, grad.data_ptr())
        prep = self._prepared.get(key)
        if prep is None:                                    # marshal once per (staging buffer, gradient array)
            suf = _lib.suffix(self.dtype)
            prep = (getattr(_lib.lib(), "sg_exchange_push_" + suf),
                    (_lib.ptr(grad), self.peer_ptrs[b], C.c_int(self.world), C.c_int(self.rank), C.c_int64(self.plane_elems),
                     C.c_int64(self.c_last), C.c_int(self.nout), C.c_int64(self.k0[self.rank]), C.c_int64(self.np_[self.rank]),
                     C.c_int64(self.max_planes)))
            if len(self._prepared) >= 16:
                self._prepared.clear()
            self._prepared[key] = prep
        _lib.check(prep[0](*prep[1], st), "sg_exchange_push")
        return self._wait_reduce_(grad, b, st)

    def status(self):
        """(exchanges completed on this rank, whether a wait ever timed out) -- blocking."""
        from . import _lib
        C = _lib.C
        ep, to = C.c_ulonglong(0), C.c_int(0)
        _lib.check(_lib.lib().sg_exchange_status(_lib.ptr(self.sync), C.byref(ep), C.byref(to), _lib.stream_ptr(self.device)),
                   "sg_exchange_status")
        return int(ep.value), bool(to.value)


class SlabShardedGrid:
    """This rank's slab of a global spline grid.

    ``global_dims``: the full per-dimension objects (cheap: 1-D tables).  ``local`` is a ``SplineGrid``
    whose last dimension holds rows ``[lo, hi)`` only; its ``eval`` is the local ``(n_1..n_{D-1}, hi-lo, Nout)``
    slab and its control points (and weights) are full replicas.
    """

    def __init__(self, global_dims: Sequence[SplineDimension], Nout: int, rank: int, world_size: int,
                 nurbs: bool = False, group=None, peer_exchange: bool = False, exchange_mode: str = "replicated"):
        self.global_dims = tuple(global_dims)
        self.rank, self.world_size, self.group = rank, world_size, group
        self.exchange = None
        self.fused_push = True        # push fused into the adjoint's last kernel (False: separate push kernel)
        self.exchange_kind = "none" if world_size == 1 else "nccl_allreduce"
        n_last = self.global_dims[-1].n_sample_points
        assert n_last >= world_size, "fewer sample rows along the slowest axis than ranks"
        self.lo, self.hi = slab_bounds(n_last, world_size, rank)
        dims = self.global_dims[:-1] + (slice_dimension(self.global_dims[-1], self.lo, self.hi),)
        self.local: SplineGrid = NURBSGrid(dims, Nout) if nurbs else SplineGrid(dims, Nout)
        if peer_exchange and world_size > 1:
            try:
                self.exchange = PeerGradientExchange(self.global_dims, Nout, rank, world_size, group, mode=exchange_mode)
                self.exchange_kind = ("peer_memory_push + device-side flag barrier + reduce (push fused into the adjoint's last "
                                      "kernel; signal, wait and reduce are one kernel)")
                if exchange_mode == "support":
                    self.exchange_kind += ("; support-plane exchange: each rank ends with the summed gradient on the control "
                                           "planes its slab reads, halo planes go to the neighbouring ranks only")
            except Exception as e:   # symmetric memory unavailable: keep the NCCL all-reduce
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable ({e!r}); using the NCCL all-reduce")

    @property
    def global_sample_grid_size(self) -> Tuple[int, ...]:
        return tuple(sd.n_sample_points for sd in self.global_dims)

    def evaluate_(self, **kw) -> None:
        """Local ``evaluate!``: no communication."""
        evaluate_(self.local, **kw)

    def evaluate_adjoint_(self, *, control_points: Optional[torch.Tensor] = None, **kw) -> None:
        """Local ``evaluate_adjoint!`` followed by the gradient all-reduce; afterwards every rank holds the
        full gradient, exactly what the single-device call produces (up to summation order)."""
        cp = obtain(self.local.control_points if control_points is None else control_points)
        if self.exchange is not None and self.fused_push:
            self.exchange.adjoint_and_exchange_(self.local, control_points=cp, **kw)
            return
        evaluate_adjoint_(self.local, control_points=control_points, **kw)
        if self.exchange is not None:
            self.exchange.exchange_(cp)
        else:
            allreduce_gradient_(cp, self.group)
