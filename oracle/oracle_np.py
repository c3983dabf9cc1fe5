"""CPU oracle (numpy) for the SplineGrids.jl grid-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  This module restates, on the CPU, the algorithms of the
reference's KernelAbstractions kernels so the CUDA path can be checked against them.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``splinegrids.jl_b200``) never does: it fails loudly when its CUDA library is missing.

Parity pinning: the reference is pure Julia and Julia is not installed in the build
image, so the reference itself cannot be executed here.  This restatement is pinned
against every golden vector / known answer the reference's own tests hold for the path
(see ``tests/test_oracle_golden.py``).  The NURBS adjoint
has no reference behaviour at all ("parity unpinned" for that single function, see
``evaluate_adjoint`` below).

Conventions
-----------
* Arrays use the *Julia shapes*: ``eval`` is ``(n_1..n_D, Nout)``, control points are
  ``(c_1..c_D, Nout)``, basis tables are ``(n, p+1, mdo+1)``.  All arrays are numpy arrays in
  Fortran order, so the memory layout equals Julia's column-major layout byte for byte.
* Index arrays (``sample_indices``, ``row_pointer``, ``column_start``,
  ``refinement_indices``) hold **1-based** values exactly as the reference stores them.
* All citations are ``path:line`` relative to the reference checkout.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from fractions import Fraction
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# L1: knot vectors and sample points (host-side set-up in the reference)
# --------------------------------------------------------------------------------------


def linspace_exact(a: float, b: float, n: int, dtype) -> np.ndarray:
    """Correctly rounded ``a + i*(b-a)/(n-1)``.

    Stands in for Julia's ``range(a, b; length=n)`` (twice-precision stepping) followed by
    ``float_type.(...)`` -- src/knot_vector.jl:95-101, src/spline_dimension.jl:121-128.
    Exact rational arithmetic, rounded once to float64 and then to ``dtype``.
    """
    dtype = np.dtype(dtype)
    if n == 1:
        return np.array([a], dtype=dtype)
    fa, fb = Fraction(float(a)), Fraction(float(b))
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        out[i] = float(fa + (fb - fa) * Fraction(i, n - 1))
    return out.astype(dtype)


def expand_knot_vector(knot_values: np.ndarray, multiplicities: np.ndarray) -> np.ndarray:
    """K9 ``expand_knot_vector_kernel`` -- src/util_kernels.jl:1-20."""
    knots_all = np.empty(int(np.sum(multiplicities)), dtype=knot_values.dtype)
    for i in range(len(knot_values)):
        mult_sum = 0
        for j in range(i):
            mult_sum += int(multiplicities[j])
        for k in range(mult_sum, mult_sum + int(multiplicities[i])):
            knots_all[k] = knot_values[i]
    return knots_all


def clamped_knot_vector(n_basis: int, degree: int, dtype=np.float32, extent=(0, 1),
                        distribution: str = "equispaced", rng=None):
    """``KnotVector(n_basis_functions, degree; ...)`` -- src/knot_vector.jl:75-113.

    Returns ``(knot_values, multiplicities[int32], knots_all)``.
    """
    assert n_basis - degree >= 1                     # src/knot_vector.jl:84
    n_knot_values = n_basis - degree + 1
    dtype = np.dtype(dtype)
    if distribution == "random":                     # src/knot_vector.jl:87-92
        rng = np.random.default_rng(0) if rng is None else rng
        kv = np.cumsum(rng.random(n_knot_values).astype(dtype)).astype(dtype)
        kv = kv - kv[0]
        kv = kv / (kv[-1] - kv[0])
        kv = kv * dtype.type(extent[1] - extent[0])
        kv = (kv + dtype.type(extent[0])).astype(dtype)
    elif distribution == "equispaced":               # src/knot_vector.jl:94-101
        kv = linspace_exact(extent[0], extent[1], n_knot_values, dtype)
    else:
        raise ValueError(f"Unsupported knot distribution type {distribution}.")
    mult = np.ones(n_knot_values, dtype=np.int32)    # src/knot_vector.jl:105-108
    mult[0] = degree + 1
    mult[-1] = degree + 1
    return kv, mult, expand_knot_vector(kv, mult)


def default_sample_points(knot_values: np.ndarray, n: int) -> np.ndarray:
    """Equispaced samples over the knot extent -- src/spline_dimension.jl:121-128."""
    return linspace_exact(float(knot_values[0]), float(knot_values[-1]), n, knot_values.dtype)


# --------------------------------------------------------------------------------------
# L2: per-dimension basis tables (K1, K2, K10)
# --------------------------------------------------------------------------------------


def span_indices(sample_points: np.ndarray, knots_all: np.ndarray, degree: int) -> np.ndarray:
    """K1 ``set_sample_indices_kernel`` -- src/util_kernels.jl:22-49 (literal scan).

    ``idx = #{leading knots with not (t < knot)}``, clamped to ``[p+1, n_knots-p-1]``.
    """
    n_knots = len(knots_all)
    out = np.empty(len(sample_points), dtype=np.int32)
    for i, t in enumerate(sample_points):
        idx = 0
        for knot in knots_all:
            if t < knot:
                break
            idx += 1
        out[i] = min(max(idx, degree + 1), n_knots - degree - 1)
    return out


def basis_tables(knots_all: np.ndarray, sample_points: np.ndarray, sample_indices: np.ndarray,
                 degree: int, max_derivative_order: int = 0) -> np.ndarray:
    """K2 ``spline_dimension_kernel`` -- src/spline_dimension.jl:160-215.

    Cox-de Boor triangle with derivative rows, literal to the reference (same operation
    order), vectorised over the samples only.  Returns ``eval[n, p+1, mdo+1]`` (F order).
    """
    T = knots_all.dtype.type
    n = len(sample_points)
    p, mdo = degree, max_derivative_order
    assert 0 <= mdo <= p                                      # src/spline_dimension.jl:108
    t = sample_points
    i = sample_indices.astype(np.int64)                        # 1-based span index
    cur = np.zeros((n, p + 1, mdo + 1), dtype=knots_all.dtype, order="F")
    prev = np.zeros_like(cur)
    cur[:, 0, 0] = T(1)                                        # :179-180
    prev[:, 0, 0] = T(1)
    for k in range(1, p + 1):                                  # :183
        cur[:] = T(0)                                          # :185-187
        for k_ in range(1, k + 1):                             # :189
            t_min = knots_all[i + k_ - k - 1]                  # knots_all[i + k_ - k] (1-based)
            t_max = knots_all[i + k_ - 1]                      # knots_all[i + k_]
            dt = t_max - t_min
            b_prev = prev[:, k_ - 1, 0]
            frac = b_prev / dt
            cur[:, k_ - 1, 0] += frac * (t_max - t)            # :196
            cur[:, k_, 0] = frac * (t - t_min)                 # :197
            for d in range(1, mdo + k - p + 1):                # :200
                b_prev = prev[:, k_ - 1, d - 1]
                contrib = b_prev * T(k) / dt                   # :202
                cur[:, k_ - 1, d] -= contrib
                cur[:, k_, d] = contrib
        if k != p:                                             # :209-213
            prev[:] = cur
    return cur


def decompress(table: np.ndarray, sample_indices: np.ndarray, degree: int, n_basis: int,
               derivative_order: int = 0) -> np.ndarray:
    """K10 ``decompress_basis_function_eval_kernel`` -- src/util_kernels.jl:51-67."""
    n = table.shape[0]
    out = np.zeros((n, n_basis), dtype=table.dtype, order="F")
    for i in range(n):
        l = int(sample_indices[i])
        for j, k in enumerate(range(l - degree, l + 1)):
            out[i, k - 1] = table[i, j, derivative_order]
    return out


# --------------------------------------------------------------------------------------
# L4: grid evaluation (K3) and adjoint (K4)
# --------------------------------------------------------------------------------------


def _window_offsets(degrees: Sequence[int]):
    """CartesianIndices(control_point_kernel_size): first dimension fastest."""
    rev = [range(p + 1) for p in reversed(degrees)]
    for I in itertools.product(*rev):
        yield tuple(reversed(I))


def evaluate(tables: Sequence[np.ndarray], sample_indices: Sequence[np.ndarray],
             degrees: Sequence[int], derivative_order: Sequence[int],
             control_points: np.ndarray, weights: Optional[np.ndarray] = None) -> np.ndarray:
    """K3 ``spline_eval_kernel`` -- src/spline_grid.jl:119-183.

    Literal in accumulation order (window offsets column-major, basis product
    ``((1*B1)*B2)...``, NURBS weight applied after the product, divide at the end),
    vectorised over the sample grid only.
    """
    nin = len(tables)
    nout = control_points.shape[-1]
    dt = control_points.dtype
    n = tuple(len(ix) for ix in sample_indices)
    out = np.zeros(n + (nout,), dtype=dt, order="F")           # :136-138
    denom = np.zeros(n, dtype=dt)
    base = [sample_indices[d].astype(np.int64) - degrees[d] - 1 for d in range(nin)]  # :143-144 (0-based cp idx of I=1)
    for I in _window_offsets(degrees):                         # :146
        prod = np.ones(n, dtype=dt)
        for d in range(nin):                                   # :153-156
            b = tables[d][:, I[d], derivative_order[d]]
            shape = [1] * nin
            shape[d] = n[d]
            prod = prod * b.reshape(shape)
        idx = np.ix_(*[base[d] + I[d] for d in range(nin)])
        if weights is not None:                                # :159-166
            prod = prod * weights[idx]
            denom = denom + prod
        for o in range(nout):                                  # :171-174
            out[..., o] += prod * control_points[idx + (o,)]
    if weights is not None:                                    # :178-182
        for o in range(nout):
            out[..., o] /= denom
    return out


def evaluate_loops(tables, sample_indices, degrees, derivative_order, control_points, weights=None):
    """K3 as pure Python loops, one "thread" per sample (tiny cases; cross-checks `evaluate`)."""
    nin = len(tables)
    nout = control_points.shape[-1]
    T = control_points.dtype.type
    n = tuple(len(ix) for ix in sample_indices)
    out = np.zeros(n + (nout,), dtype=control_points.dtype, order="F")
    for J in itertools.product(*[range(m) for m in n]):
        denom = T(0)
        base = [int(sample_indices[d][J[d]]) - degrees[d] - 1 for d in range(nin)]
        for I in _window_offsets(degrees):
            cpi = tuple(base[d] + I[d] for d in range(nin))
            prod = T(1)
            for d in range(nin):
                prod = T(prod * tables[d][J[d], I[d], derivative_order[d]])
            if weights is not None:
                prod = T(prod * weights[cpi])
                denom = T(denom + prod)
            for o in range(nout):
                out[J + (o,)] = T(out[J + (o,)] + T(prod * control_points[cpi + (o,)]))
        if weights is not None:
            for o in range(nout):
                out[J + (o,)] = T(out[J + (o,)] / denom)
    return out


def evaluate_adjoint(tables, sample_indices, degrees, derivative_order, eval_in: np.ndarray,
                     cp_shape: Sequence[int], weights: Optional[np.ndarray] = None) -> np.ndarray:
    """K4 ``spline_eval_adjoint_kernel`` + zero fill -- src/adjoint.jl:1-40, :61.

    ``cp[base(J)+I, o] += b * eval[J, o]``.  The reference uses float atomics, i.e. an
    unspecified summation order; here the order is window-offset-major with numpy's
    unbuffered ``add.at`` inside one offset.

    ``weights`` (NURBS) is an EXTENSION with no reference behaviour (the reference method is
    restricted to non-rational grids, src/adjoint.jl:52,57): PARITY UNPINNED.  It is the
    transpose of the fixed-weights rational map: b -> b*w[base+I]/denom(J).
    """
    nin = len(tables)
    nout = eval_in.shape[-1]
    dt = eval_in.dtype
    n = eval_in.shape[:-1]
    cp = np.zeros(tuple(cp_shape), dtype=dt, order="F")        # :61
    base = [sample_indices[d].astype(np.int64) - degrees[d] - 1 for d in range(nin)]

    def products(I):
        prod = np.ones(n, dtype=dt)
        for d in range(nin):
            shape = [1] * nin
            shape[d] = n[d]
            prod = prod * tables[d][:, I[d], derivative_order[d]].reshape(shape)
        return prod

    denom = None
    if weights is not None:
        denom = np.zeros(n, dtype=dt)
        for I in _window_offsets(degrees):
            idx = np.ix_(*[base[d] + I[d] for d in range(nin)])
            denom = denom + products(I) * weights[idx]
    for I in _window_offsets(degrees):                         # :21
        prod = products(I)
        idx = np.ix_(*[base[d] + I[d] for d in range(nin)])
        if weights is not None:
            prod = prod * weights[idx] / denom
        for o in range(nout):                                  # :33-38
            np.add.at(cp[..., o], idx, prod * eval_in[..., o])
    return cp


# --------------------------------------------------------------------------------------
# L3b: refinement matrices (storage A.5, apply K5, adjoint K6) and their construction
# --------------------------------------------------------------------------------------


@dataclass
class RefMat:
    """``RefinementMatrix`` -- src/refinement_matrix.jl:19-57.  1-based index arrays."""
    m: int
    n: int
    row_pointer: np.ndarray
    column_start: np.ndarray
    nzval: np.ndarray

    def column_range(self, i: int) -> Tuple[int, int]:
        """``get_column_range`` for 1-based row i -- src/refinement_matrix.jl:103-125."""
        rp_next = len(self.nzval) + 1 if i == self.m else int(self.row_pointer[i])
        rp = int(self.row_pointer[i - 1])
        cs = int(self.column_start[i - 1])
        return cs, cs + (rp_next - rp) - 1

    def validate(self) -> List[int]:
        """``validate_refinement_matrix_kernel`` -- src/refinement_matrix.jl:134-181.
        Returns the list of invalid (1-based) rows."""
        bad = []
        for i in range(1, self.m + 1):
            cs, ce = self.column_range(i)
            ok = ce >= cs
            if ok:
                ok = cs >= 1 and ce <= self.n
            first = i == 1
            if first:
                ok = cs == 1 and int(self.row_pointer[0]) == 1
            if ok and not first:
                csp, cep = self.column_range(i - 1)
                ok = csp <= cs <= cep + 1
                if ok:
                    ok = ce >= cep
            if not ok:
                bad.append(i)
        return bad

    def dense(self) -> np.ndarray:
        """``collect`` -- src/refinement_matrix.jl:329-363."""
        out = np.zeros((self.m, self.n), dtype=self.nzval.dtype)
        for i in range(1, self.m + 1):
            cs, ce = self.column_range(i)
            rp = int(self.row_pointer[i - 1])
            out[i - 1, cs - 1:ce] = self.nzval[rp - 1:rp - 1 + (ce - cs + 1)]
        return out


def refmat_from_dense(A: np.ndarray, int_dtype=np.int32) -> RefMat:
    """``RefinementMatrix(A::Matrix)`` -- src/refinement_matrix.jl:476-504."""
    rp, cs, nz = [], [], []
    pointer = 1
    for row in A:
        nzi = np.flatnonzero(row)
        rp.append(pointer)
        if len(nzi) == 0:
            cs.append(0)
        else:
            cs.append(int(nzi[0]) + 1)
            nz.extend(row[nzi[0]:nzi[-1] + 1].tolist())
            pointer += int(nzi[-1] - nzi[0]) + 1
    return RefMat(A.shape[0], A.shape[1], np.array(rp, dtype=int_dtype),
                  np.array(cs, dtype=int_dtype), np.array(nz, dtype=A.dtype))


def rmeye(n: int, dtype=np.float32, int_dtype=np.int32) -> RefMat:
    """``rmeye`` -- src/refinement_matrix.jl:463-474."""
    r = np.arange(1, n + 1, dtype=int_dtype)
    return RefMat(n, n, r.copy(), r.copy(), np.ones(n, dtype=dtype))


def refmat_matmul(A: RefMat, B: RefMat) -> RefMat:
    """``A * B`` -- src/refinement_matrix.jl:184-327 (structural product: the support of row i
    of C is the union of the supports of the rows of B hit by row i of A)."""
    assert A.n == B.m
    n_nonzero = np.zeros(A.m, dtype=A.row_pointer.dtype)
    cs_C = np.zeros(A.m, dtype=A.row_pointer.dtype)
    for i in range(1, A.m + 1):                      # :232-271
        a0, a1 = A.column_range(i)
        cnt, first = 0, 0
        for j in range(B.n, 0, -1):
            for k in range(a0, a1 + 1):
                b0, b1 = B.column_range(k)
                if b0 <= j <= b1:
                    cnt += 1
                    first = j
                    break
        n_nonzero[i - 1] = cnt
        cs_C[i - 1] = first
    rp_C = np.zeros(A.m, dtype=A.row_pointer.dtype)  # :300-303
    rp_C[1:] = np.cumsum(n_nonzero[:-1])
    rp_C += 1
    nz_C = np.zeros(int(n_nonzero.sum()), dtype=A.nzval.dtype)
    C = RefMat(A.m, B.n, rp_C, cs_C, nz_C)
    for i in range(1, A.m + 1):                      # :184-228
        a0, a1 = A.column_range(i)
        c0, c1 = C.column_range(i)
        pa = int(A.row_pointer[i - 1])
        for k in range(a0, a1 + 1):
            b0, b1 = B.column_range(k)
            pc = int(rp_C[i - 1])
            for j in range(c0, c1 + 1):
                if b0 <= j <= b1:
                    pb = int(B.row_pointer[k - 1]) + j - b0
                    nz_C[pc - 1] += A.nzval[pa - 1] * B.nzval[pb - 1]
                pc += 1
            pa += 1
    return C


def _row_extends(I, refmat_index_all, As):
    """``get_row_extends`` -- src/utils.jl:204-235 (I is a 1-based index tuple)."""
    cs, nc = [], []
    for dim, r in enumerate(refmat_index_all):
        if r == 0:
            cs.append(I[dim])
            nc.append(1)
        else:
            c0, c1 = As[r - 1].column_range(I[dim])
            cs.append(c0)
            nc.append(c1 - c0 + 1)
    return cs, nc


def _refmat_index_all(ndims: int, dims_refinement: Sequence[int]):
    """src/refinement_matrix.jl:431-432 (dims_refinement is 1-based)."""
    dims_refinement = list(dims_refinement)
    return [dims_refinement.index(d) + 1 if d in dims_refinement else 0 for d in range(1, ndims + 1)]


def validate_mult_input(Y, As, B, dims_refinement):
    """src/validation.jl:90-107."""
    assert len(set(dims_refinement)) == len(dims_refinement), "Refinement dimensions must be unique."
    assert len(As) == len(dims_refinement), \
        "There must be exactly one refinement dimension per refinement matrix."
    for dim, (sb, sy) in enumerate(zip(B.shape, Y.shape), start=1):
        if dim in dims_refinement:
            A = As[list(dims_refinement).index(dim)]
            if not (A.m == sy and A.n == sb):
                raise ValueError(
                    f"Size of refinement matrix does not match `B` and `Y` along refinement dimension {dim}.")
        elif sb != sy:
            raise ValueError(f"`B` and `Y` don't have the same size along dimension {dim}.")


def mult(Y: np.ndarray, As: Sequence[RefMat], B: np.ndarray, dims_refinement: Sequence[int]) -> None:
    """K5 ``refinement_matrix_array_mul_kernel`` / ``mult!`` -- src/refinement_matrix.jl:365-445.
    Literal loops (one "thread" per element of Y, all dims)."""
    validate_mult_input(Y, As, B, dims_refinement)
    ridx = _refmat_index_all(Y.ndim, dims_refinement)
    T = Y.dtype.type
    for I0 in np.ndindex(*Y.shape):
        I = tuple(i + 1 for i in I0)
        cs, nc = _row_extends(I, ridx, As)
        out = T(0)
        for Jb in itertools.product(*[range(1, c + 1) for c in reversed(nc)]):
            Jb = tuple(reversed(Jb))
            J = tuple(Jb[d] + cs[d] - 1 for d in range(Y.ndim))
            contrib = B[tuple(j - 1 for j in J)]
            for d in range(Y.ndim):
                r = ridx[d]
                if r != 0:
                    rp = int(As[r - 1].row_pointer[I[d] - 1])
                    contrib = T(contrib * As[r - 1].nzval[rp + Jb[d] - 2])
            out = T(out + contrib)
        Y[I0] = out


def mult_adjoint(B: np.ndarray, As: Sequence[RefMat], Y: np.ndarray, dims_refinement: Sequence[int]) -> None:
    """K6 ``refinement_matrix_array_mul_adjoint_kernel`` / ``mult_adjoint!`` -- src/adjoint.jl:85-152
    (float path; atomics replaced by a sequential scatter)."""
    validate_mult_input(Y, As, B, dims_refinement)
    B[...] = 0                                       # :135
    ridx = _refmat_index_all(Y.ndim, dims_refinement)
    T = Y.dtype.type
    for I0 in np.ndindex(*Y.shape):
        I = tuple(i + 1 for i in I0)
        cs, nc = _row_extends(I, ridx, As)
        for Jb in itertools.product(*[range(1, c + 1) for c in reversed(nc)]):
            Jb = tuple(reversed(Jb))
            J0 = tuple(Jb[d] + cs[d] - 2 for d in range(Y.ndim))
            contrib = Y[I0]
            for d in range(Y.ndim):
                r = ridx[d]
                if r != 0:
                    rp = int(As[r - 1].row_pointer[I[d] - 1])
                    contrib = T(contrib * As[r - 1].nzval[rp + Jb[d] - 2])
            B[J0] = T(B[J0] + contrib)


def mult_dense(As: Sequence[RefMat], B: np.ndarray, dims_refinement: Sequence[int]) -> np.ndarray:
    """Fast equivalent of `mult` through dense mode-n products (for mid-size cross-checks)."""
    Y = B
    for A, d in zip(As, dims_refinement):
        Y = np.moveaxis(np.tensordot(A.dense(), Y, axes=(1, d - 1)), 0, d - 1)
    return np.asfortranarray(Y)


def mult_adjoint_dense(As: Sequence[RefMat], Y: np.ndarray, dims_refinement: Sequence[int]) -> np.ndarray:
    B = Y
    for A, d in zip(As, dims_refinement):
        B = np.moveaxis(np.tensordot(A.dense().T, B, axes=(1, d - 1)), 0, d - 1)
    return np.asfortranarray(B)


# --- construction of refinement matrices (set-up path; needed to BUILD config 5) ---------


def boehm_matrix(knots_all: np.ndarray, degree: int, knot_span_index: int, knot_new) -> RefMat:
    """K13 ``build_refinement_matrix_kernel`` + ctor -- src/refinement.jl:3-36, 53-88."""
    T = knots_all.dtype.type
    n = len(knots_all) - degree - 1
    m = n + 1
    rp = np.empty(m, dtype=np.int32)
    cs = np.empty(m, dtype=np.int32)
    nz = np.ones(n + degree + 1, dtype=knots_all.dtype)
    k, p = knot_span_index, degree
    for i in range(1, m + 1):
        if i <= k - p:
            rp[i - 1] = i
            cs[i - 1] = i
        elif i <= k:
            alpha = (T(knot_new) - knots_all[i - 1]) / (knots_all[i + p - 1] - knots_all[i - 1])
            r = 2 * i - k + p - 1
            rp[i - 1] = r
            cs[i - 1] = i - 1
            nz[r - 1] = T(1) - alpha
            nz[r] = alpha
        else:
            rp[i - 1] = i + p
            cs[i - 1] = i - 1
    return RefMat(m, n, rp, cs, nz)


def insert_knot(knot_values: np.ndarray, multiplicities: np.ndarray, knot_new):
    """``insert_knot(::KnotVector, knot_new)`` -- src/refinement.jl:107-133.
    Returns ``(knot_values_new, multiplicities_new, knot_span_index)``."""
    T = knot_values.dtype.type
    knot_new = T(knot_new)
    assert not np.any(knot_values == knot_new)
    pos = int(np.searchsorted(knot_values, knot_new, side="left"))     # searchsortedfirst - 1
    kv = np.insert(knot_values, pos, knot_new)                           # K11
    mu = np.insert(multiplicities, pos, 1)
    return kv, mu, int(np.sum(multiplicities[:pos]))


def refine_dimension(knot_values: np.ndarray, multiplicities: np.ndarray, degree: int,
                     knots_new: Optional[np.ndarray] = None):
    """``refine(::SplineDimension)`` -- src/refinement.jl:256-284.
    Returns ``(knot_values_new, multiplicities_new, R)`` with R the accumulated matrix."""
    if knots_new is None:                                                # :264-266
        knots_new = knot_values[:-1] + np.diff(knot_values) / knot_values.dtype.type(2)
    n_basis = int(np.sum(multiplicities)) - degree - 1
    R = rmeye(n_basis, dtype=knot_values.dtype)
    kv, mu = knot_values, multiplicities
    for knot in knots_new:                                               # :275-279
        knots_all_old = expand_knot_vector(kv, mu)
        kv, mu, k = insert_knot(kv, mu, knot)
        R = refmat_matmul(boehm_matrix(knots_all_old, degree, k, knot), R)
    return kv, mu, R


# --------------------------------------------------------------------------------------
# L3: locally refined control points (K7, K8) and THB bookkeeping
# --------------------------------------------------------------------------------------


@dataclass
class LocalRefinementO:
    """``LocalRefinement`` -- src/control_points.jl:70-110."""
    dims_refinement: List[int]
    refinement_matrices: List[RefMat]
    refinement_indices: np.ndarray          # (n_active, Nin) int32, 1-based
    refinement_values: np.ndarray           # (n_active, Nout)


@dataclass
class LocallyRefinedCP:
    """``LocallyRefinedControlPoints`` -- src/control_points.jl:173-199."""
    control_points_refined: List[np.ndarray]
    local_refinements: List[LocalRefinementO]

    def n_control_points(self) -> int:       # src/control_points.jl:281-288
        return sum(lr.refinement_indices.shape[0] for lr in self.local_refinements)

    def get_values(self) -> np.ndarray:      # vcat view, :201-233
        return np.concatenate([lr.refinement_values for lr in self.local_refinements], axis=0)

    def set_values(self, M: np.ndarray) -> None:
        off = 0
        for lr in self.local_refinements:
            k = lr.refinement_values.shape[0]
            lr.refinement_values[...] = M[off:off + k]
            off += k


def scatter_active(cp: np.ndarray, refinement_indices: np.ndarray, refinement_values: np.ndarray) -> None:
    """K7 ``local_refinement_kernel`` -- src/control_points.jl:296-311."""
    nin = cp.ndim - 1
    for i in range(refinement_indices.shape[0]):
        idx = tuple(int(refinement_indices[i, d]) - 1 for d in range(nin))
        for o in range(cp.shape[-1]):
            cp[idx + (o,)] = refinement_values[i, o]


def gather_zero_active(refinement_values: np.ndarray, cp: np.ndarray, refinement_indices: np.ndarray) -> None:
    """K8 ``local_refinement_adjoint_kernel`` -- src/adjoint.jl:154-170."""
    nin = cp.ndim - 1
    for i in range(refinement_indices.shape[0]):
        idx = tuple(int(refinement_indices[i, d]) - 1 for d in range(nin))
        for o in range(cp.shape[-1]):
            refinement_values[i, o] = cp[idx + (o,)]
            cp[idx + (o,)] = 0


def lrcp_evaluate(lrcp: LocallyRefinedCP, fast: bool = True) -> None:
    """``evaluate!(::LocallyRefinedControlPoints)`` -- src/control_points.jl:319-349."""
    for i, lr in enumerate(lrcp.local_refinements):
        cp_new = lrcp.control_points_refined[i]
        if i > 0:
            cp_prev = lrcp.control_points_refined[i - 1]
            if fast:
                cp_new[...] = mult_dense(lr.refinement_matrices, cp_prev, lr.dims_refinement)
            else:
                mult(cp_new, lr.refinement_matrices, cp_prev, lr.dims_refinement)
        if lr.refinement_indices.shape[0] > 0:
            scatter_active(cp_new, lr.refinement_indices, lr.refinement_values)


def lrcp_evaluate_adjoint(lrcp: LocallyRefinedCP, fast: bool = True) -> None:
    """``evaluate_adjoint!(::LocallyRefinedControlPoints)`` -- src/adjoint.jl:174-205."""
    for i in reversed(range(len(lrcp.local_refinements))):
        lr = lrcp.local_refinements[i]
        cp_new = lrcp.control_points_refined[i]
        if lr.refinement_indices.shape[0] > 0:
            gather_zero_active(lr.refinement_values, cp_new, lr.refinement_indices)
        if i > 0:
            cp_prev = lrcp.control_points_refined[i - 1]
            if fast:
                cp_prev[...] = mult_adjoint_dense(lr.refinement_matrices, cp_new, lr.dims_refinement)
            else:
                mult_adjoint(cp_prev, lr.refinement_matrices, cp_new, lr.dims_refinement)


@dataclass
class DimO:
    """Plain-data stand-in for ``SplineDimension`` -- src/spline_dimension.jl:25-66."""
    degree: int
    max_derivative_order: int
    knot_values: np.ndarray
    multiplicities: np.ndarray
    knots_all: np.ndarray
    sample_points: np.ndarray
    sample_indices: np.ndarray = field(default=None)
    eval: np.ndarray = field(default=None)

    @property
    def n_basis(self) -> int:
        return len(self.knots_all) - self.degree - 1

    def rebuild(self) -> None:
        self.sample_indices = span_indices(self.sample_points, self.knots_all, self.degree)
        self.eval = basis_tables(self.knots_all, self.sample_points, self.sample_indices,
                                 self.degree, self.max_derivative_order)


def make_dimension(n_basis: int, degree: int, n_samples: int, dtype=np.float32,
                   max_derivative_order: int = 0, extent=(0, 1), distribution="equispaced",
                   rng=None, knot_values=None, multiplicities=None) -> DimO:
    """``SplineDimension(n_basis, degree, n_samples; ...)`` -- src/spline_dimension.jl:97-156."""
    if knot_values is None:
        kv, mu, ka = clamped_knot_vector(n_basis, degree, dtype, extent, distribution, rng)
    else:
        kv, mu = np.asarray(knot_values), np.asarray(multiplicities, dtype=np.int32)
        ka = expand_knot_vector(kv, mu)
        assert len(ka) == n_basis + degree + 1, "Incompatible knot vector supplied."
    d = DimO(degree, max_derivative_order, kv, mu, ka, default_sample_points(kv, n_samples))
    d.rebuild()
    return d


def unit_cp_grid(cp_shape: Sequence[int], dtype) -> np.ndarray:
    """``set_unit_cp_grid!`` -- src/utils.jl:1-15."""
    cp = np.zeros(tuple(cp_shape), dtype=dtype, order="F")
    nin = len(cp_shape) - 1
    for nn in range(min(nin, cp_shape[-1])):
        coords = np.linspace(0, 1, cp_shape[nn])
        shape = [1] * nin
        shape[nn] = cp_shape[nn]
        cp[..., nn] = coords.reshape(shape)
    return cp


def add_default_local_refinement(dims: List[DimO], control_points):
    """``add_default_local_refinement`` -- src/control_points.jl:364-425.
    ``control_points`` is a dense array (DefaultControlPoints) or a `LocallyRefinedCP`.
    Returns ``(dims_new, LocallyRefinedCP)``."""
    nin = len(dims)
    dims_new, mats = [], []
    for d in dims:
        kv, mu, R = refine_dimension(d.knot_values, d.multiplicities, d.degree)
        dn = DimO(d.degree, d.max_derivative_order, kv, mu, expand_knot_vector(kv, mu), d.sample_points)
        dn.rebuild()
        dims_new.append(dn)
        mats.append(R)
    cp_fine_old = control_points.control_points_refined[-1] if isinstance(control_points, LocallyRefinedCP) \
        else control_points
    nout = cp_fine_old.shape[-1]
    dr = list(range(1, nin + 1))
    cp_new = mult_dense(mats, cp_fine_old, dr).astype(cp_fine_old.dtype)
    lr = LocalRefinementO(dr, mats, np.zeros((0, nin), dtype=np.int32),
                          np.zeros((0, nout), dtype=cp_fine_old.dtype))
    if isinstance(control_points, LocallyRefinedCP):
        control_points.control_points_refined.append(cp_new)
        control_points.local_refinements.append(lr)
        return dims_new, control_points
    # base level: every control point active, values = copy (:126-154)
    shape = cp_fine_old.shape[:-1]
    n_cp = int(np.prod(shape))
    idx = np.stack(np.unravel_index(np.arange(n_cp), shape, order="F"), axis=1).astype(np.int32) + 1
    vals = cp_fine_old.reshape((n_cp, nout), order="F").copy()
    base = LocalRefinementO([], [], idx, vals)
    return dims_new, LocallyRefinedCP([cp_fine_old, cp_new], [base, lr])


def _unique_rows_first(a: np.ndarray) -> np.ndarray:
    """Julia ``unique(a, dims=1)``: first occurrences, original order."""
    seen, keep = set(), []
    for r in range(a.shape[0]):
        key = tuple(int(v) for v in a[r])
        if key not in seen:
            seen.add(key)
            keep.append(r)
    return a[keep]


def activate_local_refinement(lrcp: LocallyRefinedCP, refinement_indices: np.ndarray,
                              refinement_index: Optional[int] = None) -> None:
    """``activate_local_refinement!`` + K18 -- src/control_points.jl:427-516."""
    level = len(lrcp.local_refinements) - 1 if refinement_index is None else refinement_index - 1
    lr = lrcp.local_refinements[level]
    cp = lrcp.control_points_refined[level]
    nin = cp.ndim - 1
    assert refinement_indices.shape[1] == nin
    idx_new = _unique_rows_first(np.concatenate([lr.refinement_indices, refinement_indices.astype(np.int32)], axis=0))
    n_old = lr.refinement_values.shape[0]
    vals = np.empty((idx_new.shape[0], cp.shape[-1]), dtype=cp.dtype)
    vals[:n_old] = lr.refinement_values
    for i in range(n_old, idx_new.shape[0]):
        vals[i] = cp[tuple(int(v) - 1 for v in idx_new[i])]
    lrcp.local_refinements[level] = LocalRefinementO(lr.dims_refinement, lr.refinement_matrices, idx_new, vals)


def activate_local_control_point_range(lrcp: LocallyRefinedCP, *ranges: Tuple[int, int]) -> None:
    """``activate_local_control_point_range!`` -- src/control_points.jl:531-539.
    Each range is a 1-based inclusive ``(lo, hi)``.  Row order: the reference builds the Nin-dimensional array of
    index vectors with ``Iterators.product`` (dim 1 fastest) and flattens its ADJOINT, ``reduce(vcat, A')`` (:537),
    i.e. for Nin = 2 the rows come out with dimension 2 fastest; for Nin = 1 the order is unchanged; for Nin >= 3 the
    adjoint of an N-d array does not exist and the reference call errors (kept here as dim 1 fastest)."""
    rs = [range(lo, hi + 1) for lo, hi in ranges]
    if len(rs) == 2:
        rows = [(i, j) for i in rs[0] for j in rs[1]]
    else:
        rows = [tuple(reversed(t)) for t in itertools.product(*reversed(rs))]
    activate_local_refinement(lrcp, np.array(rows, dtype=np.int32).reshape(len(rows), len(rs)))


def deactivate_overwritten_control_points(lrcp: LocallyRefinedCP) -> None:
    """``deactivate_overwritten_control_points!`` -- src/control_points.jl:584-680.
    The ``Flag`` arithmetic (src/utils.jl:237-247; adjoint branch src/adjoint.jl:117-121) is an
    OR-propagation of "still visible" through the structural support of the refinement matrices."""
    for level in range(len(lrcp.local_refinements) - 2, -1, -1):
        lr, lr_next = lrcp.local_refinements[level], lrcp.local_refinements[level + 1]
        shape_next = lrcp.control_points_refined[level + 1].shape[:-1]
        shape_cur = lrcp.control_points_refined[level].shape[:-1]
        flag_next = np.ones(shape_next, dtype=bool)
        for r in lr_next.refinement_indices:
            flag_next[tuple(int(v) - 1 for v in r)] = False
        flag = flag_next                                    # L* with boolean support matrices
        for A, d in zip(lr_next.refinement_matrices, lr_next.dims_refinement):
            S = np.zeros((A.m, A.n), dtype=bool)
            for i in range(1, A.m + 1):
                c0, c1 = A.column_range(i)
                S[i - 1, c0 - 1:c1] = True
            flag = np.moveaxis(np.tensordot(S.T.astype(np.int64), flag.astype(np.int64), axes=(1, d - 1)), 0, d - 1) > 0
        assert flag.shape == shape_cur
        keep = [i for i, r in enumerate(lr.refinement_indices) if flag[tuple(int(v) - 1 for v in r)]]
        lrcp.local_refinements[level] = LocalRefinementO(
            lr.dims_refinement, lr.refinement_matrices,
            lr.refinement_indices[keep], lr.refinement_values[keep])


def error_informed_local_refinement(dims: List[DimO], lrcp: LocallyRefinedCP, error: np.ndarray,
                                    threshold_factor: float = 1.0) -> None:
    """``error_informed_local_refinement!`` -- src/control_points.jl:541-575."""
    cp_shape = lrcp.control_points_refined[-1].shape
    nin = len(dims)
    g = evaluate_adjoint([d.eval for d in dims], [d.sample_indices for d in dims],
                         [d.degree for d in dims], [0] * nin, error, cp_shape)
    grid_err = g.sum(axis=nin)
    threshold = threshold_factor * grid_err.sum() / grid_err.size
    hit = np.argwhere(np.asfortranarray(grid_err > threshold))
    # findall order = column-major
    order = np.lexsort(tuple(hit[:, d] for d in range(nin)))
    activate_local_refinement(lrcp, (hit[order] + 1).astype(np.int32))
