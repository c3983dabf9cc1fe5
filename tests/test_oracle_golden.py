"""Pins the CPU oracle (numpy + C restatements) against every golden vector / known answer the
reference's own tests hold for the hot path.  CPU only."""
import numpy as np
import pytest
from helpers import julia_isapprox, load_g1, load_g2, rel_err

from oracle import oracle_c as OC
from oracle import oracle_np as O


def _dims(n_cp, deg, n_s, dtype=np.float32, **kw):
    return [O.make_dimension(c, p, n, dtype=dtype, **kw) for c, p, n in zip(n_cp, deg, n_s)]


def _eval(dims, cp, der=None, weights=None, impl=O):
    der = [0] * len(dims) if der is None else der
    return impl.evaluate([d.eval for d in dims], [d.sample_indices for d in dims],
                         [d.degree for d in dims], der, cp, weights)


# ---- G10: knot vector -- test/test_knot_vector.jl:26-41 -------------------------------------


def test_g10_knot_vector_values():
    kv, mu, ka = O.clamped_knot_vector(5, 2)
    assert julia_isapprox(kv, np.array([0, 1 / 3, 2 / 3, 1], dtype=np.float32))
    assert mu.tolist() == [3, 1, 1, 3]
    assert julia_isapprox(ka, np.array([0, 0, 0, 1 / 3, 2 / 3, 1, 1, 1], dtype=np.float32))
    kv2, mu2, ka2 = O.clamped_knot_vector(5, 2, extent=(5.0, 7.0))
    assert julia_isapprox(kv2, 2 * kv + 5)
    assert mu2.tolist() == mu.tolist()
    assert julia_isapprox(ka2, 2 * ka + 5)
    O.clamped_knot_vector(10, 3, extent=(4, 8), distribution="random")
    with pytest.raises(AssertionError):
        O.clamped_knot_vector(5, 10)


# ---- G5: partition of unity -- test/test_spline_dimension.jl:10-29 --------------------------


@pytest.mark.parametrize("degree", range(6))
@pytest.mark.parametrize("distribution", ["equispaced", "random"])
def test_g5_partition_of_unity(degree, distribution):
    d = O.make_dimension(25, degree, 500, distribution=distribution, rng=np.random.default_rng(degree))
    B = d.eval[:, :, 0]
    assert B.shape == (500, degree + 1)
    assert np.all(B >= 0)
    assert np.allclose(B.sum(axis=1), 1, rtol=np.sqrt(np.finfo(np.float32).eps))
    # C restatement agrees bit for bit with the numpy one (same operation order, no contraction)
    idx_c = OC.span_indices(d.sample_points, d.knots_all, degree)
    assert np.array_equal(idx_c, d.sample_indices)
    tab_c = OC.basis_tables(d.knots_all, d.sample_points, idx_c, degree, 0)
    assert np.array_equal(tab_c, d.eval)


# ---- G6: derivatives vs finite differences -- test/test_spline_dimension.jl:31-66 ------------


def test_g6_derivative_tables_vs_finite_differences():
    d = O.make_dimension(10, 3, 5000, dtype=np.float64, max_derivative_order=2)
    dt = np.diff(d.sample_points)
    data = O.decompress(d.eval, d.sample_indices, 3, d.n_basis, 0)
    data1 = O.decompress(d.eval, d.sample_indices, 3, d.n_basis, 1)
    data2 = O.decompress(d.eval, d.sample_indices, 3, d.n_basis, 2)
    fd = np.diff(data, axis=0) / dt[:, None]
    fd2 = np.diff(fd, axis=0) / dt[1:, None]
    assert julia_isapprox(data1[:-1], fd, rtol=1e-2)
    assert julia_isapprox(data2[2:], fd2, rtol=1e-2)
    tab_c = OC.basis_tables(d.knots_all, d.sample_points, d.sample_indices, 3, 2)
    assert np.array_equal(tab_c, d.eval)


# ---- G4: ones in, ones out -- test/test_spline_grid.jl:11-27 --------------------------------


def test_g4_ones_in_ones_out_all_degrees():
    for n_basis in range(2, 11):
        for degree in range(1, n_basis):
            d = O.make_dimension(n_basis, degree, 100)
            cp = np.ones((n_basis, 1), dtype=np.float32, order="F")
            ev = _eval([d], cp)
            assert np.allclose(ev, 1, rtol=np.sqrt(np.finfo(np.float32).eps)), (n_basis, degree)
            assert np.array_equal(_eval([d], cp, impl=OC), ev)


# ---- G1: the golden grid -- test/test_spline_grid.jl:29-53 -----------------------------------


def test_g1_golden_grid():
    g, cp, ev = load_g1()
    dims = _dims(g["n_control_points"], g["degree"], g["n_sample_points"])
    assert dims[0].sample_indices.tolist() == [4, 4, 4, 5, 5, 5, 5]
    assert dims[1].sample_indices.tolist() == [3, 3, 4, 4, 5, 5, 6, 6, 6]
    cp32 = np.asfortranarray(cp.astype(np.float32))
    for impl in (O, OC):
        out = _eval(dims, cp32, impl=impl)
        assert out.shape == (7, 9, 2)
        assert julia_isapprox(out, ev.astype(np.float32))      # the reference's own criterion
        assert rel_err(out, ev) < 2e-7                         # and much tighter than that
    out_loops = O.evaluate_loops([d.eval for d in dims], [d.sample_indices for d in dims],
                                 g["degree"], [0, 0], cp32)
    assert np.array_equal(out_loops, _eval(dims, cp32))
    assert np.array_equal(_eval(dims, cp32, impl=OC), out_loops)
    # Float64 tables reproduce the printed Float64 digits much more closely
    dims64 = _dims(g["n_control_points"], g["degree"], g["n_sample_points"], dtype=np.float64)
    out64 = _eval(dims64, np.asfortranarray(cp))
    assert rel_err(out64, ev) < 1e-7


# ---- G3: NURBS circle -- test/test_nurbs_grid.jl:25-52 ----------------------------------------


def test_g3_nurbs_circle():
    kv = np.array([0, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi], dtype=np.float32)
    d = O.make_dimension(9, 2, 500, knot_values=kv, multiplicities=[3, 2, 2, 2, 3])
    w = np.ones(9, dtype=np.float32)
    w[1::2] = np.float32(1 / np.sqrt(2))
    cp = np.asfortranarray(np.array([[1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1], [1, -1], [1, 0]],
                                    dtype=np.float32))
    for impl in (O, OC):
        ev = _eval([d], cp, weights=w, impl=impl)
        r2 = ev[:, 0].astype(np.float64) ** 2 + ev[:, 1].astype(np.float64) ** 2
        assert np.all(np.abs(r2 - 1) <= np.sqrt(np.finfo(np.float32).eps))
        assert len({tuple(r) for r in ev[1:]}) == 499


# ---- G7: refinement preserves geometry -- test/test_refinement.jl:11-58 ------------------------


def _g7_grid():
    rng = np.random.default_rng(1)
    dims = [O.make_dimension(c, p, n, distribution="random", rng=rng)
            for c, p, n in zip((5, 8, 6), (4, 2, 3), (15, 20, 25))]
    cp = np.asfortranarray(rng.random((5, 8, 6, 2)).astype(np.float32))
    return dims, cp


def test_g7_knot_insertion_preserves_geometry():
    dims, cp = _g7_grid()
    ev0 = _eval(dims, cp)
    for dr in range(3):
        d = dims[dr]
        new_knot = d.knot_values.dtype.type(0.25)
        ka_old = d.knots_all
        kv, mu, k = O.insert_knot(d.knot_values, d.multiplicities, new_knot)
        R = O.boehm_matrix(ka_old, d.degree, k, new_knot)
        assert R.validate() == []
        dn = O.DimO(d.degree, 0, kv, mu, O.expand_knot_vector(kv, mu), d.sample_points)
        dn.rebuild()
        dims[dr] = dn
        shape = list(cp.shape)
        shape[dr] += 1
        cp_new = np.zeros(shape, dtype=cp.dtype, order="F")
        O.mult(cp_new, [R], cp, [dr + 1])
        assert (R.m, R.n) == (cp_new.shape[dr], cp_new.shape[dr] - 1)
        cp = cp_new
        assert julia_isapprox(_eval(dims, cp), ev0)


def test_g7_default_refine_preserves_geometry():
    dims, cp = _g7_grid()
    ev0 = _eval(dims, cp)
    old_shape = cp.shape
    for dr in range(3):
        d = dims[dr]
        kv, mu, R = O.refine_dimension(d.knot_values, d.multiplicities, d.degree)
        assert R.validate() == []
        dn = O.DimO(d.degree, 0, kv, mu, O.expand_knot_vector(kv, mu), d.sample_points)
        dn.rebuild()
        dims[dr] = dn
        shape = list(cp.shape)
        shape[dr] = R.m
        cp_new = np.zeros(shape, dtype=cp.dtype, order="F")
        O.mult(cp_new, [R], cp, [dr + 1])
        # literal loops == dense mode-n product == C restatement
        assert np.allclose(cp_new, O.mult_dense([R], cp, [dr + 1]), rtol=1e-5, atol=1e-6)
        cp_c = np.zeros_like(cp_new)
        OC.mult(cp_c, [R], cp, [dr + 1])
        assert np.array_equal(cp_c, cp_new)
        assert (R.m, R.n) == (cp_new.shape[dr], old_shape[dr])
        cp = cp_new
        assert julia_isapprox(_eval(dims, cp), ev0)


# ---- RefinementMatrix unit tests -- test/test_refinement_matrix.jl:16-75 -----------------------


def test_refinement_matrix_unit_and_products():
    M = O.rmeye(100)
    M2 = O.refmat_matmul(M, M)
    assert (M2.m, M2.n) == (100, 100)
    assert np.array_equal(M2.row_pointer, M.row_pointer)
    assert np.array_equal(M2.column_start, M.column_start)
    assert np.array_equal(M2.nzval, M.nzval)
    rng = np.random.default_rng(1)
    A = O.refmat_from_dense(np.diag(rng.random(100)))
    B = O.refmat_from_dense(np.diag(rng.random(100)))
    assert np.array_equal(O.refmat_matmul(A, B).nzval, A.nzval * B.nzval)

    def brand(m, n, lo, up):
        a = rng.random((m, n))
        i, j = np.indices((m, n))
        a[(i - j > lo) | (j - i > up)] = 0
        return a

    Ad, Bd = brand(100, 100, 4, 3), brand(100, 200, 3, 5)
    Cm = O.refmat_matmul(O.refmat_from_dense(Ad), O.refmat_from_dense(Bd))
    assert Cm.validate() == []
    assert julia_isapprox(Cm.dense(), Ad @ Bd)


def test_refinement_matrix_invalid_rows():
    M = np.zeros((2, 50))
    M[0, :20] = 1
    M[1, 38:] = 1
    assert O.refmat_from_dense(M).validate() == [2]          # "Invalid rows: [2]."
    M = np.zeros((3, 50))
    M[0, :25] = 1
    M[2, 25:] = 1
    assert O.refmat_from_dense(M).validate() == [2, 3]       # "Invalid rows: [2, 3]."


def test_g8_dyadic_degree2_rows():
    d = O.make_dimension(6, 2, 50)
    _, _, R = O.refine_dimension(d.knot_values, d.multiplicities, 2)
    D = R.dense()
    assert D.shape == (10, 6)
    assert np.allclose(D[0, :1], [1]) and np.allclose(D[1, :2], [.5, .5])
    assert np.allclose(D[2, 1:3], [.75, .25]) and np.allclose(D[3, 1:3], [.25, .75])
    assert np.allclose(D[-2, -2:], [.5, .5]) and np.allclose(D[-1, -1:], [1])
    assert np.allclose(D.sum(axis=1), 1)


# ---- G8: THB bookkeeping -- test/test_local_refinement.jl:11-46 --------------------------------


def _thb_docs(n_samples=(500, 500), dtype=np.float32, levels=1):
    dims = _dims((6, 6), (2, 2), n_samples, dtype=dtype)
    cp0 = O.unit_cp_grid((6, 6, 3), dtype)
    dims, lr = O.add_default_local_refinement(dims, cp0)
    return dims, lr


def test_g8_local_refinement_counts_and_geometry():
    dims, lr = _thb_docs()
    assert lr.control_points_refined[-1].shape == (10, 10, 3)
    assert lr.n_control_points() == 36
    before = lr.control_points_refined[-1].copy()
    O.activate_local_control_point_range(lr, (1, 4), (1, 6))
    O.activate_local_control_point_range(lr, (1, 6), (1, 2))
    O.activate_local_control_point_range(lr, (9, 10), (7, 10))
    O.deactivate_overwritten_control_points(lr)
    O.lrcp_evaluate(lr, fast=False)
    assert lr.n_control_points() == 63
    assert lr.local_refinements[0].refinement_indices.shape[0] == 27
    assert julia_isapprox(lr.control_points_refined[-1], before)
    # second level of the docs example (docs/src/theory_local_refinement.md:56-63)
    dims, lr = O.add_default_local_refinement(dims, lr)
    assert lr.control_points_refined[-1].shape == (18, 18, 3)
    O.activate_local_control_point_range(lr, (5, 12), (1, 4))
    O.activate_local_control_point_range(lr, (7, 8), (5, 6))
    O.deactivate_overwritten_control_points(lr)
    fine_before = lr.control_points_refined[-1].copy()
    O.lrcp_evaluate(lr)
    assert julia_isapprox(lr.control_points_refined[-1], fine_before)


# ---- G2: error-informed refinement -- test/test_local_refinement.jl:48-67 ----------------------


def test_g2_error_informed_index_set():
    g, expected = load_g2()
    dims, lr = _thb_docs(n_samples=(50, 50))
    err = np.zeros((50, 50, 3), dtype=np.float32, order="F")
    err[19:40, 9:30, 1] = 1
    O.error_informed_local_refinement(dims, lr, err)
    assert lr.local_refinements[-1].refinement_indices.dtype == np.int32
    assert np.array_equal(lr.local_refinements[-1].refinement_indices, expected)
    # the C adjoint (atomics) gives the same set
    gc = OC.evaluate_adjoint([d.eval for d in dims], [d.sample_indices for d in dims], [2, 2], [0, 0], err,
                             (10, 10, 3))
    ge = gc.sum(axis=2)
    hit = np.argwhere(ge > ge.sum() / ge.size)
    order = np.lexsort((hit[:, 0], hit[:, 1]))
    assert np.array_equal((hit[order] + 1).astype(np.int32), expected)


# ---- G9: forward/adjoint consistency -- test/test_LinearMapsExt.jl:14-62 -----------------------


def _lsqr(matvec, rmatvec, m, n, b):
    from scipy.sparse.linalg import LinearOperator, lsqr
    A = LinearOperator((m, n), matvec=matvec, rmatvec=rmatvec, dtype=np.float64)
    return lsqr(A, b, atol=1e-12, btol=1e-12, iter_lim=2000)[0]


def test_g9_lsqr_recovers_control_points():
    rng = np.random.default_rng(1)
    dims = _dims((12, 10), (3, 4), (26, 73), dtype=np.float64)
    cp_shape = (12, 10, 3)
    cp = np.asfortranarray(rng.random(cp_shape))
    tabs, idx, deg = [d.eval for d in dims], [d.sample_indices for d in dims], [3, 4]
    target = OC.evaluate(tabs, idx, deg, [0, 0], cp)
    assert np.allclose(target, O.evaluate(tabs, idx, deg, [0, 0], cp), rtol=1e-13, atol=1e-15)

    def fwd(x):
        return OC.evaluate(tabs, idx, deg, [0, 0], np.asfortranarray(x.reshape(cp_shape, order="F"))).ravel(order="F")

    def adj(y):
        return OC.evaluate_adjoint(tabs, idx, deg, [0, 0], np.asfortranarray(y.reshape(target.shape, order="F")),
                                   cp_shape).ravel(order="F")

    # adjoint identity <Ax, y> == <x, A'y>
    x, y = rng.random(cp.size), rng.random(target.size)
    assert abs(fwd(x) @ y - x @ adj(y)) <= 1e-12 * abs(fwd(x) @ y)
    g_np = O.evaluate_adjoint(tabs, idx, deg, [0, 0], np.asfortranarray(y.reshape(target.shape, order="F")), cp_shape)
    assert rel_err(adj(y), g_np.ravel(order="F")) < 1e-13
    fit = _lsqr(fwd, adj, target.size, cp.size, target.ravel(order="F"))
    assert np.allclose(fit, cp.ravel(order="F"), rtol=1e-5)


def test_g9_locally_refined_lsqr():
    rng = np.random.default_rng(1)
    dims = _dims((12, 10), (3, 4), (26, 73), dtype=np.float64)
    dims, lr = O.add_default_local_refinement(dims, O.unit_cp_grid((12, 10, 3), np.float64))
    O.activate_local_control_point_range(lr, (1, 4), (1, 6))
    O.activate_local_control_point_range(lr, (1, 6), (1, 2))
    O.activate_local_control_point_range(lr, (9, 10), (7, 10))
    O.deactivate_overwritten_control_points(lr)
    n_cp = lr.n_control_points()
    vals = rng.random((n_cp, 3))
    tabs, idx, deg = [d.eval for d in dims], [d.sample_indices for d in dims], [3, 4]
    fine_shape = lr.control_points_refined[-1].shape

    def fwd(x):
        lr.set_values(x.reshape((n_cp, 3), order="F"))
        O.lrcp_evaluate(lr)
        return OC.evaluate(tabs, idx, deg, [0, 0], np.asfortranarray(lr.control_points_refined[-1])).ravel(order="F")

    def adj(y):
        g = OC.evaluate_adjoint(tabs, idx, deg, [0, 0], np.asfortranarray(y.reshape((26, 73, 3), order="F")), fine_shape)
        lr.control_points_refined[-1][...] = g
        O.lrcp_evaluate_adjoint(lr)
        return lr.get_values().ravel(order="F")

    target = fwd(vals.ravel(order="F")).copy()
    x, y = rng.random(n_cp * 3), rng.random(target.size)
    lhs, rhs = fwd(x) @ y, x @ adj(y)
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    fit = _lsqr(fwd, adj, target.size, n_cp * 3, target)
    assert np.allclose(fit, vals.ravel(order="F"), rtol=1e-5)


# ---- hierarchy apply: literal K5-K8 (numpy loops) == C restatement == dense ------------------------


def test_hierarchy_literal_vs_c():
    rng = np.random.default_rng(5)
    dims, lr = _thb_docs(n_samples=(20, 20))
    O.activate_local_control_point_range(lr, (2, 5), (3, 8))
    O.deactivate_overwritten_control_points(lr)
    lr.set_values(rng.random((lr.n_control_points(), 3)).astype(np.float32))
    O.lrcp_evaluate(lr, fast=False)
    fine = lr.control_points_refined[-1].copy()
    # C path
    c0 = np.zeros((6, 6, 3), dtype=np.float32, order="F")
    OC.scatter_active(c0, lr.local_refinements[0].refinement_indices, lr.local_refinements[0].refinement_values)
    c1 = np.zeros((10, 10, 3), dtype=np.float32, order="F")
    OC.mult(c1, lr.local_refinements[1].refinement_matrices, c0, [1, 2])
    OC.scatter_active(c1, lr.local_refinements[1].refinement_indices, lr.local_refinements[1].refinement_values)
    assert np.array_equal(c1, fine)
    # adjoint chain
    g = np.asfortranarray(rng.random((10, 10, 3)).astype(np.float32))
    lr.control_points_refined[-1][...] = g
    O.lrcp_evaluate_adjoint(lr, fast=False)
    v1_np = lr.local_refinements[1].refinement_values.copy()
    v0_np = lr.local_refinements[0].refinement_values.copy()
    gc = g.copy(order="F")
    v1 = np.zeros(v1_np.shape, dtype=np.float32, order="F")
    OC.gather_zero_active(v1, gc, lr.local_refinements[1].refinement_indices)
    b = np.zeros((6, 6, 3), dtype=np.float32, order="F")
    OC.mult_adjoint(b, lr.local_refinements[1].refinement_matrices, gc, [1, 2])
    v0 = np.zeros(v0_np.shape, dtype=np.float32, order="F")
    OC.gather_zero_active(v0, b, lr.local_refinements[0].refinement_indices)
    assert np.array_equal(v1, v1_np)
    assert np.allclose(v0, v0_np, rtol=1e-6, atol=1e-7)
