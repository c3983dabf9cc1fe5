"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs, the
reference's golden fixtures, and the reference's own test assertions re-stated for the package API.

Tolerances (BASELINE.json north_star): bit-exact for span / sample indices, <= 1e-12 relative for
Float64, <= 1e-5 relative for Float32 (norm-wise, like the reference's `≈`).
"""
import logging

import numpy as np
import pytest
from helpers import TOL, julia_isapprox, load_g1, load_g2, max_rel_err, rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def S():
    from gpu_helpers import sg
    return sg()


@pytest.fixture(autouse=True)
def _restore_switches(S):
    """Process-wide test hooks (kernel policy, adjoint plans) go back to their defaults after every test."""
    yield
    S.set_adjoint_plans(True)
    S.set_kernel_policy(0)


def _tol(ft):
    return 1e-5 if ft == "Float32" else 1e-12


# ---------------------------------------------------------------------------------------------
# K1 / K2 / K9 / K10: per-dimension tables
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("ft", ["Float32", "Float64"])
@pytest.mark.parametrize("distribution", ["equispaced", "random"])
def test_basis_tables_bit_exact_all_degrees(S, ft, distribution):
    """test/test_spline_dimension.jl:10-29 shapes (25 basis functions, 500 samples, degree 0..5) plus degrees
    up to 9; indices AND tables must be bit-identical to the oracle."""
    from gpu_helpers import O, OC, oracle_dim
    rng = np.random.default_rng(3)
    for degree in range(0, 10):
        for mdo in sorted({0, min(1, degree), min(2, degree), degree}):
            sd = S.SplineDimension(25, degree, 500, float_type=ft, max_derivative_order=mdo,
                                   distribution=distribution, rng=rng)
            tab, idx = oracle_dim(sd, OC)
            assert np.array_equal(S.to_numpy(sd.sample_indices), idx), (degree, mdo)
            got = S.to_numpy(sd.eval)
            assert got.shape == (500, degree + 1, mdo + 1)
            assert np.array_equal(got, tab), (degree, mdo, np.abs(got - tab).max())
            B = got[:, :, 0]
            assert np.all(B >= 0)
            assert np.allclose(B.sum(axis=1), 1, rtol=np.sqrt(np.finfo(B.dtype).eps))
        # the numpy oracle agrees too (same operation order)
        tab_np, idx_np = oracle_dim(sd, O)
        assert np.array_equal(idx_np, idx) and np.array_equal(tab_np, tab)


def test_knot_vector_expand_and_values(S):
    """test/test_knot_vector.jl:26-41 (G10) through K9 on the device."""
    kv = S.KnotVector.clamped(5, 2)
    assert julia_isapprox(S.to_numpy(kv.knot_values), np.array([0, 1 / 3, 2 / 3, 1], dtype=np.float32))
    assert S.to_numpy(kv.multiplicities).tolist() == [3, 1, 1, 3]
    assert julia_isapprox(S.to_numpy(kv.knots_all), np.array([0, 0, 0, 1 / 3, 2 / 3, 1, 1, 1], dtype=np.float32))
    kv2 = S.KnotVector.clamped(5, 2, extent=(5.0, 7.0))
    assert julia_isapprox(S.to_numpy(kv2.knots_all), 2 * S.to_numpy(kv.knots_all) + 5)
    S.KnotVector.clamped(10, 3, extent=(4, 8), distribution="random")
    with pytest.raises(AssertionError):
        S.KnotVector.clamped(5, 10)


def test_span_indices_edge_cases(S):
    """Repeated interior knots, samples on knots, out-of-range and NaN samples: bit-exact vs the literal scan."""
    from gpu_helpers import O
    for npdt, ft in ((np.float32, "Float32"), (np.float64, "Float64")):
        kvals = np.array([0, 0.25, 0.5, 0.75, 1.0], dtype=npdt)
        kv = S.KnotVector(kvals, np.array([3, 2, 1, 2, 3], dtype=np.int32))
        sp = np.array([-1.0, 0.0, 0.1, 0.25, np.nextafter(npdt(0.25), npdt(0)), 0.5, 0.75, 0.9999, 1.0, 2.0, np.nan],
                      dtype=npdt)
        n_basis = int(kv.knots_all.numel()) - 3
        sd = S.SplineDimension.from_fields(2, 0, kv, S.to_device(sp), torch.zeros(len(sp), dtype=torch.int32, device="cuda"),
                                           S.jl_zeros((len(sp), 3, 1), kv.dtype, kv.device))
        S.set_sample_indices_(sd)
        ref = O.span_indices(sp, S.to_numpy(kv.knots_all), 2)
        assert np.array_equal(S.to_numpy(sd.sample_indices), ref)
        assert ref[-1] == n_basis and ref[-2] == n_basis          # NaN and t > t_max -> last span


def test_derivative_tables_vs_finite_differences(S):
    """test/test_spline_dimension.jl:31-66 (G6) through K2 + K10 (decompress)."""
    sd = S.SplineDimension(10, 3, 5000, max_derivative_order=2, float_type="Float64")
    sp = S.to_numpy(sd.sample_points)
    dt = np.diff(sp)
    data, d1, d2 = (S.to_numpy(S.decompress(sd, derivative_order=k)) for k in range(3))
    fd = np.diff(data, axis=0) / dt[:, None]
    fd2 = np.diff(fd, axis=0) / dt[1:, None]
    assert julia_isapprox(d1[:-1], fd, rtol=1e-2)
    assert julia_isapprox(d2[2:], fd2, rtol=1e-2)


# ---------------------------------------------------------------------------------------------
# K3: evaluate!
# ---------------------------------------------------------------------------------------------


def test_g4_ones_in_ones_out(S):
    """test/test_spline_grid.jl:11-27: degree 1..n_basis-1 up to 9, including the Bezier case."""
    for n_basis in range(2, 11):
        for degree in range(1, n_basis):
            grid = S.SplineGrid(S.SplineDimension(n_basis, degree, 100), 1)
            grid.control_points.fill_(1)
            S.evaluate_(grid)
            assert np.allclose(S.to_numpy(grid.eval), 1, rtol=np.sqrt(np.finfo(np.float32).eps)), (n_basis, degree)


def test_g1_golden_grid(S):
    """test/test_spline_grid.jl:29-53: the reference's golden vector, end to end (K1+K2+K3)."""
    from gpu_helpers import oracle_evaluate
    g, cp, ev = load_g1()
    dims = tuple(S.SplineDimension(c, p, n) for c, p, n in zip(g["n_control_points"], g["degree"], g["n_sample_points"]))
    assert S.to_numpy(dims[0].sample_indices).tolist() == [4, 4, 4, 5, 5, 5, 5]
    assert S.to_numpy(dims[1].sample_indices).tolist() == [3, 3, 4, 4, 5, 5, 6, 6, 6]
    grid = S.SplineGrid(dims, g["Nout"])
    S.copyto_(grid.control_points, cp.astype(np.float32))
    S.evaluate_(grid)
    got = S.to_numpy(grid.eval)
    assert got.shape == (7, 9, 2)
    assert julia_isapprox(got, ev.astype(np.float32))             # the reference's own assertion
    assert rel_err(got, ev) < 2e-7
    assert rel_err(got, oracle_evaluate(grid, cp.astype(np.float32))) < 1e-6


def test_g3_nurbs_circle(S):
    """test/test_nurbs_grid.jl:25-52."""
    kv = S.KnotVector(np.array([0, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi], dtype=np.float32),
                      np.array([3, 2, 2, 2, 3], dtype=np.int32))
    sd = S.SplineDimension(9, 2, 500, knot_vector=kv)
    grid = S.NURBSGrid(sd, 2)
    grid.weights[1::2] = 1 / np.sqrt(2)
    S.copyto_(grid.control_points,
              np.array([[1, 0], [1, 1], [0, 1], [-1, 1], [-1, 0], [-1, -1], [0, -1], [1, -1], [1, 0]], dtype=np.float32))
    S.evaluate_(grid)
    pts = S.to_numpy(grid.eval).astype(np.float64)
    assert np.all(np.abs(pts[:, 0] ** 2 + pts[:, 1] ** 2 - 1) <= np.sqrt(np.finfo(np.float32).eps))
    assert len({tuple(r) for r in S.to_numpy(grid.eval)[1:]}) == 499


CASES = [
    # n_cp, degree, n_samples, nout, float type, mdo, nurbs
    ((7,), (3,), (41,), 1, "Float64", 0, False),
    ((9,), (5,), (33,), 3, "Float32", 0, True),
    ((10, 10, 5), (2, 3, 2), (50, 50, 25), 4, "Float64", 0, False),       # BASELINE config C1
    ((5, 6), (3, 2), (7, 9), 2, "Float32", 1, False),
    ((12, 10), (3, 4), (26, 73), 3, "Float64", 2, False),
    ((12, 10), (3, 4), (26, 73), 3, "Float64", 0, True),
    ((6, 6), (2, 2), (50, 50), 3, "Float32", 0, False),
    ((5, 8, 6), (4, 2, 3), (15, 20, 25), 2, "Float32", 0, False),
    ((4, 5, 4, 3), (1, 2, 3, 2), (6, 7, 5, 4), 5, "Float64", 0, False),   # Nin = 4, Nout > chunk
    ((10,), (9,), (100,), 1, "Float32", 0, False),                        # Bezier, degree 9
    ((4, 4), (3, 3), (9, 300), 1, "Float64", 1, False),                   # single span per dimension
    ((16, 16), (3, 3), (128, 96), 3, "Float32", 1, False),
    ((16, 16, 16), (3, 3, 3), (40, 36, 44), 1, "Float64", 0, False),
    ((16, 16), (3, 3), (128, 96), 3, "Float32", 0, True),
    ((20, 3), (2, 2), (3, 50), 2, "Float64", 0, False),                   # fewer samples than spans
    # 3-D / 4-D shapes with n_1 a multiple of 64
    ((20, 11, 9), (3, 3, 3), (128, 400, 20), 1, "Float64", 0, False),
    ((40, 9, 7), (2, 3, 2), (256, 400, 12), 2, "Float32", 1, False),
    ((100, 8, 6), (2, 2, 2), (128, 400, 20), 1, "Float64", 0, False),     # > 32 control indices per tile: device-side fallback
    ((12, 6, 9, 5), (3, 2, 3, 1), (64, 20, 40, 5), 2, "Float64", 0, False),   # Nin = 4
]


@pytest.mark.parametrize("policy", [1, 0, 2])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-{c[1]}-{c[4]}{'-nurbs' if c[6] else ''}" for c in CASES])
def test_evaluate_and_adjoint_vs_oracle(S, case, policy):
    """K3 and K4 against the C oracle (the reference's algorithm) for every kernel policy
    (1 = generic kernels, 0 = automatic dispatch, 2 = tiled fast paths forced)."""
    from gpu_helpers import make_grid, oracle_adjoint, oracle_evaluate
    n_cp, deg, n_s, nout, ft, mdo, nurbs = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=mdo, nurbs=nurbs, seed=11)
    tol = _tol(ft)
    S.set_kernel_policy(policy)
    try:
        ders = [(0,) * len(n_cp)]
        if mdo and not nurbs:
            ders += [tuple(min(mdo, p) if d == k else 0 for d, p in enumerate(deg)) for k in range(len(n_cp))]
            ders += [tuple(min(mdo, p) for p in deg)]
        for der in ders:
            S.evaluate_(grid, derivative_order=der)
            ref = oracle_evaluate(grid, cp, der, w)
            assert rel_err(S.to_numpy(grid.eval), ref) <= tol, (der, S.last_variant())
            assert max_rel_err(S.to_numpy(grid.eval), ref) <= 10 * tol, (der, S.last_variant())
            e = np.asfortranarray(rng.random(ref.shape).astype(ref.dtype))
            g = torch.full_like(grid.control_points.obtain(), 7.0)    # must be overwritten (zero fill, adjoint.jl:61)
            S.evaluate_adjoint_(grid, derivative_order=der, eval=S.to_device(e), control_points=g, allow_nurbs=nurbs)
            gref = oracle_adjoint(grid, e, der, w)
            assert rel_err(S.to_numpy(g), gref) <= tol, (der, S.last_variant())
            assert max_rel_err(S.to_numpy(g), gref) <= 10 * tol, (der, S.last_variant())
    finally:
        S.set_kernel_policy(0)


MARCH2_CASES = [((20, 11, 9), (3, 3, 3), (128, 400, 20), 1, "Float64", 0, False),
                ((100, 8, 6), (2, 2, 2), (128, 400, 20), 1, "Float64", 0, False),
                ((30, 12, 25), (2, 1, 1), (256, 40, 300), 2, "Float32", 0, False),
                ((9, 40, 7), (3, 3, 3), (130, 45, 33), 1, "Float64", 0, False)]      # ~1 sample per span in dim 2, ragged n1


@pytest.mark.parametrize("m2f", ["1", "0"], ids=["fused", "post2"])
@pytest.mark.parametrize("case,env,variant", [(c, "SG_ADJ_MARCH2", "adjoint_march2") for c in MARCH2_CASES],
                         ids=[f"march2-{c[0]}-{c[1]}" for c in MARCH2_CASES])
def test_forced_double_march_vs_oracle(S, case, env, variant, m2f, monkeypatch):
    """SG_ADJ_MARCH2=1 forces the 3-D double march over dimensions 3 and 2 on small shapes; against the C oracle."""
    from gpu_helpers import make_grid, oracle_adjoint
    n_cp, deg, n_s, nout, ft, mdo, nurbs = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=0, seed=21)
    e = np.asfortranarray(rng.random(tuple(n_s) + (nout,)).astype(cp.dtype))
    g = torch.full_like(grid.control_points.obtain(), -3.0)
    monkeypatch.setenv(env, "1")
    monkeypatch.setenv("SG_ADJ_M2G", m2f)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g)
    assert S.last_variant().startswith(variant)
    gref = oracle_adjoint(grid, e)
    assert rel_err(S.to_numpy(g), gref) <= _tol(ft)
    assert max_rel_err(S.to_numpy(g), gref) <= 10 * _tol(ft)


POST2_CASES = [
    # n_cp, degree, n_samples, nout, float type: shapes that exercise sg_adj_post2_kernel (halo combine + dimension 1)
    ((20, 11, 9), (3, 3, 3), (1400, 40, 20), 1, "Float64"),      # sample range of an i1 block > SG_POST2_JMAX: three pieces
    ((300, 9, 8), (3, 2, 2), (640, 30, 24), 2, "Float64"),       # three blocks of control indices, Nout 2
    ((150, 14, 12), (1, 1, 1), (256, 50, 40), 3, "Float32"),     # degree 1, Nout 3, Float32
    ((40, 23, 21), (5, 3, 3), (200, 70, 64), 1, "Float64"),      # degree 5 in dimension 1; tiles2 not a multiple of G2
    ((16, 10, 40), (2, 3, 3), (128, 33, 17), 1, "Float64"),      # fewer samples than spans in dimension 3
]


@pytest.mark.parametrize("m2f", ["1", "0"], ids=["fused", "post2"])
@pytest.mark.parametrize("distribution", ["equispaced", "random"])
@pytest.mark.parametrize("case", POST2_CASES, ids=[f"{c[0]}-{c[1]}-{c[4]}" for c in POST2_CASES])
def test_double_march_post_kernel_vs_oracle(S, case, distribution, m2f, monkeypatch):
    """Double march + fused post kernel (sg_adjoint_post2.cuh) against the C oracle; control points pre-filled with
    garbage because this pipeline does its own zero fill; deterministic."""
    from gpu_helpers import make_grid, oracle_adjoint
    n_cp, deg, n_s, nout, ft = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=0, seed=41, distribution=distribution)
    e = np.asfortranarray(rng.random(tuple(n_s) + (nout,)).astype(cp.dtype))
    g = torch.full_like(grid.control_points.obtain(), -7.0)
    monkeypatch.setenv("SG_ADJ_MARCH2", "1")
    monkeypatch.setenv("SG_ADJ_M2G", m2f)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g)
    assert S.last_variant().startswith("adjoint_march2")
    gref = oracle_adjoint(grid, e)
    assert rel_err(S.to_numpy(g), gref) <= _tol(ft)
    assert max_rel_err(S.to_numpy(g), gref) <= 10 * _tol(ft)
    g2 = torch.full_like(g, 9.0)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g2)
    if distribution == "equispaced":
        assert torch.equal(g, g2)
    # other chunkings of dimension 3 (CTA-count target) give the same result up to summation order
    for ctas in ("1", "100000"):
        monkeypatch.setenv("SG_ADJ_M2_CTAS", ctas)
        g3 = torch.full_like(g, 1.0)
        S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g3)
        assert rel_err(S.to_numpy(g3), gref) <= _tol(ft), ctas


FUSED_CASES = [
    # n_cp, degree, n_samples, nout, float type, distribution: shapes the fused double march (sg_adjoint_march2g.cuh) takes
    ((36, 35, 20), (3, 3, 3), (256, 128, 64), 1, "Float64", "equispaced"),     # ~7.8 samples per span in dimension 1
    ((70, 20, 12), (2, 2, 2), (384, 72, 40), 2, "Float32", "equispaced"),      # Nout 2, Float32, degree 2
    ((50, 30, 9), (1, 1, 1), (200, 100, 30), 3, "Float64", "equispaced"),      # ragged last column block, degree 1
    ((40, 23, 21), (5, 3, 3), (200, 70, 64), 1, "Float64", "equispaced"),      # degree 5 in dimension 1
    ((10, 20, 12), (3, 3, 3), (512, 80, 40), 1, "Float64", "equispaced"),      # 73 samples per span: support = whole block
    ((67, 20, 12), (3, 3, 3), (128, 72, 40), 1, "Float64", "equispaced"),      # 2 samples per span: 70 control indices per block
    ((36, 35, 20), (3, 3, 3), (256, 128, 64), 1, "Float64", "random"),         # random sample points (may leave the fused path)
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=[f"{c[0]}-{c[1]}-{c[4]}-{c[5]}" for c in FUSED_CASES])
def test_fused_double_march_vs_oracle(S, case, monkeypatch):
    """Planned adjoint of 3-D grids: all three contractions in the march kernel + the halo-sum kernel, against the C
    oracle; deterministic; same result (up to summation order) as the plain entry point and the unfused pipeline."""
    from gpu_helpers import make_grid, oracle_adjoint
    n_cp, deg, n_s, nout, ft, distribution = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=0, seed=43, distribution=distribution)
    e = np.asfortranarray(rng.random(tuple(n_s) + (nout,)).astype(cp.dtype))
    g = torch.full_like(grid.control_points.obtain(), -7.0)
    monkeypatch.setenv("SG_ADJ_MARCH2", "1")
    monkeypatch.setenv("SG_ADJ_M2G", "1")                            # opt-in variant
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g)
    if distribution == "equispaced":
        assert S.last_variant() == "adjoint_march2_fused"
    else:
        assert S.last_variant().startswith("adjoint_march2")
    gref = oracle_adjoint(grid, e)
    assert rel_err(S.to_numpy(g), gref) <= _tol(ft)
    assert max_rel_err(S.to_numpy(g), gref) <= 10 * _tol(ft)
    g2 = torch.full_like(g, 9.0)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g2)
    assert torch.equal(g, g2)                                        # deterministic
    for ctas in ("1", "100000"):                                     # other chunkings of dimension 3
        monkeypatch.setenv("SG_ADJ_M2_CTAS", ctas)
        grid2, _, _, _ = make_grid(n_cp, deg, n_s, nout, ft, mdo=0, seed=43, distribution=distribution)
        g3 = torch.full_like(g, 1.0)
        S.evaluate_adjoint_(grid2, eval=S.to_device(e), control_points=g3)
        assert rel_err(S.to_numpy(g3), gref) <= _tol(ft), ctas
    monkeypatch.delenv("SG_ADJ_M2_CTAS")
    S.set_adjoint_plans(False)                                       # plain sg_evaluate_adjoint: decisions on device
    try:
        g4 = torch.full_like(g, 2.0)
        S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g4)
        assert S.last_variant() == "adjoint_march2"
        assert rel_err(S.to_numpy(g4), gref) <= _tol(ft)
    finally:
        S.set_adjoint_plans(True)


F2D_CASES = [
    # n_cp, degree, n_samples, nout, float type, nurbs, distribution: 2-D shapes the fused march (F1) takes
    ((20, 12), (3, 3), (1024, 200), 3, "Float32", False, "equispaced"),
    ((20, 12), (2, 3), (512, 90), 2, "Float64", False, "equispaced"),        # mixed degrees
    ((9, 40), (3, 2), (768, 70), 1, "Float64", False, "equispaced"),         # 128 samples per span of dimension 1
    ((30, 10), (3, 3), (600, 64), 2, "Float32", False, "equispaced"),        # ragged last column block
    ((24, 18), (3, 3), (1024, 160), 3, "Float32", True, "equispaced"),       # rational (NURBS extension)
    ((24, 18), (2, 2), (512, 100), 2, "Float64", True, "equispaced"),
    ((40, 12), (3, 3), (1024, 200), 3, "Float32", False, "random"),          # random knots
    ((200, 9), (1, 1), (512, 40), 4, "Float64", False, "equispaced"),        # 2.6 samples per span: 100+ control indices per block
]


@pytest.mark.parametrize("case", F2D_CASES, ids=[f"{c[0]}-{c[1]}-{c[4]}{'-nurbs' if c[5] else ''}-{c[6]}" for c in F2D_CASES])
def test_fused_2d_march_vs_oracle(S, case, monkeypatch):
    """Planned adjoint of 2-D grids: dimension 2 marched per sample column and dimension 1 contracted in the same kernel
    (sg_adj_march_kernel<F1>) + halo-sum kernel, against the C oracle; deterministic; slab of a sharded grid."""
    from gpu_helpers import make_grid, oracle_adjoint
    n_cp, deg, n_s, nout, ft, nurbs, distribution = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=0, nurbs=nurbs, seed=53, distribution=distribution)
    e = np.asfortranarray(rng.random(tuple(n_s) + (nout,)).astype(cp.dtype))
    g = torch.full_like(grid.control_points.obtain(), -7.0)
    monkeypatch.setenv("SG_ADJ_F2D", "2")                            # also where the cost heuristic would decline
    S.set_kernel_policy(2)
    try:
        S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g, allow_nurbs=nurbs)
        assert S.last_variant() == ("adjoint_passes_rational2d" if nurbs else "adjoint_fused2d")   # (rational: multi-pass)
        gref = oracle_adjoint(grid, e, None, w)
        assert rel_err(S.to_numpy(g), gref) <= _tol(ft)
        assert max_rel_err(S.to_numpy(g), gref) <= 10 * _tol(ft)
        g2 = torch.full_like(g, 9.0)
        S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g2, allow_nurbs=nurbs)
        assert torch.equal(g, g2)
        S.set_adjoint_plans(False)                                   # plain entry point: multi-pass pipeline
        g3 = torch.full_like(g, 2.0)
        S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g3, allow_nurbs=nurbs)
        assert S.last_variant().startswith("adjoint_passes")
        assert rel_err(S.to_numpy(g3), gref) <= _tol(ft)
    finally:
        S.set_adjoint_plans(True)
        S.set_kernel_policy(0)


def test_adjoint_plan_invalidated_by_dimension_rebuild(S):
    """A plan captures the inverse sample map; evaluate!(spline_dimension) / set_sample_indices! on new sample points
    must not leave a stale plan behind (the prepared-call key holds the dimensions' versions)."""
    from gpu_helpers import make_grid, oracle_adjoint
    grid, cp, w, rng = make_grid((36, 35, 20), (3, 3, 3), (256, 128, 64), 1, "Float64", mdo=0, seed=47)
    e = np.asfortranarray(rng.random((256, 128, 64, 1)))
    g = torch.zeros_like(grid.control_points.obtain())
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g)
    sd = grid.spline_dimensions[0]
    sd.sample_points.copy_(torch.sort(torch.rand_like(sd.sample_points)).values)
    S.set_sample_indices_(sd)
    S.evaluate_(sd)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g)
    assert rel_err(S.to_numpy(g), oracle_adjoint(grid, e)) <= 1e-12


MULTI_CASES = [
    # n_cp, degree, n_samples, nout, float type, mdo, derivative tuples
    ((16, 12), (3, 3), (300, 280), 3, "Float32", 1, [(0, 0), (1, 0), (0, 1)]),            # the optics caller: u, d1 u, d2 u
    ((16, 12), (3, 3), (300, 280), 2, "Float64", 2, [(0, 0), (2, 0), (0, 2), (1, 1)]),    # four tuples
    ((9, 14), (2, 2), (513, 77), 1, "Float64", 1, [(1, 0), (0, 1)]),                      # ragged n1, degree 2
    ((30, 8), (1, 1), (128, 400), 2, "Float32", 1, [(0, 0), (1, 1)]),                     # degree 1
    ((5, 8, 6), (3, 3, 3), (40, 36, 44), 1, "Float64", 2, [(0, 0, 0), (2, 0, 0), (0, 2, 0), (0, 0, 2)]),   # 3-D (PDE caller): one launch per tuple
    ((12, 10), (3, 4), (96, 73), 3, "Float64", 1, [(0, 0), (1, 0)]),                      # mixed degrees: one launch per tuple
]


@pytest.mark.parametrize("case", MULTI_CASES, ids=[f"{c[0]}-{c[1]}-{c[4]}-{len(c[6])}ders" for c in MULTI_CASES])
def test_evaluate_multi_vs_separate_oracle_calls(S, case):
    """sg_evaluate_multi (row f1): value + partial derivatives in one call == separate oracle evaluations."""
    from gpu_helpers import make_grid, oracle_evaluate
    n_cp, deg, n_s, nout, ft, mdo, ders = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=mdo, seed=61)
    evs = [torch.full_like(grid.eval, float("nan")) for _ in ders]
    S.set_kernel_policy(2)
    try:
        S.launch_count_reset()
        S.evaluate_multi_(grid, ders, evs)
        fused = len(n_cp) == 2 and len(set(deg)) == 1 and deg[0] <= 3
        assert S.launch_count() == (1 if fused else len(ders))
        if fused:
            assert S.last_variant() == "evaluate_multi2d"
    finally:
        S.set_kernel_policy(0)
    tol = _tol(ft)
    for der, ev in zip(ders, evs):
        ref = oracle_evaluate(grid, cp, der)
        assert rel_err(S.to_numpy(ev), ref) <= tol, der
        assert max_rel_err(S.to_numpy(ev), ref) <= 10 * tol, der
    with pytest.raises(S.SplineGridsError):
        S.evaluate_multi_(grid, [tuple(mdo + 1 for _ in n_cp)], evs[:1])


MIXED_CASES = [
    ((10, 10, 5), (2, 3, 2), (96, 80, 64), 4, "Float64", 1, False),          # the README degrees on a mid-size grid
    ((9, 12, 8), (3, 1, 2), (100, 70, 50), 2, "Float32", 1, False),
    ((14, 9), (1, 3), (700, 500), 3, "Float32", 1, False),
    ((14, 9), (3, 2), (640, 480), 2, "Float64", 0, True),                     # rational, mixed degrees
    ((6, 20, 7), (0, 3, 2), (64, 90, 60), 1, "Float64", 0, False),           # a degree-0 dimension
]


@pytest.mark.parametrize("case", MIXED_CASES, ids=[f"{c[0]}-{c[1]}-{c[4]}{'-nurbs' if c[6] else ''}" for c in MIXED_CASES])
def test_mixed_degree_forward_uses_padded_fast_path(S, case):
    """Mixed degrees: the tables are padded to the largest degree on the device and the uniform-degree march kernel runs
    (variant "..._mixed"); against the C oracle for the value and the first derivatives."""
    from gpu_helpers import make_grid, oracle_evaluate
    n_cp, deg, n_s, nout, ft, mdo, nurbs = case
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, mdo=mdo, nurbs=nurbs, seed=67)
    ders = [(0,) * len(n_cp)]
    if mdo and not nurbs:
        ders += [tuple(min(mdo, p) if d == k else 0 for d, p in enumerate(deg)) for k in range(len(n_cp))]
    for der in ders:
        grid.eval.fill_(float("nan"))
        S.evaluate_(grid, derivative_order=der)
        assert S.last_variant().endswith("_mixed"), S.last_variant()
        ref = oracle_evaluate(grid, cp, der, w)
        assert rel_err(S.to_numpy(grid.eval), ref) <= _tol(ft), der
        assert max_rel_err(S.to_numpy(grid.eval), ref) <= 10 * _tol(ft), der


def test_evaluate_with_raw_and_reshaped_arrays(S):
    """control_points / eval kwargs accept raw arrays and reshaped flat vectors (test/test_EnzymeExt.jl:24-28,
    ext/SplineGridsLinearMapsExt.jl:26-30)."""
    from gpu_helpers import make_grid, oracle_evaluate
    grid, cp, _, rng = make_grid((8, 9), (3, 2), (30, 40), 2, "Float64")
    flat_cp = S.to_device(cp.ravel(order="F"))
    flat_ev = torch.empty(30 * 40 * 2, dtype=torch.float64, device="cuda")
    S.evaluate_(grid, control_points=S.reshape_colmajor(flat_cp, (8, 9, 2)), eval=S.reshape_colmajor(flat_ev, (30, 40, 2)))
    ref = oracle_evaluate(grid, cp)
    assert rel_err(flat_ev.cpu().numpy(), ref.ravel(order="F")) <= 1e-12
    with pytest.raises(AssertionError):
        S.evaluate_(grid, eval=torch.empty(30, 41, 2, dtype=torch.float64, device="cuda"))


def test_nonmonotone_sample_points_use_scatter_path(S):
    """User-built SplineDimension with UNSORTED sample points (struct constructor, src/spline_dimension.jl:40-65):
    forward and adjoint must still match the oracle (adjoint falls back to the atomic scatter on device)."""
    from gpu_helpers import O
    rng = np.random.default_rng(5)
    dims, odims = [], []
    for c, p, n in zip((9, 7), (3, 2), (300, 280)):
        kv = S.KnotVector.clamped(c, p, float_type_="Float64")
        sp = rng.random(n)                                           # unsorted
        sd = S.SplineDimension.from_fields(p, 0, kv, S.to_device(sp), torch.zeros(n, dtype=torch.int32, device="cuda"),
                                           S.jl_zeros((n, p + 1, 1), torch.float64, "cuda"))
        S.build_(sd)
        idx = O.span_indices(sp, S.to_numpy(kv.knots_all), p)
        assert np.array_equal(idx, S.to_numpy(sd.sample_indices))
        odims.append((O.basis_tables(S.to_numpy(kv.knots_all), sp, idx, p, 0), idx))
        dims.append(sd)
    grid = S.SplineGrid(tuple(dims), 3)
    cp = np.asfortranarray(rng.random((9, 7, 3)))
    S.copyto_(grid.control_points, cp)
    S.evaluate_(grid)
    ref = O.evaluate([t for t, _ in odims], [i for _, i in odims], [3, 2], [0, 0], cp)
    assert rel_err(S.to_numpy(grid.eval), ref) <= 1e-12
    e = np.asfortranarray(rng.random(ref.shape))
    S.evaluate_adjoint_(grid, eval=S.to_device(e))
    gref = O.evaluate_adjoint([t for t, _ in odims], [i for _, i in odims], [3, 2], [0, 0], e, (9, 7, 3))
    assert rel_err(S.to_numpy(grid.control_points.obtain()), gref) <= 1e-12


@pytest.mark.parametrize("shape", [((9, 8, 40), (3, 2, 3), (40, 36, 96)), ((12, 9, 40), (2, 3, 3), (128, 40, 96))],
                         ids=["multipass", "march2"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_slab_sharded_grid_matches_full_grid(S, world, shape):
    """Multi-GPU data flow on one device: every rank's slab (sliced last-dimension arrays, replicated control
    points) evaluated separately; concatenated slabs == full evaluate!, summed partial gradients == full adjoint,
    and each partial gradient is non-zero only on the slab's support planes."""
    from gpu_helpers import oracle_adjoint, oracle_evaluate
    rng = np.random.default_rng(4)
    (n_cp, deg, n_s), nout = shape, 2
    gdims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip(n_cp, deg, n_s))
    full = S.SplineGrid(gdims, nout)
    cp = np.asfortranarray(rng.random(n_cp + (nout,)))
    e = np.asfortranarray(rng.random(n_s + (nout,)))
    S.copyto_(full.control_points, cp)
    ref_eval = oracle_evaluate(full, cp)
    ref_grad = oracle_adjoint(full, e)
    total = np.zeros_like(ref_grad)
    for policy in (0, 2):
        S.set_kernel_policy(policy)
        try:
            total[...] = 0
            for rank in range(world):
                sh = S.SlabShardedGrid(gdims, nout, rank, world)
                S.copyto_(sh.local.control_points, cp)
                sh.evaluate_()
                assert rel_err(S.to_numpy(sh.local.eval), ref_eval[:, :, sh.lo:sh.hi, :]) <= 1e-12
                g = torch.full_like(sh.local.control_points.obtain(), 5.0)
                S.evaluate_adjoint_(sh.local, eval=S.to_device(e[:, :, sh.lo:sh.hi, :]), control_points=g)
                gp = S.to_numpy(g)
                idx3 = S.to_numpy(gdims[2].sample_indices)
                k0, k1 = int(idx3[sh.lo]) - deg[2] - 1, int(idx3[sh.hi - 1])
                assert np.all(gp[:, :, :k0, :] == 0) and np.all(gp[:, :, k1:, :] == 0), S.last_variant()
                total += gp
            assert rel_err(total, ref_grad) <= 1e-12, S.last_variant()
        finally:
            S.set_kernel_policy(0)


@pytest.mark.parametrize("ft", ["Float32", "Float64"])
def test_peer_exchange_kernels_single_device(S, ft):
    """sg_exchange_push / sg_exchange_reduce with all "ranks" simulated on one device (the peer pointers simply
    point at buffers of the same GPU): the exchanged gradient equals the sum of the partial gradients."""
    import ctypes as C
    lib = S._lib.lib()
    suf = "f32" if ft == "Float32" else "f64"
    dt = torch.float32 if ft == "Float32" else torch.float64
    rng = np.random.default_rng(8)
    world, c1, c2, c_last, nout = 4, 10, 6, 23, 2
    plane = c1 * c2
    k0 = [0, 5, 11, 17]
    npl = [8, 9, 9, 6]                      # supports overlap by 3 planes
    max_planes = max(npl)
    grads = []
    for r in range(world):
        g = np.zeros((c1, c2, c_last, nout), dtype=np.float64, order="F")
        g[:, :, k0[r]:k0[r] + npl[r], :] = rng.random((c1, c2, npl[r], nout))
        grads.append(g)
    expected = sum(grads)
    stages = [torch.full((world * nout * max_planes * plane,), float("nan"), dtype=dt, device="cuda") for _ in range(world)]
    peer = (C.c_void_p * world)(*[t.data_ptr() for t in stages])
    dev = [S.to_device(g, dtype=dt) for g in grads]
    for r in range(world):
        S._lib.check(getattr(lib, "sg_exchange_push_" + suf)(
            S._lib.ptr(dev[r]), peer, C.c_int(world), C.c_int(r), C.c_int64(plane), C.c_int64(c_last), C.c_int(nout),
            C.c_int64(k0[r]), C.c_int64(npl[r]), C.c_int64(max_planes), None), "push")
    torch.cuda.synchronize()
    for r in range(world):
        S._lib.check(getattr(lib, "sg_exchange_reduce_" + suf)(
            S._lib.ptr(dev[r]), S._lib.ptr(stages[r]), C.c_int(world), S._lib.i64_array(k0), S._lib.i64_array(npl),
            C.c_int64(plane), C.c_int64(c_last), C.c_int(nout), C.c_int64(max_planes), None), "reduce")
        assert rel_err(S.to_numpy(dev[r]), expected) <= _tol(ft)
    for r in range(1, world):
        assert torch.equal(dev[r], dev[0])                          # rank-order summation: bit-identical on every rank


@pytest.mark.parametrize("fused_signal", [False, True], ids=["signal-kernel", "signal-in-reduce"])
def test_flag_synchronised_exchange_single_device(S, fused_signal):
    """sg_exchange_signal / sg_exchange_wait_reduce (the C ABI's own device-side barrier) with all "ranks" simulated on
    one device.  Every rank's wait needs every rank's signal, so the simulated ranks run on separate streams (signals
    first when they are separate kernels); two consecutive exchanges check the device-resident exchange counter."""
    import ctypes as C
    lib = S._lib.lib()
    rng = np.random.default_rng(18)
    world, c1, c2, c_last, nout = 3, 16, 6, 20, 2
    plane = c1 * c2
    k0, npl = [0, 5, 12], [8, 10, 8]
    max_planes = max(npl)
    flags = [torch.zeros(16, dtype=torch.int64, device="cuda") for _ in range(world)]
    syncs = [torch.zeros(8, dtype=torch.int64, device="cuda") for _ in range(world)]
    peer_flags = (C.c_void_p * world)(*[t.data_ptr() for t in flags])
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for exchange in range(2):
        grads = []
        for r in range(world):
            g = np.zeros((c1, c2, c_last, nout), order="F")
            g[:, :, k0[r]:k0[r] + npl[r], :] = rng.random((c1, c2, npl[r], nout))
            grads.append(g)
        expected = sum(grads)
        stages = [torch.full((world * nout * max_planes * plane,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(world)]
        peer = (C.c_void_p * world)(*[t.data_ptr() for t in stages])
        dev = [S.to_device(g) for g in grads]
        torch.cuda.synchronize()
        for r in range(world):
            st = C.c_void_p(streams[r].cuda_stream)
            S._lib.check(lib.sg_exchange_push_f64(S._lib.ptr(dev[r]), peer, C.c_int(world), C.c_int(r), C.c_int64(plane), C.c_int64(c_last),
                                                  C.c_int(nout), C.c_int64(k0[r]), C.c_int64(npl[r]), C.c_int64(max_planes), st), "push")
            if not fused_signal:
                S._lib.check(lib.sg_exchange_signal(peer_flags, C.c_int(world), C.c_int(r), S._lib.ptr(syncs[r]), st), "signal")
        for r in range(world):
            st = C.c_void_p(streams[r].cuda_stream)
            S._lib.check(lib.sg_exchange_wait_reduce_f64(
                S._lib.ptr(dev[r]), S._lib.ptr(stages[r]), S._lib.ptr(flags[r]), S._lib.ptr(syncs[r]),
                peer_flags if fused_signal else None, C.c_int(world), C.c_int(r), S._lib.i64_array(k0), S._lib.i64_array(npl),
                C.c_int64(plane), C.c_int64(c_last), C.c_int(nout), C.c_int64(max_planes), st), "wait_reduce")
        torch.cuda.synchronize()
        for r in range(world):
            assert rel_err(S.to_numpy(dev[r]), expected) <= 1e-14
            assert torch.equal(dev[r], dev[0])
            ep, to = C.c_ulonglong(0), C.c_int(0)
            S._lib.check(lib.sg_exchange_status(S._lib.ptr(syncs[r]), C.byref(ep), C.byref(to), None), "status")
            assert ep.value == exchange + 1 and to.value == 0
            assert flags[r][:world].tolist() == [exchange + 1] * world


@pytest.mark.parametrize("shape", [((12, 9, 40), (2, 3, 3), (128, 40, 96), "adjoint_march2"),
                                   ((9, 8, 40), (3, 2, 3), (40, 36, 96), "adjoint_passes")], ids=["fused", "fallback"])
@pytest.mark.parametrize("plans", [True, False], ids=["planned", "plain"])
@pytest.mark.parametrize("world", [2, 3])
def test_fused_gradient_push_single_device(S, world, shape, plans):
    """sg_evaluate_adjoint_planned with peer pointers / sg_evaluate_adjoint_push with every "rank" simulated on one device (peer pointers = local buffers): after all
    ranks' calls, every staging buffer holds exactly what adjoint + sg_exchange_push writes, and the reduce gives the
    full-grid gradient.  "fused": the double march's post kernel does the peer stores itself; "fallback": another
    pipeline runs, followed by the separate push kernel inside the same C call."""
    import ctypes as C
    from gpu_helpers import oracle_adjoint
    lib = S._lib.lib()
    rng = np.random.default_rng(14)
    n_cp, deg, n_s, variant = shape
    nout = 2
    S.set_adjoint_plans(plans)
    gdims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip(n_cp, deg, n_s))
    full = S.SplineGrid(gdims, nout)
    e = np.asfortranarray(rng.random(n_s + (nout,)))
    ref_grad = oracle_adjoint(full, e)
    idx3 = S.to_numpy(gdims[2].sample_indices)
    shards = [S.SlabShardedGrid(gdims, nout, r, world) for r in range(world)]
    k0 = [int(idx3[sh.lo]) - deg[2] - 1 for sh in shards]
    npl = [int(idx3[sh.hi - 1]) - k for sh, k in zip(shards, k0)]
    max_planes, plane = max(npl), n_cp[0] * n_cp[1]
    slot = world * nout * max_planes * plane

    def run(fused):
        stages = [torch.full((slot,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(world)]
        peer = (C.c_void_p * world)(*[t.data_ptr() for t in stages])
        grads = []
        for r, sh in enumerate(shards):
            g = torch.full_like(sh.local.control_points.obtain(), 3.0)
            ein = S.to_device(e[:, :, sh.lo:sh.hi, :])
            if fused:
                S.evaluate_adjoint_(sh.local, eval=ein, control_points=g,
                                    _push=(("t", r, fused), peer, world, r, k0[r], npl[r], max_planes, 0 if fused == "nolocal" else 1))
                assert S.last_variant().startswith(variant)
                if fused == "nolocal" and variant.startswith("adjoint_march2"):
                    assert float(g.min()) == 3.0 and float(g.max()) == 3.0     # the local array was not touched
            else:
                S.evaluate_adjoint_(sh.local, eval=ein, control_points=g)
                S._lib.check(lib.sg_exchange_push_f64(S._lib.ptr(g), peer, C.c_int(world), C.c_int(r), C.c_int64(plane),
                                                      C.c_int64(n_cp[2]), C.c_int(nout), C.c_int64(k0[r]), C.c_int64(npl[r]),
                                                      C.c_int64(max_planes), None), "push")
            grads.append(g)
        torch.cuda.synchronize()
        return stages, grads

    st_fused, grads = run(True)
    st_plain, _ = run(False)
    st_nolocal, _ = run("nolocal")
    for a, b, c in zip(st_fused, st_plain, st_nolocal):
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))   # same slots written, same bits
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(c, nan=-7.0))
    for r in range(world):
        S._lib.check(lib.sg_exchange_reduce_f64(S._lib.ptr(grads[r]), S._lib.ptr(st_fused[r]), C.c_int(world),
                                                S._lib.i64_array(k0), S._lib.i64_array(npl), C.c_int64(plane),
                                                C.c_int64(n_cp[2]), C.c_int(nout), C.c_int64(max_planes), None), "reduce")
        assert rel_err(S.to_numpy(grads[r]), ref_grad) <= 1e-12


@pytest.mark.parametrize("shape", [((12, 9, 40), (2, 3, 3), (128, 40, 96), "adjoint_march2"),
                                   ((9, 8, 40), (3, 2, 3), (40, 36, 96), "adjoint_passes")], ids=["fused", "fallback"])
@pytest.mark.parametrize("world", [2, 4])
def test_support_plane_exchange_single_device(S, world, shape):
    """sg_evaluate_adjoint_planned_support + sg_exchange_wait_reduce_support (the halo variant of the gradient exchange) with
    every "rank" simulated on one device, one stream per rank: each rank ends with the full-grid gradient on the control
    planes ITS slab reads, bit-identical to the replicated exchange there; the other planes are not written by the exchange;
    the fused pipeline stores into a neighbour's staging slot only the planes the two supports share."""
    import ctypes as C
    from gpu_helpers import oracle_adjoint
    lib = S._lib.lib()
    rng = np.random.default_rng(41)
    n_cp, deg, n_s, variant = shape
    nout = 2
    gdims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip(n_cp, deg, n_s))
    full = S.SplineGrid(gdims, nout)
    e = np.asfortranarray(rng.random(n_s + (nout,)))
    ref_grad = oracle_adjoint(full, e)
    idx3 = S.to_numpy(gdims[2].sample_indices)
    shards = [S.SlabShardedGrid(gdims, nout, r, world) for r in range(world)]
    k0 = [int(idx3[sh.lo]) - deg[2] - 1 for sh in shards]
    npl = [int(idx3[sh.hi - 1]) - k for sh, k in zip(shards, k0)]
    max_planes, plane = max(npl), n_cp[0] * n_cp[1]
    slot = world * nout * max_planes * plane
    k0s, nps = S._lib.i64_array(k0), S._lib.i64_array(npl)
    streams = [torch.cuda.Stream() for _ in range(world)]
    eins = [S.to_device(e[:, :, sh.lo:sh.hi, :]) for sh in shards]

    def run(support):
        stages = [torch.full((slot,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(world)]
        peer = (C.c_void_p * world)(*[t.data_ptr() for t in stages])
        flags = [torch.zeros(16, dtype=torch.int64, device="cuda") for _ in range(world)]
        syncs = [torch.zeros(8, dtype=torch.int64, device="cuda") for _ in range(world)]
        peer_flags = (C.c_void_p * world)(*[t.data_ptr() for t in flags])
        grads = [torch.full_like(sh.local.control_points.obtain(), 3.0) for sh in shards]
        torch.cuda.synchronize()
        for r, sh in enumerate(shards):
            with torch.cuda.stream(streams[r]):
                push = (("s", r, support), peer, world, r, k0[r], npl[r], max_planes, 0, 0) + (((k0s, nps),) if support else ())
                S.evaluate_adjoint_(sh.local, eval=eins[r], control_points=grads[r], _push=push)
                assert S.last_variant().startswith(variant)
        for r in range(world):
            fn = lib.sg_exchange_wait_reduce_support_f64 if support else lib.sg_exchange_wait_reduce_f64
            S._lib.check(fn(S._lib.ptr(grads[r]), S._lib.ptr(stages[r]), S._lib.ptr(flags[r]), S._lib.ptr(syncs[r]), peer_flags,
                            C.c_int(world), C.c_int(r), k0s, nps, C.c_int64(plane), C.c_int64(n_cp[2]), C.c_int(nout),
                            C.c_int64(max_planes), C.c_void_p(streams[r].cuda_stream)), "wait_reduce")
        torch.cuda.synchronize()
        for r in range(world):
            ep, to = C.c_ulonglong(0), C.c_int(0)
            S._lib.check(lib.sg_exchange_status(S._lib.ptr(syncs[r]), C.byref(ep), C.byref(to), None), "status")
            assert ep.value == 1 and to.value == 0
        return stages, grads

    st_rep, g_rep = run(False)
    st_sup, g_sup = run(True)
    fused = variant.startswith("adjoint_march2")
    for r in range(world):
        lo, hi = k0[r], k0[r] + npl[r]
        got = S.to_numpy(g_sup[r])
        assert rel_err(got[:, :, lo:hi, :], ref_grad[:, :, lo:hi, :]) <= 1e-12
        assert torch.equal(g_sup[r][:, :, lo:hi, :], g_rep[r][:, :, lo:hi, :])          # same bits as the replicated exchange
        rest = np.concatenate([got[:, :, :lo, :].ravel(), got[:, :, hi:, :].ravel()])
        assert rest.size > 0 and np.all(rest == (3.0 if fused else 0.0))                # not written by the exchange
        if fused:
            st = S.to_numpy(st_sup[r]).reshape((plane, max_planes, nout, world), order="F")
            for q in range(world):
                if q == r:
                    continue
                shared = [l for l in range(npl[q]) if lo <= k0[q] + l < hi]
                other = [l for l in range(max_planes) if l not in shared]
                assert not np.isnan(st[:, shared, :, q]).any() and np.isnan(st[:, other, :, q]).all()


def test_nurbs_adjoint_is_transpose_of_forward(S):
    """The NURBS adjoint has no reference behaviour (src/adjoint.jl:52-57): pin it as the exact transpose of our
    own forward map via <R p, e> = <p, R' e>, and as the plain adjoint when all weights are 1."""
    from gpu_helpers import make_grid, oracle_adjoint
    grid, cp, w, rng = make_grid((11, 9), (3, 2), (60, 50), 2, "Float64", nurbs=True, seed=2)
    e = np.asfortranarray(rng.random((60, 50, 2)))
    S.evaluate_(grid)
    fwd = S.to_numpy(grid.eval)
    g = torch.zeros_like(grid.control_points.obtain())
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g, allow_nurbs=True)
    lhs, rhs = float((fwd * e).sum()), float((cp * S.to_numpy(g)).sum())
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    with pytest.raises(TypeError):
        S.evaluate_adjoint_(grid, eval=S.to_device(e))              # the reference has no such method
    grid.weights.fill_(1)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g, allow_nurbs=True)
    assert rel_err(S.to_numpy(g), oracle_adjoint(grid, e)) <= 1e-12


def test_derivative_order_validation_messages(S, caplog):
    """test/test_spline_grid.jl:55-83 and test/test_nurbs_grid.jl:11-23."""
    dims = tuple(S.SplineDimension(c, p, n, max_derivative_order=1) for c, p, n in zip((5, 6), (3, 2), (7, 9)))
    grid = S.SplineGrid(dims, 2)
    with caplog.at_level(logging.ERROR, logger="splinegrids_b200"):
        with pytest.raises(S.SplineGridsError, match=r"Invalid derivative order\(s\) supplied. If you want to evaluate "
                                                     r"\(higher order\) derivatives, specify this at construction as "
                                                     r"SplineDimension\(...; max_derivative_order\)."):
            S.evaluate_(grid, derivative_order=(2, 1))
    assert len(caplog.records) == 1
    assert caplog.records[0].getMessage() == "The maximum derivative order available for spline dimension 1 is 1, got 2."
    nurbs = S.NURBSGrid(S.SplineDimension(4, 2, 100, max_derivative_order=1), 1)
    with pytest.raises(S.SplineGridsError, match="Computing derivatives of NURBS is currently not supported."):
        S.evaluate_(nurbs, derivative_order=(1,))


# ---------------------------------------------------------------------------------------------
# K5-K8: refinement matrices and the hierarchy
# ---------------------------------------------------------------------------------------------


def test_refinement_matrix_unit_tests(S):
    """test/test_refinement_matrix.jl:16-75."""
    M = S.rmeye(100)
    assert M[25, 25] == 1 and M[25, 26] == 0
    with pytest.raises(S.SplineGridsError, match=r"Index \(101, 1\) out of bounds for refinement matrix of size \(100, 100\)."):
        M[101, 1]
    M2 = M @ M
    assert M2 == M and M2.shape == (100, 100) and len(M2) == 100 ** 2
    rng = np.random.default_rng(1)
    A, B = S.refinement_matrix_from_dense(np.diag(rng.random(100))), S.refinement_matrix_from_dense(np.diag(rng.random(100)))
    assert np.array_equal((A @ B).nzval_host, A.nzval_host * B.nzval_host)

    def brand(m, n, lo, up):
        a = rng.random((m, n))
        i, j = np.indices((m, n))
        a[(i - j > lo) | (j - i > up)] = 0
        return a

    Ad, Bd = brand(100, 100, 4, 3), brand(100, 200, 3, 5)
    Cm = S.refinement_matrix_from_dense(Ad) @ S.refinement_matrix_from_dense(Bd)
    assert julia_isapprox(Cm.collect(), Ad @ Bd)
    Mz = np.zeros((2, 50))
    Mz[0, :20] = 1
    Mz[1, 38:] = 1
    with pytest.raises(S.SplineGridsError, match=r"Invalid rows: \[2\]\."):
        S.refinement_matrix_from_dense(Mz)
    Mz = np.zeros((3, 50))
    Mz[0, :25] = 1
    Mz[2, 25:] = 1
    with pytest.raises(S.SplineGridsError, match=r"Invalid rows: \[2, 3\]\."):
        S.refinement_matrix_from_dense(Mz)


@pytest.mark.parametrize("ft", ["Float32", "Float64"])
def test_mult_and_mult_adjoint_vs_oracle(S, ft):
    """K5 / K6 with random banded matrices on 1, 2 and 3 refined dimensions of a 4-d array."""
    from gpu_helpers import O, OC
    rng = np.random.default_rng(9)
    npdt = np.float32 if ft == "Float32" else np.float64

    def banded(m, n):
        """Random valid refinement-matrix structure (m >= n): monotone, touching row windows."""
        a = np.zeros((m, n), dtype=npdt)
        lo_prev, hi_prev = 0, 1
        for i in range(m):
            c = int(round(i * (n - 1) / max(m - 1, 1)))
            lo = 0 if i == 0 else min(max(lo_prev, c - int(rng.integers(0, 3))), hi_prev)
            hi = min(n, max(hi_prev, c + 1 + int(rng.integers(0, 3))))
            a[i, lo:hi] = rng.random(hi - lo) + 0.1
            lo_prev, hi_prev = lo, hi
        return a

    sizeB = (6, 7, 5, 3)
    B = np.asfortranarray(rng.random(sizeB).astype(npdt))
    for dims_ref, ms in (((2,), (11,)), ((1, 3), (10, 9)), ((3, 1, 2), (8, 9, 12))):
        As_d = [banded(m, sizeB[d - 1]) for d, m in zip(dims_ref, ms)]
        As = [S.refinement_matrix_from_dense(a) for a in As_d]
        As_o = [O.refmat_from_dense(a) for a in As_d]
        sizeY = list(sizeB)
        for d, m in zip(dims_ref, ms):
            sizeY[d - 1] = m
        Y = S.jl_zeros(sizeY, As[0].dtype, "cuda")
        S.mult_(Y, As, S.to_device(B), dims_ref)
        Yo = np.zeros(sizeY, dtype=npdt, order="F")
        OC.mult(Yo, As_o, B, list(dims_ref))
        assert rel_err(S.to_numpy(Y), Yo) <= _tol(ft)
        Yin = np.asfortranarray(rng.random(sizeY).astype(npdt))
        Bout = S.jl_empty(sizeB, As[0].dtype, "cuda")
        Bout.fill_(3.0)                                              # must be overwritten (B .= 0, adjoint.jl:135)
        S.mult_adjoint_(Bout, As, S.to_device(Yin), dims_ref)
        Bo = np.zeros(sizeB, dtype=npdt, order="F")
        OC.mult_adjoint(Bo, As_o, Yin, list(dims_ref))
        assert rel_err(S.to_numpy(Bout), Bo) <= _tol(ft)
    with pytest.raises(S.SplineGridsError, match="Size of refinement matrix does not match `B` and `Y` along refinement dimension 2."):
        S.mult_(S.jl_zeros((6, 12, 5, 3), As[0].dtype, "cuda"), [S.rmeye(7, float_type_=ft)], S.to_device(B), (2,))


def _g7_grid(S):
    rng = np.random.default_rng(1)
    dims = tuple(S.SplineDimension(c, p, n, distribution="random", rng=rng)
                 for c, p, n in zip((5, 8, 6), (4, 2, 3), (15, 20, 25)))
    grid = S.SplineGrid(dims, 2)
    S.copyto_(grid.control_points, rng.random((5, 8, 6, 2)).astype(np.float32))
    S.evaluate_(grid)
    return grid


def test_g7_knot_insertion_and_refine_preserve_geometry(S):
    """test/test_refinement.jl:29-58."""
    grid0 = _g7_grid(S)
    ev0 = S.to_numpy(grid0.eval).copy()
    grid = grid0
    for dr in (1, 2, 3):
        grid, R = S.insert_knot(grid, dr, 0.25)
        grid = grid.replace(eval=torch.zeros_like(grid0.eval))
        S.evaluate_(grid)
        assert julia_isapprox(S.to_numpy(grid.eval), ev0)
        shp = grid.control_points.shape
        assert R.shape == (shp[dr - 1], shp[dr - 1] - 1)
    grid = grid0
    old = grid0.control_points.shape
    for dr in (1, 2, 3):
        grid, R = S.refine(grid, dr)
        grid = grid.replace(eval=torch.zeros_like(grid0.eval))
        S.evaluate_(grid)
        assert julia_isapprox(S.to_numpy(grid.eval), ev0)
        assert R.shape == (grid.control_points.shape[dr - 1], old[dr - 1])


def test_g8_local_refinement_counts(S):
    """test/test_local_refinement.jl:11-46 + the second level of docs/src/theory_local_refinement.md:19-63."""
    dims = tuple(S.SplineDimension(6, 2, 500) for _ in range(2))
    grid = S.SplineGrid(dims, 3)
    assert S.get_n_control_points(grid) == 36
    assert S.evaluate_(grid.control_points) is None
    grid = S.add_default_local_refinement(grid)
    cps = grid.control_points
    assert isinstance(cps, S.LocallyRefinedControlPoints)
    assert "LocallyRefinedControlPoints for final grid of size (10, 10) in R^3 (torch.float32). Local refinements:" in repr(cps)
    assert S.get_n_control_points(grid) == 36
    before = S.to_numpy(S.obtain(cps)).copy()
    S.activate_local_control_point_range_(grid, range(1, 5), range(1, 7))
    S.activate_local_control_point_range_(grid, range(1, 7), range(1, 3))
    S.activate_local_control_point_range_(grid, range(9, 11), range(7, 11))
    S.deactivate_overwritten_control_points_(grid.control_points)
    S.evaluate_(grid.control_points)
    assert S.get_n_control_points(grid) == 63
    assert julia_isapprox(S.to_numpy(S.obtain(cps)), before)
    grid = S.add_default_local_refinement(grid)
    assert S.obtain(grid.control_points).shape == (18, 18, 3)
    S.activate_local_control_point_range_(grid, range(5, 13), range(1, 5))
    S.activate_local_control_point_range_(grid, range(7, 9), range(5, 7))
    S.deactivate_overwritten_control_points_(grid.control_points)
    fine = S.to_numpy(S.obtain(grid.control_points)).copy()
    S.evaluate_(grid.control_points)
    assert julia_isapprox(S.to_numpy(S.obtain(grid.control_points)), fine)
    # same hierarchy through the oracle: identical active sets, finest control points and evaluation
    from gpu_helpers import O, oracle_dim
    odims = [O.make_dimension(6, 2, 500) for _ in range(2)]
    odims, olr = O.add_default_local_refinement(odims, O.unit_cp_grid((6, 6, 3), np.float32))
    O.activate_local_control_point_range(olr, (1, 4), (1, 6))
    O.activate_local_control_point_range(olr, (1, 6), (1, 2))
    O.activate_local_control_point_range(olr, (9, 10), (7, 10))
    O.deactivate_overwritten_control_points(olr)
    odims, olr = O.add_default_local_refinement(odims, olr)
    O.activate_local_control_point_range(olr, (5, 12), (1, 4))
    O.activate_local_control_point_range(olr, (7, 8), (5, 6))
    O.deactivate_overwritten_control_points(olr)
    for lr_g, lr_o in zip(grid.control_points.local_refinements, olr.local_refinements):
        assert np.array_equal(S.to_numpy(lr_g.refinement_indices).reshape(-1, 2), lr_o.refinement_indices)
    rng = np.random.default_rng(42)
    vals = rng.random((S.get_n_control_points(grid), 3)).astype(np.float32)
    S.copyto_(grid.control_points, vals)
    S.evaluate_(grid.control_points)
    olr.set_values(vals)
    O.lrcp_evaluate(olr, fast=False)
    assert rel_err(S.to_numpy(S.obtain(grid.control_points)), olr.control_points_refined[-1]) <= 1e-5
    S.evaluate_(grid)
    tabs = [oracle_dim(sd, O) for sd in grid.spline_dimensions]
    ref = O.evaluate([t for t, _ in tabs], [i for _, i in tabs], [2, 2], [0, 0],
                     np.asfortranarray(olr.control_points_refined[-1]))
    assert rel_err(S.to_numpy(grid.eval), ref) <= 1e-5


def test_g2_error_informed_refinement_golden_indices(S):
    """test/test_local_refinement.jl:48-67: exact 18x2 Int32 index set through the device adjoint."""
    _, expected = load_g2()
    dims = tuple(S.SplineDimension(6, 2, 50) for _ in range(2))
    grid = S.add_default_local_refinement(S.SplineGrid(dims, 3))
    err = torch.zeros_like(grid.eval)
    err[19:40, 9:30, 1] = 1
    for policy in (1, 0, 2):
        g = S.add_default_local_refinement(S.SplineGrid(dims, 3))
        S.set_kernel_policy(policy)
        try:
            S.error_informed_local_refinement_(g, err)
        finally:
            S.set_kernel_policy(0)
        got = S.to_numpy(g.control_points.local_refinements[-1].refinement_indices)
        assert got.dtype == np.int32 and np.array_equal(got, expected)


def _lsqr(A, b):
    """scipy lsqr over the device linear map (host vectors in, device kernels inside)."""
    from scipy.sparse.linalg import LinearOperator, lsqr
    import gpu_helpers
    S_ = gpu_helpers.sg()

    def mv(x):
        return A.matvec(S_.to_device(np.asarray(x, dtype=np.float64))).cpu().numpy()

    def rmv(y):
        return A.rmatvec(S_.to_device(np.asarray(y, dtype=np.float64))).cpu().numpy()

    return lsqr(LinearOperator(A.shape, matvec=mv, rmatvec=rmv, dtype=np.float64), b, atol=1e-12, btol=1e-12,
                iter_lim=2000)[0]


def test_g9_least_squares_fitting(S):
    """test/test_LinearMapsExt.jl:14-33."""
    rng = np.random.default_rng(1)
    dims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip((12, 10), (3, 4), (26, 73)))
    grid = S.SplineGrid(dims, 3)
    cp = rng.random((12, 10, 3))
    S.copyto_(grid.control_points, cp)
    S.evaluate_(grid)
    A = S.SplineGridLinearMap(grid)
    fit = _lsqr(A, S.to_numpy(grid.eval).ravel(order="F").copy())
    assert np.allclose(fit, cp.ravel(order="F"), rtol=1e-5)


def test_g9_locally_refined_least_squares_fitting(S):
    """test/test_LinearMapsExt.jl:35-62 (K3-K8 mutually adjoint and consistent)."""
    rng = np.random.default_rng(1)
    dims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip((12, 10), (3, 4), (26, 73)))
    grid = S.add_default_local_refinement(S.SplineGrid(dims, 3))
    S.activate_local_control_point_range_(grid, range(1, 5), range(1, 7))
    S.activate_local_control_point_range_(grid, range(1, 7), range(1, 3))
    S.activate_local_control_point_range_(grid, range(9, 11), range(7, 11))
    S.deactivate_overwritten_control_points_(grid.control_points)
    n_cp = S.get_n_control_points(grid)
    vals = rng.random((n_cp, 3))
    S.copyto_(grid.control_points, vals)
    S.evaluate_(grid.control_points)
    S.evaluate_(grid)
    target = S.to_numpy(grid.eval).ravel(order="F").copy()
    A = S.SplineGridLinearMap(grid)
    x, y = rng.random(A.shape[1]), rng.random(A.shape[0])
    lhs = float(A.matvec(S.to_device(x)).cpu().numpy() @ y)
    rhs = float(x @ A.rmatvec(S.to_device(y)).cpu().numpy())
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    fit = _lsqr(A, target)
    assert np.allclose(fit, vals.ravel(order="F"), rtol=1e-5)


# ---------------------------------------------------------------------------------------------
# K11 / K12 set-up helpers through the C ABI
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("ft", ["Float32", "Float64"])
def test_device_refinement_setup_kernels_vs_oracle(S, ft):
    """Scope row f3: K13 (Boehm matrix), K16 + scan + K15 (sparse product), K14 (validation), K17 (collect) on the device
    are BIT-identical to the oracle's restatement of the reference kernels; K18 + unique + Flag pipeline + error-informed
    flags against numpy restatements."""
    from gpu_helpers import O
    from splinegrids_jl_b200 import device_setup as D
    npdt = np.float32 if ft == "Float32" else np.float64
    tdt = torch.float32 if ft == "Float32" else torch.float64
    rng = np.random.default_rng(71)
    # K13 + product chain: insert four knots one after the other into a random clamped knot vector of degree 3
    p, nb = 3, 9
    kv, mu, ka = O.clamped_knot_vector(nb, p, npdt, distribution="random", rng=rng)
    R_dev, R_ref = None, None
    for knot_new in (0.31, 0.62, 0.05, 0.93):
        kvn, mun, k = O.insert_knot(kv, mu, npdt(knot_new))
        ref = O.boehm_matrix(ka, p, k, npdt(knot_new))
        dev = S.boehm_refinement_matrix(S.to_device(ka), p, k, npdt(knot_new))
        assert np.array_equal(dev.row_pointer_host, ref.row_pointer) and np.array_equal(dev.column_start_host, ref.column_start)
        assert np.array_equal(dev.nzval_host, ref.nzval)                    # bit-identical (no contraction)
        R_dev = dev if R_dev is None else dev @ R_dev
        R_ref = ref if R_ref is None else O.refmat_matmul(ref, R_ref)
        assert np.array_equal(R_dev.row_pointer_host, R_ref.row_pointer) and np.array_equal(R_dev.column_start_host, R_ref.column_start)
        assert np.array_equal(R_dev.nzval_host, R_ref.nzval)
        assert np.array_equal(R_dev.collect(), R_ref.dense())               # K17
        kv, mu, ka = kvn, mun, O.expand_knot_vector(kvn, mun)
    # K14: rows 2 and 3 invalid (gap after row 1, then a row ending before the previous one)
    bad = np.zeros((4, 5), dtype=npdt)
    bad[0, 0] = 1; bad[1, 3:5] = 1; bad[2, 3] = 1; bad[3, 3:5] = 1
    with pytest.raises(S.SplineGridsError, match=r"Invalid rows: \[2, 3\]"):
        S.refinement_matrix_from_dense(bad)
    # unique(vcat(old, new); dims = 1) + K18
    n_cp, nout = (7, 6), 2
    cp = S.to_device(np.asfortranarray(rng.random(n_cp + (nout,)).astype(npdt)))
    old = np.array([[1, 1], [3, 2], [7, 6]], dtype=np.int32)
    new = np.array([[3, 2], [2, 2], [2, 2], [5, 6], [1, 1], [4, 4]], dtype=np.int32)
    keep = D.unique_new_rows(S.to_device(old), S.to_device(new), n_cp)
    assert keep.cpu().tolist() == [0, 1, 0, 1, 0, 1]
    rows = D.compact_flags(keep)
    assert rows.cpu().tolist() == [1, 3, 5]
    kept = D.gather_rows(S.to_device(new), rows)
    assert np.array_equal(S.to_numpy(kept), new[[1, 3, 5]])
    idx_all = S.to_device(np.concatenate([old, new[[1, 3, 5]]]))
    vals_old = S.to_device(np.asfortranarray(rng.random((3, nout)).astype(npdt)))
    vals = S.to_numpy(D.refinement_values_new(vals_old, cp, idx_all))
    cpn = S.to_numpy(cp)
    assert np.array_equal(vals[:3], S.to_numpy(vals_old))
    assert np.array_equal(vals[3:], np.stack([cpn[1, 1], cpn[4, 5], cpn[3, 3]]))
    # error-informed flags: grid error = sum over outputs, threshold = factor * mean
    err = np.asfortranarray(rng.random((40, 30, 3)).astype(npdt))
    flags = D.error_flags(S.to_device(err), 1.25).cpu().numpy().reshape((40, 30), order="F").astype(bool)
    ge = err.sum(axis=2)
    thr = npdt(1.25) * ge.sum() / ge.size
    expect = ge > thr
    assert np.count_nonzero(flags != expect) <= 2                           # ties within rounding of the mean only
    hit = D.compact_flags(S.to_device(np.ascontiguousarray(expect.ravel(order="F").astype(np.uint8))))
    idx = S.to_numpy(D.indices_from_linear(hit, (40, 30)))
    ref_idx = np.argwhere(expect.T)[:, ::-1] + 1                            # column-major findall order, 1-based
    assert np.array_equal(idx, ref_idx.astype(np.int32))
    # scan / compaction across several scan blocks
    big = (rng.random(5000) > 0.7).astype(np.uint8)
    pos = D.compact_flags(S.to_device(big))
    assert np.array_equal(pos.cpu().numpy(), np.flatnonzero(big).astype(np.int32))
    assert np.array_equal(D.compact_flags(S.to_device(big), invert=True).cpu().numpy(), np.flatnonzero(big == 0).astype(np.int32))


def test_ad_rules_match_reference_enzyme_test(S):
    """Scope row f2.  test/test_EnzymeExt.jl:13-49 restated: loss(control_points_flat) = sum(evaluate!(grid; control_points));
    reverse mode gives a non-zero gradient -- here checked exactly: it must equal the oracle's adjoint applied to ones,
    the JVP must equal evaluate! on the tangent, and <J t, y> = <t, J' y>; make_zero! zeroes eval and the tables."""
    from gpu_helpers import make_grid, oracle_adjoint, oracle_evaluate
    grid, cp, w, rng = make_grid((10, 10), (2, 2), (50, 50), 2, "Float32", seed=1)
    flat = torch.tensor(cp.ravel(order="F"), device="cuda", requires_grad=True)
    loss = S.evaluate_autograd(grid, flat).sum()
    loss.backward()
    assert flat.grad is not None and bool((flat.grad != 0).any())              # the reference's assertion (:53)
    gref = oracle_adjoint(grid, np.ones((50, 50, 2), dtype=np.float32, order="F"))
    assert rel_err(flat.grad.cpu().numpy(), gref.ravel(order="F")) <= 1e-5
    # a derivative order and Float64: gradient of a weighted sum, forward rule, dot-product identity
    grid, cp, w, rng = make_grid((9, 7, 6), (3, 2, 2), (20, 18, 16), 2, "Float64", mdo=1, seed=2)
    der = (0, 1, 0)
    flat = torch.tensor(cp.ravel(order="F"), device="cuda", requires_grad=True)
    y = np.asfortranarray(rng.random((20, 18, 16, 2)))
    (S.evaluate_autograd(grid, flat, der) * S.to_device(y)).sum().backward()
    assert rel_err(flat.grad.cpu().numpy(), oracle_adjoint(grid, y, der).ravel(order="F")) <= 1e-12
    t = np.asfortranarray(rng.random(cp.shape))
    import torch.autograd.forward_ad as fwAD
    with fwAD.dual_level():
        dual = fwAD.make_dual(flat.detach(), torch.tensor(t.ravel(order="F"), device="cuda"))
        jt = fwAD.unpack_dual(S.evaluate_autograd(grid, dual, der)).tangent
    ref_jt = oracle_evaluate(grid, t, der)
    assert rel_err(S.to_numpy(jt), ref_jt) <= 1e-12
    lhs = float((S.to_numpy(jt) * y).sum())
    rhs = float((t * oracle_adjoint(grid, y, der)).sum())
    assert abs(lhs - rhs) <= 1e-11 * abs(lhs)
    S.make_zero_(grid)
    assert float(grid.eval.abs().max()) == 0.0 and all(float(sd.eval.abs().max()) == 0.0 for sd in grid.spline_dimensions)


def test_insert_and_collect_indices_kernels(S):
    """K11 `insert_kernel` (src/util_kernels.jl:69-79) and K12 `collect_indices_kernel` (:81-88)."""
    import ctypes as C
    lib = S._lib.lib()
    rng = np.random.default_rng(0)
    for npdt, tdt, fn, ctype in ((np.float32, torch.float32, lib.sg_insert_f32, C.c_float),
                                 (np.float64, torch.float64, lib.sg_insert_f64, C.c_double),
                                 (np.int32, torch.int32, lib.sg_insert_i32, C.c_int32)):
        v = (rng.random(37) * 100).astype(npdt)
        for pos in (1, 19, 38):                                  # 1-based insert position, incl. both ends
            out = torch.empty(38, dtype=tdt, device="cuda")
            vin = torch.from_numpy(v).cuda()
            S._lib.check(fn(S._lib.ptr(out), S._lib.ptr(vin), C.c_int64(37), C.c_int64(pos), ctype(7), None), "sg_insert")
            assert np.array_equal(out.cpu().numpy(), np.insert(v, pos - 1, npdt(7)))
    cart = rng.integers(1, 50, size=(123, 3)).astype(np.int64)   # 123 CartesianIndex{3}, AoS
    idx = torch.empty((3, 123), dtype=torch.int32, device="cuda")   # (123, 3) column-major
    S._lib.check(lib.sg_collect_indices_i32(S._lib.ptr(idx), S._lib.ptr(torch.from_numpy(cart).cuda()), C.c_int64(123),
                                            C.c_int(3), None), "sg_collect_indices")
    assert np.array_equal(idx.cpu().numpy().T, cart.astype(np.int32))


# ---------------------------------------------------------------------------------------------
# C ABI behaviour
# ---------------------------------------------------------------------------------------------


def test_c_abi_error_codes(S):
    import ctypes as C
    lib = S._lib.lib()
    assert lib.sg_span_indices_f32(None, None, C.c_int64(4), None, C.c_int64(8), C.c_int(2), None) == -1
    assert lib.sg_status_string(-1) == b"SG_ERR_INVALID_ARGUMENT"
    t = torch.zeros(8, device="cuda")
    idx = torch.zeros(8, dtype=torch.int32, device="cuda")
    # degree above SG_MAX_DEGREE
    assert lib.sg_basis_tables_f32(S._lib.ptr(t), S._lib.ptr(t), C.c_int64(40), S._lib.ptr(t), S._lib.ptr(idx),
                                   C.c_int64(8), C.c_int(16), C.c_int(0), None) == -2
    # misaligned / short workspace
    grid = S.SplineGrid(tuple(S.SplineDimension(16, 3, 200) for _ in range(2)), 3)
    from importlib import import_module
    sgmod = import_module("splinegrids_jl_b200.spline_grid")
    args = sgmod._grid_call_args(grid, (0, 0))
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")
    rc = lib.sg_evaluate_adjoint_f32(S._lib.ptr(grid.control_points.obtain()), *args, S._lib.ptr(grid.eval), None,
                                     S._lib.ptr(ws), C.c_size_t(64), None)
    assert rc == -3
    before = S.launch_count()
    S.evaluate_(grid)
    assert S.launch_count() == before + 1


def test_captured_calls_replay_matches_eager(S):
    """evaluate! + evaluate_adjoint! captured in a CUDA graph (graphs.py) and replayed on NEW array contents give
    exactly what the eager calls give: the entry points only enqueue kernels/memsets on the given stream."""
    from gpu_helpers import make_grid
    grid, cp, w, rng = make_grid((24, 20, 18), (3, 3, 3), (256, 70, 66), 2, "Float64", mdo=0, seed=5)
    e = S.to_device(np.asfortranarray(rng.random((256, 70, 66, 2))))
    g = torch.zeros_like(grid.control_points.obtain())

    def step():
        S.evaluate_(grid)
        S.evaluate_adjoint_(grid, eval=e, control_points=g)

    cap = S.CapturedCalls(step, unroll=2)
    assert cap.kernel_launches >= 4
    # new inputs, same arrays
    grid.control_points.obtain().copy_(torch.rand_like(grid.control_points.obtain()))
    e.copy_(torch.rand_like(e))
    grid.eval.fill_(-1.0)
    g.fill_(-1.0)
    cap.replay()
    torch.cuda.synchronize()
    ev_graph, g_graph = grid.eval.clone(), g.clone()
    grid.eval.fill_(-2.0)
    g.fill_(-2.0)
    step()
    torch.cuda.synchronize()
    assert torch.equal(ev_graph, grid.eval)
    assert torch.equal(g_graph, g)
