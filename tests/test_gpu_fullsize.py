"""GPU parity at BASELINE.json's FULL sizes (configs C2-C5) through size-independent properties plus oracle
comparisons on sub-blocks the CPU oracle finishes in seconds:

* partition of unity (control points == const  =>  eval == const; derivatives == 0),
* linearity, adjoint identity <A x, y> == <x, A' y>,
* the C oracle on a slab / a strided sub-grid of the same arrays,
* refinement preserves geometry and forward/adjoint consistency of the hierarchy (C5, scaled instance).
"""
import numpy as np
import pytest
from helpers import rel_err

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def S():
    from gpu_helpers import sg
    S = sg()
    yield S
    torch.cuda.empty_cache()


def _oracle_subgrid(S, grid, cp_np, sel, der=None, weights_np=None):
    """C-oracle evaluation of the sub-grid given by per-dimension index arrays `sel` (same tables/indices)."""
    from gpu_helpers import OC
    tabs, idxs = [], []
    for sd, s in zip(grid.spline_dimensions, sel):
        tabs.append(np.asfortranarray(S.to_numpy(sd.eval)[s]))
        idxs.append(np.ascontiguousarray(S.to_numpy(sd.sample_indices)[s]))
    der = [0] * grid.Nin if der is None else list(der)
    return OC.evaluate(tabs, idxs, [sd.degree for sd in grid.spline_dimensions], der, np.asfortranarray(cp_np), weights_np)


def _dot(a, b):
    return float((a.double() * b.double()).sum())


def test_c3_full_size_3d_cubic_float64(S):
    """C3: control 128^3, samples 512^3, Nout 1, Float64."""
    dims = tuple(S.SplineDimension(128, 3, 512, float_type="Float64") for _ in range(3))
    grid = S.SplineGrid(dims, 1)
    cp = grid.control_points.obtain()
    # partition of unity
    cp.fill_(2.5)
    S.evaluate_(grid)
    assert S.last_variant().startswith("evaluate_march3d")
    assert float((grid.eval - 2.5).abs().max()) <= 2.5 * 1e-12
    # oracle on a slab of 3 planes and on a strided sub-grid
    g = torch.Generator(device="cuda").manual_seed(1)
    cp.copy_(torch.rand(cp.shape, dtype=cp.dtype, device="cuda", generator=g))
    S.evaluate_(grid)
    cp_np = S.to_numpy(cp)
    ev = grid.eval
    sel = [np.arange(512), np.arange(512), np.array([0, 255, 511])]
    ref = _oracle_subgrid(S, grid, cp_np, sel)
    got = S.to_numpy(ev[:, :, torch.tensor(sel[2], device="cuda"), :])
    assert rel_err(got, ref) <= 1e-12
    sel = [np.arange(0, 512, 7), np.arange(3, 512, 11), np.arange(1, 512, 5)]
    ref = _oracle_subgrid(S, grid, cp_np, sel)
    ix = [torch.tensor(s, device="cuda") for s in sel]
    got = S.to_numpy(ev[ix[0]][:, ix[1]][:, :, ix[2]])
    assert rel_err(got, ref) <= 1e-12
    # adjoint identity at full size (same arrays, float64)
    y = torch.empty_like(ev)
    y.copy_(torch.rand(y.shape, dtype=y.dtype, device="cuda", generator=g))
    grad = torch.zeros_like(cp)
    S.evaluate_adjoint_(grid, eval=y, control_points=grad)
    assert S.last_variant().startswith("adjoint_march2")
    lhs, rhs = _dot(ev, y), _dot(cp, grad)
    assert abs(lhs - rhs) <= 1e-11 * abs(lhs)
    # ... and the multi-pass pipeline gives the same gradient
    import os
    for env, variant in (({"SG_ADJ_MARCH2": "0"}, "adjoint_passes"),):
        grad2 = torch.zeros_like(cp)
        os.environ.update(env)
        try:
            S.evaluate_adjoint_(grid, eval=y, control_points=grad2)
            assert S.last_variant() == variant
        finally:
            for k in env:
                del os.environ[k]
        assert rel_err(S.to_numpy(grad2), S.to_numpy(grad)) <= 1e-12
    # the C ORACLE's adjoint (the reference's algorithm, src/adjoint.jl:11-39) at full size: the input is zero outside
    # four sample planes (both ends and the two middle planes, which share chunk / tile boundaries), so the oracle only
    # has to scatter a 512 x 512 x 4 slab (seconds) and the FULL 128^3 gradient is compared
    from gpu_helpers import OC
    planes = np.array([0, 255, 256, 511])
    y_np = np.zeros((512, 512, 4, 1), order="F")
    y_np[...] = np.random.default_rng(7).random(y_np.shape)
    y.zero_()
    y[:, :, torch.tensor(planes, device="cuda"), :] = S.to_device(y_np)
    S.evaluate_adjoint_(grid, eval=y, control_points=grad)
    assert S.last_variant().startswith("adjoint_march2")
    tabs = [np.asfortranarray(S.to_numpy(sd.eval)) for sd in dims]
    idxs = [np.ascontiguousarray(S.to_numpy(sd.sample_indices)) for sd in dims]
    tabs[2], idxs[2] = np.asfortranarray(tabs[2][planes]), np.ascontiguousarray(idxs[2][planes])
    gref = OC.evaluate_adjoint(tabs, idxs, [3, 3, 3], [0, 0, 0], y_np, (128, 128, 128, 1))
    assert rel_err(S.to_numpy(grad), gref) <= 1e-12
    from helpers import max_rel_err
    assert max_rel_err(S.to_numpy(grad), gref) <= 1e-11
    # adjoint of a one-hot plane pattern equals basis sums: e == 1 => grad[i] = prod_d sum_j B_d[j, i]
    y.fill_(1.0)
    S.evaluate_adjoint_(grid, eval=y, control_points=grad)
    col = []
    for sd in dims:
        dense = S.to_numpy(S.decompress(sd)).sum(axis=0)
        col.append(dense)
    expect = col[0][:, None, None] * col[1][None, :, None] * col[2][None, None, :]
    assert rel_err(S.to_numpy(grad)[..., 0], expect) <= 1e-12
    del y, grad, grid
    torch.cuda.empty_cache()


def test_c2_full_size_2d_float32_with_first_derivatives(S):
    """C2: control 64x64, samples 4096x4096, Nout 3, Float32, value + first partial derivatives."""
    dims = tuple(S.SplineDimension(64, 3, 4096, max_derivative_order=1) for _ in range(2))
    grid = S.SplineGrid(dims, 3)
    cp = grid.control_points.obtain()
    cp.fill_(1.0)
    for der in ((0, 0), (1, 0), (0, 1)):
        S.evaluate_(grid, derivative_order=der)
        target = 1.0 if der == (0, 0) else 0.0
        # derivative rows of a clamped cubic basis reach ~3*61 in magnitude: cancellation error scales with that
        assert float((grid.eval - target).abs().max()) <= (1e-5 if der == (0, 0) else 2e-3)
    g = torch.Generator(device="cuda").manual_seed(2)
    cp.copy_(torch.rand(cp.shape, dtype=cp.dtype, device="cuda", generator=g))
    cp_np = S.to_numpy(cp)
    for der in ((0, 0), (1, 0), (0, 1)):
        S.evaluate_(grid, derivative_order=der)
        assert S.last_variant() == "evaluate_march2d"
        sel = [np.arange(0, 4096, 13), np.arange(5, 4096, 17)]
        ref = _oracle_subgrid(S, grid, cp_np, sel, der=der)
        ix = [torch.tensor(s, device="cuda") for s in sel]
        got = S.to_numpy(grid.eval[ix[0]][:, ix[1]])
        assert rel_err(got, ref) <= 1e-5, der
        y = torch.rand(grid.eval.shape[::-1], dtype=cp.dtype, device="cuda", generator=g).permute(2, 1, 0)
        grad = torch.zeros_like(cp)
        S.evaluate_adjoint_(grid, derivative_order=der, eval=y, control_points=grad)
        lhs, rhs = _dot(grid.eval, y), _dot(cp, grad)
        scale = float(grid.eval.double().norm() * y.double().norm())      # Cauchy-Schwarz scale (derivatives cancel)
        assert abs(lhs - rhs) <= 2e-5 * scale, der
    del grid
    torch.cuda.empty_cache()


def test_c4_full_size_2d_nurbs_float32(S):
    """C4: NURBS, control 256x256 with random weights, samples 8192x8192, Nout 3, Float32; the adjoint is this
    package's extension (transpose of the fixed-weights rational map)."""
    dims = tuple(S.SplineDimension(256, 3, 8192) for _ in range(2))
    grid = S.NURBSGrid(dims, 3)
    g = torch.Generator(device="cuda").manual_seed(3)
    grid.weights.copy_(0.5 + torch.rand(grid.weights.shape, dtype=grid.dtype, device="cuda", generator=g))
    cp = grid.control_points.obtain()
    cp.fill_(-1.5)                               # rational partition of unity: weighted average of a constant
    S.evaluate_(grid)
    assert S.last_variant() == "evaluate_march2d_nurbs"
    assert float((grid.eval + 1.5).abs().max()) <= 1.5 * 1e-5
    cp.copy_(torch.rand(cp.shape, dtype=cp.dtype, device="cuda", generator=g))
    S.evaluate_(grid)
    sel = [np.arange(0, 8192, 29), np.arange(7, 8192, 31)]
    ref = _oracle_subgrid(S, grid, S.to_numpy(cp), sel, weights_np=np.asfortranarray(S.to_numpy(grid.weights)))
    ix = [torch.tensor(s, device="cuda") for s in sel]
    assert rel_err(S.to_numpy(grid.eval[ix[0]][:, ix[1]]), ref) <= 1e-5
    y = torch.rand(grid.eval.shape[::-1], dtype=cp.dtype, device="cuda", generator=g).permute(2, 1, 0)
    grad = torch.zeros_like(cp)
    S.evaluate_adjoint_(grid, eval=y, control_points=grad, allow_nurbs=True)
    assert S.last_variant() in ("adjoint_passes_rational2d", "adjoint_fused2d_rational")
    lhs, rhs = _dot(grid.eval, y), _dot(cp, grad)
    assert abs(lhs - rhs) <= 2e-5 * abs(lhs)
    del grid, y
    torch.cuda.empty_cache()


def _build_c5(S, base, n_samples, nout=3):
    """3-level hierarchy of docs/src/theory_local_refinement.md:19-63, activation rectangles scaled by base/6."""
    f = base / 6.0
    sc = lambda lo, hi: range(max(1, int(round((lo - 1) * f)) + 1), int(round(hi * f)) + 1)
    dims = tuple(S.SplineDimension(base, 2, n_samples) for _ in range(2))
    grid = S.add_default_local_refinement(S.SplineGrid(dims, nout))
    for r in (((1, 4), (1, 6)), ((1, 6), (1, 2)), ((9, 10), (7, 10))):
        S.activate_local_control_point_range_(grid, sc(*r[0]), sc(*r[1]))
    S.deactivate_overwritten_control_points_(grid.control_points)
    grid = S.add_default_local_refinement(grid)
    for r in (((5, 12), (1, 4)), ((7, 8), (5, 6))):
        S.activate_local_control_point_range_(grid, sc(*r[0]), sc(*r[1]))
    S.deactivate_overwritten_control_points_(grid.control_points)
    return grid


def test_c5_scaled_three_level_hierarchy(S):
    """C5, throughput instance: base 64x64 -> 126^2 -> 250^2, samples 4096^2, Nout 3, Float32."""
    grid = _build_c5(S, 64, 4096)
    cps = grid.control_points
    assert [tuple(c.shape) for c in cps.control_points_refined] == [(64, 64, 3), (126, 126, 3), (250, 250, 3)]
    n_act = S.get_n_control_points(grid)
    assert n_act > 64 * 64 * 0.5
    # geometry is preserved by refinement + activation: finest control points unchanged by evaluate!(cp)
    fine0 = S.obtain(cps).clone()
    S.evaluate_(cps)
    assert rel_err(S.to_numpy(S.obtain(cps)), S.to_numpy(fine0)) <= 1e-5
    # evaluate! on the finest control points against the oracle on a strided sub-grid
    S.evaluate_(grid)
    assert S.last_variant() == "evaluate_march2d"
    sel = [np.arange(0, 4096, 19), np.arange(2, 4096, 23)]
    ref = _oracle_subgrid(S, grid, S.to_numpy(S.obtain(cps)), sel)
    ix = [torch.tensor(s_, device="cuda") for s_ in sel]
    assert rel_err(S.to_numpy(grid.eval[ix[0]][:, ix[1]]), ref) <= 1e-5
    # hierarchy + grid are mutually adjoint (what LinearMap/lsqr relies on)
    A = S.SplineGridLinearMap(grid)
    g = torch.Generator(device="cuda").manual_seed(5)
    xv = torch.rand(A.shape[1], dtype=grid.dtype, device="cuda", generator=g)
    yv = torch.rand(A.shape[0], dtype=grid.dtype, device="cuda", generator=g)
    Ax = A.matvec(xv).clone()
    Aty = A.rmatvec(yv)
    lhs, rhs = _dot(Ax, yv), _dot(xv, Aty)
    assert abs(lhs - rhs) <= 2e-5 * abs(lhs)
    # the finest level against the oracle hierarchy apply (dense mode-n products) on the same values
    from gpu_helpers import O
    vals = S.to_numpy(cps.get_values())
    lr_o = O.LocallyRefinedCP(
        [np.asfortranarray(S.to_numpy(c)).copy() for c in cps.control_points_refined],
        [O.LocalRefinementO(lr.dims_refinement,
                            [O.RefMat(A_.m, A_.n, A_.row_pointer_host, A_.column_start_host, A_.nzval_host) for A_ in lr.refinement_matrices],
                            S.to_numpy(lr.refinement_indices).reshape(-1, 2), S.to_numpy(lr.refinement_values).copy())
         for lr in cps.local_refinements])
    O.lrcp_evaluate(lr_o, fast=True)
    S.copyto_(cps, vals)
    S.evaluate_(cps)
    assert rel_err(S.to_numpy(S.obtain(cps)), lr_o.control_points_refined[-1]) <= 1e-5
    del grid
    torch.cuda.empty_cache()
