"""Helpers for the GPU parity tests: load the package, mirror package objects into oracle inputs."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import __graft_entry__ as entry  # noqa: E402

from oracle import oracle_c as OC  # noqa: E402
from oracle import oracle_np as O  # noqa: E402


def sg():
    return entry.load_package()


def oracle_dim(sd, impl=OC):
    """Recompute span indices and tables of a package SplineDimension with the oracle on the SAME
    knots_all / sample_points bit patterns.  Returns (table, indices)."""
    S = sg()
    ka, sp = S.to_numpy(sd.knot_vector.knots_all), S.to_numpy(sd.sample_points)
    idx = impl.span_indices(sp, ka, sd.degree)
    tab = impl.basis_tables(ka, sp, idx, sd.degree, sd.max_derivative_order)
    return tab, idx


def oracle_evaluate(grid, cp_np, der=None, weights_np=None, impl=OC):
    dims = [oracle_dim(sd, impl) for sd in grid.spline_dimensions]
    der = [0] * grid.Nin if der is None else list(der)
    return impl.evaluate([t for t, _ in dims], [i for _, i in dims], [sd.degree for sd in grid.spline_dimensions],
                         der, np.asfortranarray(cp_np), weights_np)


def oracle_adjoint(grid, eval_np, der=None, weights_np=None, impl=OC):
    dims = [oracle_dim(sd, impl) for sd in grid.spline_dimensions]
    der = [0] * grid.Nin if der is None else list(der)
    cp_shape = tuple(grid.control_point_grid_size) + (grid.Nout,)
    return impl.evaluate_adjoint([t for t, _ in dims], [i for _, i in dims],
                                 [sd.degree for sd in grid.spline_dimensions], der, np.asfortranarray(eval_np),
                                 cp_shape, weights_np)


def make_grid(n_cp, deg, n_s, nout, float_type="Float32", mdo=0, nurbs=False, distribution="equispaced", seed=0):
    S = sg()
    rng = np.random.default_rng(seed)
    mdos = mdo if isinstance(mdo, (tuple, list)) else [mdo] * len(n_cp)
    dims = tuple(S.SplineDimension(c, p, n, float_type=float_type, max_derivative_order=min(m, p),
                                   distribution=distribution, rng=rng)
                 for c, p, n, m in zip(n_cp, deg, n_s, mdos))
    grid = S.NURBSGrid(dims, nout) if nurbs else S.SplineGrid(dims, nout)
    npdt = np.float32 if float_type == "Float32" else np.float64
    cp = np.asfortranarray(rng.random(tuple(n_cp) + (nout,)).astype(npdt))
    S.copyto_(grid.control_points, cp)
    w = None
    if nurbs:
        w = np.asfortranarray((0.5 + rng.random(tuple(n_cp))).astype(npdt))
        grid.weights.copy_(S.to_device(w))
    return grid, cp, w, rng
