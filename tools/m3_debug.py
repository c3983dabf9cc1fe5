"""Debug helper: one small single-pass 3-D adjoint against the C oracle (prints errors / mismatch)."""
import os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
os.environ["SG_ADJ_MARCH3"] = "1"
from gpu_helpers import make_grid, oracle_adjoint, sg
from helpers import rel_err
S = sg()
cases = [((20, 11, 9), (3, 3, 3), (128, 40, 20), 1, "Float64", "equispaced"), ((30, 12, 25), (2, 1, 1), (256, 40, 300), 2, "Float64", "random"),
         ((30, 12, 25), (2, 1, 1), (256, 40, 300), 2, "Float32", "random"), ((33, 34, 35), (3, 3, 3), (136, 140, 150), 2, "Float64", "random")]
for n_cp, deg, n_s, nout, ft, dist in cases:
    grid, cp, w, rng = make_grid(n_cp, deg, n_s, nout, ft, seed=31, distribution=dist)
    e = np.asfortranarray(rng.random(tuple(n_s) + (nout,)).astype(cp.dtype))
    g = torch.full_like(grid.control_points.obtain(), -3.0)
    S.evaluate_adjoint_(grid, eval=S.to_device(e), control_points=g)
    torch.cuda.synchronize()
    ref = oracle_adjoint(grid, e)
    got = S.to_numpy(g)
    bad = np.argwhere(np.abs(got - ref) > 1e-4 * np.abs(ref).max())
    print(n_cp, n_s, ft, dist, S.last_variant(), rel_err(got, ref), "bad:", len(bad), bad[:6].tolist(), flush=True)
