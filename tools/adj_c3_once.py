"""Development helper: a few C3 (or 8-way slab) evaluate_adjoint! / evaluate! calls, for `ncu` launch lists."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

slab = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = entry.load_package()
S.set_synchronous(False)
dims = tuple(S.SplineDimension(128, 3, 512, float_type="Float64") for _ in range(3))
grid = S.SlabShardedGrid(dims, 1, slab // 2, slab).local if slab > 1 else S.SplineGrid(dims, 1)
grid.control_points.obtain().copy_(torch.rand(grid.control_points.shape, dtype=torch.float64, device="cuda"))
e_in = S.jl_empty(grid.eval.shape, torch.float64, "cuda")
e_in.copy_(torch.rand(e_in.shape, dtype=torch.float64, device="cuda"))
g = torch.zeros_like(grid.control_points.obtain())
for _ in range(reps):
    S.evaluate_(grid)
    S.evaluate_adjoint_(grid, eval=e_in, control_points=g)
torch.cuda.synchronize()
print(S.last_variant())
