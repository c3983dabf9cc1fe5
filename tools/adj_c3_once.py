"""Run evaluate_adjoint! on C3 (or a slab of it) a few times -- target for ncu captures."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
slab = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # > 1: the middle slab of a slab-sharded grid
S = entry.load_package()
S.set_synchronous(False)
dims = tuple(S.SplineDimension(128, 3, n, float_type="Float64") for n in (512, 512, n3))
grid = S.SlabShardedGrid(dims, 1, slab // 2, slab).local if slab > 1 else S.SplineGrid(dims, 1)
e = S.jl_empty(grid.eval.shape, torch.float64, "cuda")
e.copy_(torch.rand(e.shape, dtype=torch.float64, device="cuda"))
g = torch.zeros_like(grid.control_points.obtain())
for _ in range(reps):
    S.evaluate_adjoint_(grid, eval=e, control_points=g)
torch.cuda.synchronize()
print(S.last_variant(), float(g.sum()))
