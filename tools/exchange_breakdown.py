"""Development helper (torchrun, N > 1): per-phase CUDA-event times of one C3 step on every rank --
evaluate!, adjoint with fused push, (signal + wait + reduce) -- max / mean over ranks.  Not the contract benchmark."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

S = entry.load_package()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
S.set_synchronous(False)
gdims = tuple(S.SplineDimension(128, 3, 512, float_type="Float64") for _ in range(3))
sh = S.SlabShardedGrid(gdims, 1, rank, world, peer_exchange=True)
grid = sh.local
cp = grid.control_points.obtain()
cp.copy_(torch.rand(cp.shape, dtype=cp.dtype, device=dev))
e_in = torch.rand(grid.eval.shape[::-1], dtype=cp.dtype, device=dev).permute(*reversed(range(grid.eval.dim())))
grad = torch.zeros_like(cp)
ex = sh.exchange
iters = 60
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(iters)]
for it in range(iters + 5):
    k = it - 5
    dist.barrier() if it == 5 else None
    if k >= 0:
        ev[k][0].record()
    S.evaluate_(grid)
    if k >= 0:
        ev[k][1].record()
    b = ex.step & 1
    ex.step += 1
    push = (b, ex.peer_ptrs[b], ex.world, ex.rank, ex.k0[ex.rank], ex.np_[ex.rank], ex.max_planes, 0, ex.mc_ptrs[b])
    S.evaluate_adjoint_(grid, control_points=grad, eval=e_in, _push=push)
    if k >= 0:
        ev[k][2].record()
    ex._wait_reduce_(grad, b, S._lib.stream_ptr(dev))
    if k >= 0:
        ev[k][3].record()
torch.cuda.synchronize()
t = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in ev[10:]])
med = torch.tensor(np.median(t, axis=0), device=dev)
allm = [torch.zeros_like(med) for _ in range(world)]
dist.all_gather(allm, med)
if rank == 0:
    a = torch.stack(allm).cpu().numpy()
    print(json.dumps({"world": world, "phases": ["evaluate!", "adjoint + fused push", "signal + wait + reduce"],
                      "median_ms_max_over_ranks": a.max(axis=0).round(4).tolist(), "median_ms_mean_over_ranks": a.mean(axis=0).round(4).tolist(),
                      "multicast": bool(any(ex.mc_ptrs))}))
dist.destroy_process_group()
