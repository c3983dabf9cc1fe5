"""One rank's slab step of the C3 grid on ONE GPU: evaluate!, evaluate_adjoint! with the fused gradient push (support-plane
variant, keep_local = 0; the "peers" are staging buffers on the same device) -- CUDA events, 8 calls per measurement.
usage: python tools/slab_push_probe.py [world] [rank]"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402


def timed(fn, iters=9, batch=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(batch):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / batch)
    return float(np.median(ts))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else world // 2
    S = entry.load_package()
    S.set_synchronous(False)
    n_cp, deg, n_s, nout = (128, 128, 128), (3, 3, 3), (512, 512, 512), 1
    gdims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip(n_cp, deg, n_s))
    idx3 = S.to_numpy(gdims[2].sample_indices)
    shards = [S.SlabShardedGrid(gdims, nout, r, world) for r in range(world)]
    k0 = [int(idx3[sh.lo]) - deg[2] - 1 for sh in shards]
    npl = [int(idx3[sh.hi - 1]) - k for sh, k in zip(shards, k0)]
    max_planes, plane = max(npl), n_cp[0] * n_cp[1]
    sh = shards[rank]
    grid = sh.local
    grid.control_points.obtain().copy_(torch.rand(grid.control_points.shape, dtype=torch.float64, device="cuda"))
    e_in = torch.rand(grid.eval.shape[::-1], dtype=torch.float64, device="cuda").permute(*reversed(range(grid.eval.dim())))
    g = torch.zeros_like(grid.control_points.obtain())
    stages = [torch.zeros(world * nout * max_planes * plane, dtype=torch.float64, device="cuda") for _ in range(world)]
    peer = (C.c_void_p * world)(*[t.data_ptr() for t in stages])
    k0s, nps = S._lib.i64_array(k0), S._lib.i64_array(npl)
    out = {"world": world, "rank": rank, "support_planes": npl[rank]}
    out["evaluate_ms"] = timed(lambda: S.evaluate_(grid))
    out["adjoint_plain_ms"] = timed(lambda: S.evaluate_adjoint_(grid, eval=e_in, control_points=g))
    push_s = (("probe", "support"), peer, world, rank, k0[rank], npl[rank], max_planes, 0, 0, (k0s, nps))
    out["adjoint_push_support_ms"] = timed(lambda: S.evaluate_adjoint_(grid, eval=e_in, control_points=g, _push=push_s))
    push_r = (("probe", "replicated"), peer, world, rank, k0[rank], npl[rank], max_planes, 0, 0)
    out["adjoint_push_replicated_local_peers_ms"] = timed(lambda: S.evaluate_adjoint_(grid, eval=e_in, control_points=g, _push=push_r))
    out["variant"] = S.last_variant()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
