"""Development probe: time evaluate!/evaluate_adjoint! on the BASELINE configs with CUDA events and print
one JSON line per (config, op, policy).  Not the contract benchmark (that is bench.py)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

CONFIGS = {
    "C1": dict(n_cp=(10, 10, 5), deg=(2, 3, 2), n_s=(50, 50, 25), nout=4, ft="Float64"),
    "C2": dict(n_cp=(64, 64), deg=(3, 3), n_s=(4096, 4096), nout=3, ft="Float32", mdo=1),
    "C3": dict(n_cp=(128, 128, 128), deg=(3, 3, 3), n_s=(512, 512, 512), nout=1, ft="Float64"),
    "C3s": dict(n_cp=(128, 128, 128), deg=(3, 3, 3), n_s=(512, 512, 64), nout=1, ft="Float64"),
    "C4": dict(n_cp=(256, 256), deg=(3, 3), n_s=(8192, 8192), nout=3, ft="Float32", nurbs=True),
    "C5s": dict(n_cp=(250, 250), deg=(2, 2), n_s=(4096, 4096), nout=3, ft="Float32"),
    "C2p": dict(n_cp=(64, 64), deg=(3, 3), n_s=(4160, 4096), nout=3, ft="Float32", mdo=1),    # C2 with a row stride that is not a power of two
    "C5d": dict(n_cp=(18, 18), deg=(2, 2), n_s=(500, 500), nout=3, ft="Float32", mdo=1),      # docs-size finest level of C5
    "M1k": dict(n_cp=(64, 64), deg=(3, 3), n_s=(1024, 1024), nout=3, ft="Float32", mdo=1),
}


def algorithmic_bytes(cfg):
    s = 4 if cfg["ft"] == "Float32" else 8
    n, c = np.prod(cfg["n_s"], dtype=np.int64), np.prod(cfg["n_cp"], dtype=np.int64)
    b = n * cfg["nout"] * s + c * cfg["nout"] * s
    if cfg.get("nurbs"):
        b += c * s
    b += sum(nd * ((p + 1) * s + 4) for nd, p in zip(cfg["n_s"], cfg["deg"]))
    return int(b)


def time_op(fn, iters, warmup, flush=None, batch=1):
    """Median / min per-call device time.  batch > 1 enqueues several calls between the two events so the
    Python launch overhead (tens of microseconds) is hidden behind the previous kernel."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if batch > 1:
            fn()                      # keeps the queue busy while the timed launches are enqueued
        e0.record()
        for _ in range(batch):
            fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / batch)
    return float(np.median(times)), float(np.min(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3,C4")
    ap.add_argument("--policies", default="0")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--ops", default="evaluate,adjoint")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--slab", type=int, default=1, help="time the middle slab of an N-way slab-sharded grid (1 GPU)")
    args = ap.parse_args()
    S = entry.load_package()
    S.set_synchronous(False)
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for name in args.configs.split(","):
        cfg = CONFIGS[name]
        rng = np.random.default_rng(1)
        dims = tuple(S.SplineDimension(c, p, n, float_type=cfg["ft"], max_derivative_order=cfg.get("mdo", 0))
                     for c, p, n in zip(cfg["n_cp"], cfg["deg"], cfg["n_s"]))
        if args.slab > 1:
            grid = S.SlabShardedGrid(dims, cfg["nout"], args.slab // 2, args.slab, nurbs=bool(cfg.get("nurbs"))).local
        else:
            grid = S.NURBSGrid(dims, cfg["nout"]) if cfg.get("nurbs") else S.SplineGrid(dims, cfg["nout"])
        dt = torch.float32 if cfg["ft"] == "Float32" else torch.float64
        grid.control_points.obtain().copy_(torch.rand(grid.control_points.shape, dtype=dt, device="cuda"))
        if cfg.get("nurbs"):
            grid.weights.copy_(0.5 + torch.rand(grid.weights.shape, dtype=dt, device="cuda"))
        e_in = S.jl_empty(grid.eval.shape, dt, "cuda")
        e_in.copy_(torch.rand(e_in.shape, dtype=dt, device="cuda"))
        g_out = torch.zeros_like(grid.control_points.obtain())
        nbytes = algorithmic_bytes(cfg)
        values = int(np.prod(cfg["n_s"])) * cfg["nout"]
        if args.slab > 1:
            cfg_l = dict(cfg, n_s=tuple(grid.eval.shape[:-1]))
            nbytes = algorithmic_bytes(cfg_l)
            values = int(np.prod(cfg_l["n_s"])) * cfg["nout"]
        for pol in [int(p) for p in args.policies.split(",")]:
            S.set_kernel_policy(pol)
            for op in args.ops.split(","):
                if op == "multi":           # value + first partials (C2's three calls) as one library call
                    ders = [(0,) * len(cfg["n_cp"])] + [tuple(1 if d == k else 0 for d in range(len(cfg["n_cp"]))) for k in range(len(cfg["n_cp"]))]
                    evs = [torch.empty_like(grid.eval) for _ in ders]
                    fn = lambda: S.evaluate_multi_(grid, ders, evs)
                elif op == "separate":
                    ders = [(0,) * len(cfg["n_cp"])] + [tuple(1 if d == k else 0 for d in range(len(cfg["n_cp"]))) for k in range(len(cfg["n_cp"]))]
                    evs = [torch.empty_like(grid.eval) for _ in ders]
                    fn = lambda: [S.evaluate_(grid, derivative_order=d, eval=e) for d, e in zip(ders, evs)]
                elif op == "evaluate":
                    fn = lambda: S.evaluate_(grid)
                else:
                    fn = lambda: S.evaluate_adjoint_(grid, eval=e_in, control_points=g_out, allow_nurbs=True)
                big = pol == 1 and values > 5e7
                med, best = time_op(fn, 2 if big else args.iters, 1 if big else args.warmup,
                                    flush if (nbytes < 200e6 and args.batch == 1) else None,
                                    1 if big else args.batch)
                print(json.dumps({"config": name + (f"/slab{args.slab}" if args.slab > 1 else ""), "op": op, "policy": pol, "variant": S.last_variant(),
                                  "ms_median": round(med, 4), "ms_min": round(best, 4),
                                  "values_per_s": values / (med * 1e-3), "alg_GBs": nbytes / (med * 1e-3) / 1e9,
                                  "frac_of_measured_hbm": nbytes / (med * 1e-3) / 1e9 / hbm}), flush=True)
            S.set_kernel_policy(0)
        del grid, e_in, g_out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
