import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as entry
S = entry.load_package()
S.set_kernel_policy(2)
import json
CASES = json.loads(os.environ.get("CASES", '[["Float64", [[128,128,128],[3,3,3],[512,512,64]]]]'))
for ft, shape in CASES:
    n_cp, deg, n_s = shape
    dims = tuple(S.SplineDimension(c, p, n, float_type=ft) for c, p, n in zip(n_cp, deg, n_s))
    g = S.SplineGrid(dims, 1)
    g.control_points.obtain().copy_(torch.rand(g.control_points.shape, device="cuda", dtype=g.dtype))
    S.evaluate_(g)
    torch.cuda.synchronize()
    print("first call variant:", S.last_variant(), flush=True)
    a = g.eval.clone()
    os.environ["SG_EVAL_TMA"] = "0"
    S.evaluate_(g)
    torch.cuda.synchronize()
    os.environ["SG_EVAL_TMA"] = "1"
    print(ft, S.last_variant(), float((a - g.eval).abs().max()))
