"""Development helper: time sg_exchange_wait_reduce_f64 alone on ONE device with the flags already set
(world ranks simulated: C3's control grid, slab supports of an N-way split)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as entry  # noqa: E402

S = entry.load_package()
lib = S._lib.lib()
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c1 = c2 = c3 = 128
P = 3
plane = c1 * c2
gd = S.SplineDimension(128, 3, 512, float_type="Float64")
idx = S.to_numpy(gd.sample_indices)
k0s, nps = [], []
for r in range(world):
    lo, hi = S.slab_bounds(512, world, r)
    k0s.append(int(idx[lo]) - P - 1)
    nps.append(int(idx[hi - 1]) - (int(idx[lo]) - P - 1))
max_planes = max(nps)
stage = torch.rand(world * max_planes * plane, dtype=torch.float64, device="cuda")
grad = torch.empty(plane * c3, dtype=torch.float64, device="cuda")
flags = torch.full((16,), 1 << 40, dtype=torch.int64, device="cuda")
sync = torch.zeros(8, dtype=torch.int64, device="cuda")
k0a = (C.c_int64 * world)(*k0s)
npa = (C.c_int64 * world)(*nps)
st = S._lib.stream_ptr(torch.device("cuda"))


def call():
    S._lib.check(lib.sg_exchange_wait_reduce_f64(S._lib.ptr(grad), S._lib.ptr(stage), S._lib.ptr(flags), S._lib.ptr(sync), C.c_void_p(0),
                                                 C.c_int(world), C.c_int(0), k0a, npa, C.c_int64(plane), C.c_int64(c3), C.c_int(1),
                                                 C.c_int64(max_planes), st), "wait_reduce")


for _ in range(5):
    call()
torch.cuda.synchronize()
ts = []
for _ in range(9):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    call()
    e0.record()
    for _ in range(8):
        call()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 8)
print({"world": world, "wait_reduce_ms_median": round(float(np.median(ts)), 4), "bytes_MB": round((sum(nps) * plane + plane * c3) * 8 / 1e6, 1)})
