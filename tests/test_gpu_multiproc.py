"""Multi-process GPU test (one process per GPU, NCCL): skipped unless the box has at least 2 CUDA devices.

Covers what the single-device tests can only simulate:
* the slab-sharded adjoint with the PEER-MEMORY gradient exchange (fused push + the C ABI's own flag barrier + reduce)
  on real peers, eagerly and replayed from a CUDA graph, against (a) an NCCL all-reduce of the local partial gradients
  and (b) the C oracle's adjoint of the full grid;
* the C ABI's own NCCL entry points `sg_comm_unique_id / sg_comm_create / sg_allreduce_sum / sg_comm_destroy`
  (what a Julia host would call), executed through ctypes.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import ctypes as C

    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    from oracle import oracle_c as OC
    S = entry.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lib = S._lib.lib()
        # ---- sg_comm_* / sg_allreduce_sum through the C ABI -------------------------------------------------
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            S._lib.check(lib.sg_comm_unique_id(buf), "sg_comm_unique_id")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        uid_c = (C.c_ubyte * 128)(*uid.cpu().tolist())
        comm = C.c_void_p()
        S._lib.check(lib.sg_comm_create(C.byref(comm), C.c_int(world), C.c_int(rank), uid_c), "sg_comm_create")
        for dt, suf in ((torch.float64, "f64"), (torch.float32, "f32")):
            x = torch.arange(1000, dtype=dt, device=dev) * (rank + 1)
            S._lib.check(getattr(lib, "sg_allreduce_sum_" + suf)(S._lib.ptr(x), C.c_int64(x.numel()), comm,
                                                                 S._lib.stream_ptr(dev)), "sg_allreduce_sum")
            torch.cuda.synchronize()
            expect = torch.arange(1000, dtype=dt, device=dev) * (world * (world + 1) // 2)
            assert torch.equal(x, expect)
        S._lib.check(lib.sg_comm_destroy(comm), "sg_comm_destroy")

        # ---- slab-sharded adjoint + peer-memory exchange ------------------------------------------------------
        n_cp, deg, n_s, nout = (24, 20, 40), (3, 3, 3), (128, 72, 160), 2
        gdims = tuple(S.SplineDimension(c, p, n, float_type="Float64") for c, p, n in zip(n_cp, deg, n_s))
        rng = np.random.default_rng(5)                             # same full input on every rank
        e_full = np.asfortranarray(rng.random(n_s + (nout,)))
        sh = S.SlabShardedGrid(gdims, nout, rank, world, peer_exchange=True)
        assert sh.exchange is not None, "peer-memory exchange unavailable"
        e_loc = S.to_device(np.asfortranarray(e_full[:, :, sh.lo:sh.hi, :]))
        grad = torch.full_like(sh.local.control_points.obtain(), float("nan"))
        # (a) NCCL all-reduce of the local partial gradients
        ref = torch.zeros_like(grad)
        S.evaluate_adjoint_(sh.local, eval=e_loc, control_points=ref)
        S.allreduce_gradient_(ref)
        # (b) the oracle on the full grid
        tabs = [np.asfortranarray(S.to_numpy(sd.eval)) for sd in gdims]
        idxs = [np.ascontiguousarray(S.to_numpy(sd.sample_indices)) for sd in gdims]
        gref = OC.evaluate_adjoint(tabs, idxs, list(deg), [0, 0, 0], e_full, n_cp + (nout,))
        for it in range(3):                                         # alternating staging buffers
            grad.fill_(float("nan"))
            sh.evaluate_adjoint_(eval=e_loc, control_points=grad)
            torch.cuda.synchronize()
            assert S.last_variant().startswith("adjoint_march2")
            assert float((grad - ref).norm() / ref.norm()) <= 1e-13, it
            got = S.to_numpy(grad)
            assert np.linalg.norm(got - gref) / np.linalg.norm(gref) <= 1e-12
        # separate push kernel (not fused into the adjoint)
        sh.fused_push = False
        grad.fill_(float("nan"))
        sh.evaluate_adjoint_(eval=e_loc, control_points=grad)
        torch.cuda.synchronize()
        assert float((grad - ref).norm() / ref.norm()) <= 1e-13
        sh.fused_push = True
        # the exchange inside a CUDA graph (device-side barrier: no host call in the step)
        def step():
            sh.evaluate_adjoint_(eval=e_loc, control_points=grad)
        cap = S.CapturedCalls(step, unroll=2, warmup=1)
        for _ in range(3):
            grad.fill_(float("nan"))
            cap.replay()
            torch.cuda.synchronize()
            assert float((grad - ref).norm() / ref.norm()) <= 1e-13
        ep, timed_out = sh.exchange.status()
        assert not timed_out
        counts = [None] * world
        dist.all_gather_object(counts, ep)
        assert len(set(counts)) == 1, counts                        # every rank completed the same number of exchanges
        # every rank holds bit-identical gradients (rank-order summation)
        mine = grad.permute(*reversed(range(grad.dim()))).contiguous()   # column-major array as a contiguous tensor
        g0 = mine.clone()
        dist.broadcast(g0, 0)
        assert torch.equal(g0, mine)
        # ---- support-plane exchange (halo variant): summed gradient on the planes this rank's slab reads -------------
        full_grad = grad.clone()
        sh2 = S.SlabShardedGrid(gdims, nout, rank, world, peer_exchange=True, exchange_mode="support")
        assert sh2.exchange is not None and sh2.exchange.mode == "support"
        lo, hi = sh2.exchange.support_planes()
        def step2():
            sh2.evaluate_adjoint_(eval=e_loc, control_points=grad)
        cap2 = S.CapturedCalls(step2, unroll=2, warmup=1)
        for it in range(4):
            grad.fill_(float("nan"))
            if it < 2:
                step2()
            else:
                cap2.replay()
            torch.cuda.synchronize()
            assert S.last_variant().startswith("adjoint_march2")
            assert torch.equal(grad[:, :, lo:hi, :], full_grad[:, :, lo:hi, :]), it     # same bits as the replicated exchange
            assert bool(torch.isnan(grad[:, :, :lo, :]).all()) and bool(torch.isnan(grad[:, :, hi:, :]).all())
        assert not sh2.exchange.status()[1]
        # end of a fit: the replicated array is rebuilt from the planes every rank owns (NCCL broadcasts)
        sh2.exchange.allgather_support_planes_(grad)
        torch.cuda.synchronize()
        assert torch.equal(grad, full_grad)
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_exchange_and_comm_abi_multiprocess(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices, found {torch.cuda.device_count()}")
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() + world) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").exists()
