"""world_size-2 gloo test (CPU) of the multi-GPU host logic: slab partition + gradient all-reduce.

The kernels need a GPU, so here each rank computes its slab's partial adjoint with the ORACLE (the
checker) and the product's `allreduce_gradient_` combines them; the result must equal the single-process
adjoint.  This covers the N>1 data flow: slab bounds, sliced span indices, partial gradients, all-reduce.
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world_size: int, port: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    from oracle import oracle_np as O
    S = entry.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        rng = np.random.default_rng(7)                        # same inputs on every rank
        n_cp, deg, n_s, nout = (7, 6, 5), (3, 2, 2), (12, 10, 9), 2
        dims = [O.make_dimension(c, p, n, dtype=np.float64) for c, p, n in zip(n_cp, deg, n_s)]
        e = np.asfortranarray(rng.random(n_s + (nout,)))
        cp = np.asfortranarray(rng.random(n_cp + (nout,)))
        lo, hi = S.slab_bounds(n_s[-1], world_size, rank)
        tabs = [d.eval for d in dims[:-1]] + [np.asfortranarray(dims[-1].eval[lo:hi])]
        idxs = [d.sample_indices for d in dims[:-1]] + [dims[-1].sample_indices[lo:hi].copy()]
        # forward: slab of the global evaluation, no communication
        full = O.evaluate([d.eval for d in dims], [d.sample_indices for d in dims], deg, [0, 0, 0], cp)
        part = O.evaluate(tabs, idxs, deg, [0, 0, 0], cp)
        assert np.array_equal(part, full[:, :, lo:hi, :])
        # adjoint: partial gradient of the slab, then the product's all-reduce
        g_part = O.evaluate_adjoint(tabs, idxs, deg, [0, 0, 0], np.asfortranarray(e[:, :, lo:hi, :]), n_cp + (nout,))
        # the slab's partial gradient is zero outside control planes [idx(lo)-p, idx(hi-1)]
        k0 = int(dims[-1].sample_indices[lo]) - deg[-1] - 1
        k1 = int(dims[-1].sample_indices[hi - 1])
        assert np.all(g_part[:, :, :k0, :] == 0) and np.all(g_part[:, :, k1:, :] == 0)
        t = torch.from_numpy(np.ascontiguousarray(g_part.T)).permute(3, 2, 1, 0)   # column-major view (shares memory)
        assert S.is_colmajor(t)
        S.allreduce_gradient_(t)
        g_full = O.evaluate_adjoint([d.eval for d in dims], [d.sample_indices for d in dims], deg, [0, 0, 0], e,
                                    n_cp + (nout,))
        err = np.linalg.norm(t.numpy() - g_full) / np.linalg.norm(g_full)
        assert err < 1e-13, err
        # support-plane exchange, host side: supports of all ranks, owned ranges, end-of-fit all-gather.  The exchange kernels
        # need GPUs (tests/test_gpu_parity.py, test_gpu_multiproc.py); their RESULT is emulated here: the summed gradient on
        # this rank's support planes, NaN elsewhere.
        from types import SimpleNamespace
        last = SimpleNamespace(sample_indices=torch.from_numpy(dims[-1].sample_indices.astype(np.int32)), degree=deg[-1],
                               n_sample_points=n_s[-1])
        k0s, nps = S.slab_supports(last, world_size)
        assert (k0s[rank], k0s[rank] + nps[rank]) == (k0, k1)
        owned = [S.owned_planes(k0s, nps, r) for r in range(world_size)]
        covered = sorted(k for lo_, hi_ in owned for k in range(lo_, hi_))
        assert covered == sorted(set(covered))                                   # disjoint
        assert set(covered) == {k for a, n in zip(k0s, nps) for k in range(a, a + n)}   # union of the supports
        assert all(k0s[r] <= owned[r][0] and owned[r][1] <= k0s[r] + nps[r] for r in range(world_size))
        g_sup = torch.full_like(t, float("nan"))
        g_sup[..., k0:k1, :] = t[..., k0:k1, :]
        S.allgather_support_planes_(g_sup, k0s, nps)
        assert set(covered) == set(range(n_cp[-1])) and np.array_equal(g_sup.numpy(), t.numpy())
        Path(out_dir, f"ok{rank}").write_text(f"{err}")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_slab_sharded_adjoint_allreduce_world_size_2(tmp_path, world):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + 7 * world) % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
