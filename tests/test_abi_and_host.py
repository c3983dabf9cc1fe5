"""CPU-only tests: the C-ABI library loads and exports every declared symbol, host-side logic
(validation messages, ranges, slab partition), and loud failure without a device."""
import ctypes as C
import logging

import numpy as np
import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def S():
    return entry.load_package()


def test_library_exports_every_declared_symbol(S):
    lib = S._lib.lib()
    names = S._lib.declared_symbols()
    assert len(names) >= 38
    for fam in ("sg_evaluate_", "sg_evaluate_adjoint_", "sg_basis_tables_", "sg_span_indices_", "sg_refmat_mul_",
                "sg_refmat_mul_adjoint_", "sg_scatter_active_", "sg_gather_zero_active_", "sg_allreduce_sum_"):
        assert fam + "f32" in names and fam + "f64" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.sg_version() == 100
    assert lib.sg_status_string(0) == b"SG_OK"
    assert lib.sg_status_string(-2) == b"SG_ERR_UNSUPPORTED"


def test_argument_validation_needs_no_device(S):
    """Argument errors are reported before any CUDA call (no compute without a GPU)."""
    lib = S._lib.lib()
    assert lib.sg_span_indices_f64(None, None, C.c_int64(0), None, C.c_int64(0), C.c_int(0), None) == -1
    n = (C.c_int64 * 1)(4)
    assert lib.sg_evaluate_f32(None, C.c_int(1), n, n, C.c_int(1), None, None, None, None, None, None, None, None) == -1
    assert lib.sg_evaluate_adjoint_workspace_bytes(C.c_int(0), None, None, C.c_int(1), None, C.c_int(4), C.c_int(0)) == 0
    ns, nc, deg = (C.c_int64 * 2)(4096, 4096), (C.c_int64 * 2)(64, 64), (C.c_int * 2)(3, 3)
    lib.sg_evaluate_adjoint_workspace_bytes.restype = C.c_size_t
    small = lib.sg_evaluate_adjoint_workspace_bytes(C.c_int(2), ns, nc, C.c_int(3), deg, C.c_int(4), C.c_int(0))
    big = lib.sg_evaluate_adjoint_workspace_bytes(C.c_int(2), ns, nc, C.c_int(3), deg, C.c_int(4), C.c_int(1))
    assert 0 < small and big >= small + 4096 * 4096 * 4


def test_no_cpu_fallback(S):
    """Without a CUDA device the product path must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.SplineDimension(5, 2, 10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.rmeye(4)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under splinegrids.jl_b200/ may reference it."""
    import re
    for p in entry.PKG_DIR.rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".jl") and p.is_file():
            text = p.read_text()
            assert not re.search(r"\b(from|import)\s+oracle\b|oracle_np|oracle_c|liboracle", text), p


def test_knot_vector_validation_messages(S):
    """test/test_knot_vector.jl:11-24: raised before any device work."""
    with pytest.raises(AssertionError, match="knot_values and multiplicities must be of the same length."):
        S.KnotVector(np.arange(1, 6), np.arange(1, 4))
    with pytest.raises(AssertionError, match="knot_values must be sorted."):
        S.KnotVector(np.array([3, 2, 1]), np.arange(1, 4))
    with pytest.raises(AssertionError, match="knot_values must be unique."):
        S.KnotVector(np.array([1, 1, 2]), np.arange(1, 4))


def test_julia_range_is_correctly_rounded(S):
    from importlib import import_module
    kvmod = import_module("splinegrids_jl_b200.knot_vector")
    r = kvmod.julia_range(0, 1, 4, np.float64)
    assert r.tolist() == [0.0, 1 / 3, 2 / 3, 1.0]
    r32 = kvmod.julia_range(0, 1, 7, np.float32)
    assert r32[3] == np.float32(0.5) and r32.dtype == np.float32
    assert kvmod.julia_range(5.0, 7.0, 3, np.float64).tolist() == [5.0, 6.0, 7.0]


def test_derivative_validation_host_side(S, caplog):
    from importlib import import_module
    val = import_module("splinegrids_jl_b200.validation")

    class FakeDim:
        def __init__(self, mdo):
            self.max_derivative_order = mdo

    with caplog.at_level(logging.ERROR, logger="splinegrids_b200"):
        with pytest.raises(val.SplineGridsError, match=r"Invalid derivative order\(s\) supplied"):
            val.validate_partial_derivatives((FakeDim(1), FakeDim(1)), (2, 3))
    assert [r.getMessage() for r in caplog.records] == [
        "The maximum derivative order available for spline dimension 1 is 1, got 2.",
        "The maximum derivative order available for spline dimension 2 is 1, got 3."]
    val.validate_partial_derivatives((FakeDim(1), FakeDim(0)), (1, 0))
    with pytest.raises(val.SplineGridsError, match="Computing derivatives of NURBS is currently not supported."):
        val.validate_partial_derivatives((FakeDim(1),), (1,), is_nurbs=True)


def test_slab_bounds_partition(S):
    for n in (1, 7, 64, 512, 8192, 1000):
        for ws in (1, 2, 3, 4, 8):
            if n < ws:
                continue
            b = [S.slab_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert S.slab_bounds(512, 8, 3) == (192, 256)


def test_support_plane_ranges_host_side(S):
    """sg_exchange_support_ranges (pure host code): which local planes of a rank's support go to which rank in the
    support-plane gradient exchange -- the overlap of the two supports, everything into the own slot, nothing to ranks
    whose slabs do not share a plane; and the owned ranges used by the end-of-fit all-gather."""
    import ctypes as C
    lib = S._lib.lib()
    k0, npl = [0, 5, 11, 17], [8, 9, 9, 6]                     # supports [0,8) [5,14) [11,20) [17,23): neighbours share 3 planes
    world, max_planes = 4, max(npl)
    k0s, nps = S._lib.i64_array(k0), S._lib.i64_array(npl)
    expect = {0: [(0, 9), (5, 8), (8, 8), (8, 8)],             # local plane l of rank r is global plane k0[r] + l
              1: [(0, 3), (0, 9), (6, 9), (9, 9)],
              2: [(0, 0), (0, 3), (0, 9), (6, 9)],
              3: [(0, 0), (0, 0), (0, 3), (0, 9)]}
    for r in range(world):
        lo, hi = (C.c_int * 16)(), (C.c_int * 16)()
        assert lib.sg_exchange_support_ranges(C.c_int(world), C.c_int(r), k0s, nps, C.c_int64(max_planes), lo, hi) == 0
        got = [(lo[q], hi[q]) for q in range(world)]
        for q in range(world):
            planes = set(range(k0[r] + got[q][0], k0[r] + got[q][1]))
            if q == r:
                assert got[q] == (0, max_planes)
            else:                                                # exactly the planes both supports contain
                assert planes == set(range(k0[r], k0[r] + npl[r])) & set(range(k0[q], k0[q] + npl[q])), (r, q, got)
        assert [b - a for a, b in got] == [b - a for a, b in expect[r]]
    assert lib.sg_exchange_support_ranges(C.c_int(4), C.c_int(4), k0s, nps, C.c_int64(9), lo, hi) != 0   # bad rank
    assert [S.owned_planes(k0, npl, r) for r in range(world)] == [(0, 5), (5, 11), (11, 17), (17, 23)]


def test_slab_supports_and_owned_planes_random_layouts(S):
    """slab_supports / owned_planes on random monotone span-index arrays (irregular sample spacing, thin slabs whose
    supports reach past the direct neighbour): a rank's support is exactly the set of control planes its rows touch, the owned
    ranges are disjoint, lie inside the supports and cover the union of the supports."""
    import torch
    from types import SimpleNamespace
    rng = np.random.default_rng(3)
    for trial in range(200):
        p = int(rng.integers(1, 5))
        c = int(rng.integers(p + 1, 40))
        n = int(rng.integers(1, 60))
        world = int(rng.integers(1, min(n, 8) + 1))
        idx = np.sort(rng.integers(p + 1, c + 1, size=n)).astype(np.int32)      # 1-based span indices in [p + 1, c]
        last = SimpleNamespace(sample_indices=torch.from_numpy(idx), degree=p, n_sample_points=n)
        k0s, nps = S.slab_supports(last, world)
        union = set()
        for r in range(world):
            lo, hi = S.slab_bounds(n, world, r)
            touched = {int(i) - p - 1 + a for i in idx[lo:hi] for a in range(p + 1)}
            sup = set(range(k0s[r], k0s[r] + nps[r]))
            assert touched <= sup and min(touched) == k0s[r] and max(touched) == k0s[r] + nps[r] - 1
            assert 0 <= k0s[r] and k0s[r] + nps[r] <= c
            union |= sup
        owned = [S.owned_planes(k0s, nps, r) for r in range(world)]
        flat = [k for lo, hi in owned for k in range(lo, hi)]
        assert len(flat) == len(set(flat)), (trial, owned)
        assert set(flat) == union, (trial, k0s, nps, owned)
        assert all(k0s[r] <= lo and hi <= k0s[r] + nps[r] for r, (lo, hi) in enumerate(owned))


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (the CPU arm: the oracle on the host cores) prints ONE JSON line with the contract keys,
    for the default workload and for another BASELINE config; all host threads are used whatever OMP_NUM_THREADS says."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, OMP_NUM_THREADS="1")            # what torchrun exports to its workers
    for extra in ([], ["--config", "C2"]):
        out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                              "--cpu-planes", "2", *extra], capture_output=True, text=True, env=env, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "dtype", "data", "config", "cpu_baseline", "e2e"):
            assert key in d, key
        assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
        assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_oracle_range_activation_row_order_matches_the_reference():
    """activate_local_control_point_range! flattens the ADJOINT of the Iterators.product array (src/control_points.jl:535-537):
    for Nin = 2 the rows come out with dimension 2 fastest."""
    import numpy as np
    from oracle import oracle_np as O
    seen = {}
    O_activate = O.activate_local_refinement
    try:
        O.activate_local_refinement = lambda lrcp, idx, *a, **k: seen.setdefault("idx", np.array(idx))
        O.activate_local_control_point_range(None, (2, 3), (5, 7))
    finally:
        O.activate_local_refinement = O_activate
    assert seen["idx"].tolist() == [[2, 5], [2, 6], [2, 7], [3, 5], [3, 6], [3, 7]]
