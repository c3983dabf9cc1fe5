#!/usr/bin/env python
"""Benchmark of the grid-evaluation hot path (BASELINE.json metric: evaluate!/adjoint sample-values/sec).

Workload (config.workload): BASELINE config 3 -- 3-D cubic volume, 128^3 control points, 512^3 samples, Nout=1,
Float64 (1.07 GB output, > L2).  One STEP = one `evaluate!` + one `evaluate_adjoint!` on that grid (what one
iteration of the reference's LinearMap/lsqr fitting loop does, ext/SplineGridsLinearMapsExt.jl:16-48).
With N GPUs the sample grid is sharded in slabs along its slowest axis (strong scaling: total work fixed),
control points replicated, and the adjoint's per-GPU gradients are summed with an NCCL all-reduce.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

`value`   = sample-values/s, kernels only, inputs resident in HBM (CUDA events, max over ranks).
`e2e`     = same metric through the public API with HOST buffers: pinned H2D of the inputs and D2H of the
            results inside the timed region.
`roofline`= dominant kernel's algorithmic bytes / its live CUDA-event duration vs MEASURED_PEAKS.json.
The CPU arm is the C/OpenMP restatement of the reference's algorithm (oracle/): Julia is not installed in the
image, so the reference package itself cannot run (DESIGN.md "Oracle").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs.  C3 is the one the metric is quoted on: the default and the only one the
    # driver runs; the others (N = 1 only, `--config`) give one line per config for profiles/.
    "C1": dict(name="C1: README 3-D grid, control (10,10,5), degree (2,3,2), samples (50,50,25), Nout=4, Float64",
               n_cp=(10, 10, 5), degree=(2, 3, 2), n_samples=(50, 50, 25), nout=4, float_type="Float64"),
    "C2": dict(name="C2: 2-D cubic surface, control 64x64, samples 4096x4096, Nout=3, Float32",
               n_cp=(64, 64), degree=(3, 3), n_samples=(4096, 4096), nout=3, float_type="Float32"),
    "C3": dict(name="C3: 3-D cubic volume, control 128^3, samples 512^3, Nout=1, Float64",
               n_cp=(128, 128, 128), degree=(3, 3, 3), n_samples=(512, 512, 512), nout=1, float_type="Float64"),
    "C4": dict(name="C4: 2-D NURBS surface, control 256x256 with random weights, samples 8192x8192, Nout=3, Float32 "
                    "(adjoint = this package's extension: transpose of the fixed-weights rational map)",
               n_cp=(256, 256), degree=(3, 3), n_samples=(8192, 8192), nout=3, float_type="Float32", nurbs=True),
}
WORKLOAD = WORKLOADS["C3"]
METRIC = "evaluate!+evaluate_adjoint! sample-values/sec"
UNIT = "sample-values/s"


def algorithmic_bytes(n_samples, n_cp, degree, nout, elem=8, nurbs=False):
    """output (or adjoint input) + control points (+ weights) + selected table slices + span indices."""
    n, c = int(np.prod(n_samples)), int(np.prod(n_cp))
    return (n * nout * elem + c * nout * elem + (c * elem if nurbs else 0) +
            sum(nd * ((p + 1) * elem + 4) for nd, p in zip(n_samples, degree)))


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=2.0):
        """nvidia-smi needs a few hundred ms to start: do not begin the timed region before it samples."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.lines and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self, t0=None, t1=None):
        """Summary of the samples received between host times t0 and t1 (the timed region); if the region was shorter
        than the sampling interval, the samples closest to it."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        stamped = list(self.lines)
        if t0 is not None and t1 is not None:
            inside = [ln for ts, ln in stamped if t0 <= ts <= t1 + 0.03]
            if not inside and stamped:               # region shorter than the interval: first sample after it began
                after = [ln for ts, ln in stamped if ts >= t0]
                inside = after[:1] if after else [stamped[-1][1]]
        else:
            inside = [ln for _, ln in stamped]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(np.max(power))}


# --------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (C/OpenMP restatement) on a bounded slab of the workload
# --------------------------------------------------------------------------------------------


def cpu_reference_setup(planes: int):
    from oracle import oracle_c as OC
    from oracle import oracle_np as O
    w = WORKLOAD
    npdt = np.float64 if w["float_type"] == "Float64" else np.float32
    dims = []
    for d, (c, p, n) in enumerate(zip(w["n_cp"], w["degree"], w["n_samples"])):
        kv, mu, ka = O.clamped_knot_vector(c, p, npdt)
        sp = O.default_sample_points(kv, n)
        if d == len(w["n_cp"]) - 1:
            sp = np.ascontiguousarray(sp[:planes])           # a slab of `planes` rows along the slowest axis
        idx = OC.span_indices(sp, ka, p)
        dims.append((OC.basis_tables(ka, sp, idx, p, 0), idx))
    rng = np.random.default_rng(1)
    cp = np.asfortranarray(rng.random(w["n_cp"] + (w["nout"],)).astype(npdt))
    wts = np.asfortranarray((0.5 + np.random.default_rng(2).random(w["n_cp"])).astype(npdt)) if w.get("nurbs") else None
    n_s = w["n_samples"][:-1] + (planes,)
    e_in = np.asfortranarray(np.random.default_rng(3).random(n_s + (w["nout"],)).astype(npdt))
    out = np.empty(n_s + (w["nout"],), dtype=npdt, order="F")
    g = np.empty(w["n_cp"] + (w["nout"],), dtype=npdt, order="F")
    tabs, idxs = [t for t, _ in dims], [i for _, i in dims]
    der0 = [0] * len(w["n_cp"])

    def step():
        OC.evaluate(tabs, idxs, w["degree"], der0, cp, wts, out)
        OC.evaluate_adjoint(tabs, idxs, w["degree"], der0, e_in, g.shape, wts, g)

    values = 2 * int(np.prod(n_s)) * w["nout"]
    return step, values, OC.max_threads()


try:
    _AFFINITY0 = set(os.sched_getaffinity(0))       # before any NUMA binding of the GPU arm
except AttributeError:
    _AFFINITY0 = set(range(os.cpu_count() or 1))


def host_cores() -> int:
    """Host threads this process may use (the CPU affinity mask at start-up; torchrun exports OMP_NUM_THREADS=1,
    which is a default for ITS workers, not a limit of the box)."""
    return max(1, len(_AFFINITY0))


def run_cpu_reference(steps: int, warmup: int, planes: int, budget_s: float = None):
    """Times `steps` steps (fewer if they would not fit `budget_s` seconds; the number run is returned)."""
    from oracle import oracle_c as OC
    try:
        os.sched_setaffinity(0, _AFFINITY0)         # undo the GPU arm's NUMA binding: the CPU arm uses the whole box
    except (AttributeError, OSError):
        pass
    OC.set_threads(host_cores())                    # regardless of an inherited OMP_NUM_THREADS
    step, values, cores = cpu_reference_setup(planes)
    t0 = time.perf_counter()
    for _ in range(max(1, warmup)):
        step()
    dt_est = (time.perf_counter() - t0) / max(1, warmup)
    if budget_s is not None:
        steps = max(1, min(steps, int(budget_s / max(dt_est, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return values / dt, dt, cores, steps


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    planes = min(args.cpu_planes, WORKLOAD["n_samples"][-1])
    warmup = max(1, args.warmup)
    # as many of the requested steps as fit ~150 s of CPU time (one step of the 32-plane sample takes ~0.5 s)
    value, dt, cores, steps = run_cpu_reference(args.steps, warmup, planes, budget_s=150.0)
    sample = (f"slab of {planes}/{WORKLOAD['n_samples'][-1]} rows along the slowest axis of the {args.config} grid (all control "
              f"points), evaluate!+adjoint, C/OpenMP restatement of the reference algorithm (Julia unavailable), {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if WORKLOAD["float_type"] == "Float64" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "cpu_sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this process (and therefore its pinned host buffers, first touch) to the NUMA node of its GPU: eight ranks
    whose staging buffers sit on the wrong socket share the inter-socket link.  Best effort; returns the node or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node")
        node = int(path.read_text().strip())
        if node < 0:
            return None
        cpus = []
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def check_exchange(S, sh, grid, e_in, grad, rank, world, dev):
    """N > 1 parity gate, always on: (1) the exchanged gradient against an NCCL all-reduce of the local partial
    gradients; (2) on rank 0, control planes 4..6 of the exchanged gradient against the C oracle's adjoint (the
    reference's algorithm, src/adjoint.jl:11-39) of the first 32 sample planes (those planes see no other sample)."""
    import torch
    import torch.distributed as dist
    S.evaluate_adjoint_(grid, eval=e_in, control_points=grad)
    ref = grad.clone()
    S.allreduce_gradient_(ref)
    grad.fill_(float("nan"))
    sh.evaluate_adjoint_(eval=e_in, control_points=grad)
    torch.cuda.synchronize()
    if sh.exchange is not None and sh.exchange.mode == "support":
        # support-plane exchange: the summed gradient is defined on the control planes this rank's slab reads
        lo, hi = sh.exchange.support_planes()
        sl = (Ellipsis, slice(lo, hi), slice(None))
        err = float((grad[sl] - ref[sl]).norm() / ref[sl].norm())
        untouched = bool(torch.isnan(grad[..., :lo, :]).all() and torch.isnan(grad[..., hi:, :]).all())
    else:
        err = float((grad - ref).norm() / ref.norm())
        untouched = None
    out = {"kind": sh.exchange_kind, "rel_err_vs_allreduce": err}
    if untouched is not None:
        out["support_planes"] = [lo, hi]
        out["planes_outside_support_untouched"] = untouched
    if sh.exchange is not None:
        ep, timed_out = sh.exchange.status()
        out["exchanges_completed"], out["timed_out"] = ep, timed_out
    if rank == 0:
        from oracle import oracle_c as OC
        OC.set_threads(host_cores())
        w = WORKLOAD
        n_pl = 32
        assert grid.eval.shape[2] >= n_pl
        tabs = [np.asfortranarray(S.to_numpy(sd.eval)) for sd in grid.spline_dimensions]
        idxs = [np.ascontiguousarray(S.to_numpy(sd.sample_indices)) for sd in grid.spline_dimensions]
        tabs[2], idxs[2] = np.asfortranarray(tabs[2][:n_pl]), np.ascontiguousarray(idxs[2][:n_pl])
        assert int(idxs[2][-1]) > 6 + 4                      # spans up to 10 (1-based) lie completely inside the sample
        e_np = np.asfortranarray(S.to_numpy(e_in[:, :, :n_pl, :]))
        gref = OC.evaluate_adjoint(tabs, idxs, list(w["degree"]), [0, 0, 0], e_np, tuple(w["n_cp"]) + (w["nout"],))
        got = S.to_numpy(grad)[:, :, 4:7, :]
        out["rel_err_vs_oracle_planes_4_6"] = float(np.linalg.norm(got - gref[:, :, 4:7, :]) / np.linalg.norm(gref[:, :, 4:7, :]))
    flag = torch.tensor([0 if (err <= 1e-12 and out.get("rel_err_vs_oracle_planes_4_6", 0.0) <= 1e-12
                               and not out.get("timed_out", False)) else 1], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    out["ok"] = int(flag.item()) == 0
    return out


def main_gpu(args):
    import torch
    import torch.distributed as dist

    # Libraries print to stdout while we run (NCCL's version banner, for one); the contract is ONE JSON line there, so
    # everything else is sent to stderr and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import __graft_entry__ as entry
    S = entry.load_package()                       # fails loudly if libsplinegrids_b200.so is missing
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    world_env = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world_env > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)

    w = WORKLOAD
    S.set_synchronous(False)                       # the benchmark synchronises once per timed region
    gdims = tuple(S.SplineDimension(c, p, n, float_type=w["float_type"])
                  for c, p, n in zip(w["n_cp"], w["degree"], w["n_samples"]))
    nurbs = bool(w.get("nurbs"))
    akw = {"allow_nurbs": True} if nurbs else {}
    elem = 8 if w["float_type"] == "Float64" else 4
    sh = S.SlabShardedGrid(gdims, w["nout"], rank, world, nurbs=nurbs, peer_exchange=(world > 1 and not args.nccl_allreduce),
                           exchange_mode=args.exchange)
    grid = sh.local
    if nurbs:
        gw = torch.Generator(device=dev)
        gw.manual_seed(2)
        grid.weights.copy_(0.5 + torch.rand(grid.weights.shape, dtype=grid.weights.dtype, device=dev, generator=gw))
    n_local = tuple(grid.eval.shape[:-1])
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    cp = grid.control_points.obtain()
    cp.copy_(torch.rand(cp.shape, dtype=cp.dtype, device=dev, generator=gen))        # replicated (same seed)
    gen.manual_seed(3 + rank)
    e_in = torch.empty_like(grid.eval)
    e_in.copy_(torch.rand(e_in.shape, dtype=e_in.dtype, device=dev, generator=gen))
    grad = torch.zeros_like(cp)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        S.evaluate_(grid)
        sh.evaluate_adjoint_(eval=e_in, control_points=grad, **akw)   # local adjoint + gradient exchange (N > 1)

    values_per_step = 2 * int(np.prod(w["n_samples"])) * w["nout"]      # whole job, both ops
    exchange_check = None
    if world > 1:
        exchange_check = check_exchange(S, sh, grid, e_in, grad, rank, world, dev)
        if not exchange_check["ok"]:
            if rank == 0:
                print(f"[bench] gradient exchange FAILED its parity gate: {exchange_check}", file=sys.stderr)
            dist.destroy_process_group()
            return 3

    # ---- kernel-only timing: K steps between two events, max over ranks ---------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    # The step is captured once in a CUDA graph (two steps per replay: the peer-memory exchange alternates between two
    # staging buffers) and replayed: the same kernels, without the per-call host cost that bounds thin slabs.
    cap, cap_err = None, None
    # N > 1: the peer-memory exchange is plain kernels with a device-side flag barrier, so the step is captured as well;
    # with the NCCL all-reduce (--nccl-allreduce) the step launches eagerly unless --graph forces the attempt.
    if (not args.no_graph and (world == 1 or sh.exchange is not None)) or args.graph:
        try:
            cap = S.CapturedCalls(step, unroll=2, warmup=1)
            for _ in range(2):
                cap.replay()
        except Exception as ex:                      # capture unsupported for some collective: eager launches
            cap, cap_err = None, repr(ex)[:200]
            torch.cuda.synchronize()
    ok = torch.tensor([1 if cap is not None else 0], device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # all ranks replay, or none
    if int(ok.item()) == 0:
        cap = None
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    S.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    if cap is not None:
        for _ in range(args.steps // 2):
            cap.replay()
        for _ in range(args.steps % 2):
            step()
    else:
        for _ in range(args.steps):
            step()
    e1.record()
    barrier()
    t_host1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    launches = S.launch_count() + (cap.kernel_launches * (args.steps // 2) if cap is not None else 0)
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = values_per_step / (ms_step * 1e-3)

    # ---- per-op and dominant-kernel timings (live CUDA events on the launch stream) ----------------
    def time_fn(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()                                        # keep the queue busy so launch overhead is hidden
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    reps = max(3, min(args.steps, 10))
    ms_fwd = time_fn(lambda: S.evaluate_(grid), reps)
    var_fwd = S.last_variant()
    ms_adj = time_fn(lambda: S.evaluate_adjoint_(grid, eval=e_in, control_points=grad, **akw), reps)
    var_adj = S.last_variant()
    peaks, peak_kind = load_peaks()
    hbm = float(peaks["hbm_gbs"])
    nb_local = algorithmic_bytes(n_local, w["n_cp"], w["degree"], w["nout"], elem, nurbs)
    n_local_values = int(np.prod(n_local)) * w["nout"]
    ops = {
        "evaluate": {"ms": ms_fwd, "values_per_s": n_local_values / (ms_fwd * 1e-3), "variant": var_fwd,
                     "achieved_GBs": nb_local / (ms_fwd * 1e-3) / 1e9, "frac_of_hbm_roofline": nb_local / (ms_fwd * 1e-3) / 1e9 / hbm},
        "evaluate_adjoint": {"ms": ms_adj, "values_per_s": n_local_values / (ms_adj * 1e-3), "variant": var_adj,
                             "achieved_GBs": nb_local / (ms_adj * 1e-3) / 1e9,
                             "frac_of_hbm_roofline": nb_local / (ms_adj * 1e-3) / 1e9 / hbm,
                             "note": "local kernels only (no all-reduce)"},
    }
    # dominant kernel of the step (~42 % of it, profiles/r02_ncu_launches_bench_step.csv): the TMA-fed double march of the
    # adjoint.  Its launch duration is measured live with CUDA events recorded by the library on the launch stream
    # right before / after the kernel (sg_profile_adjoint_main, include/splinegrids_b200.h).
    ms_main = None
    if var_adj.startswith("adjoint_march2"):
        lib = S._lib.lib()
        lib.sg_profile_adjoint_main(1)
        samples = []
        for _ in range(reps):
            S.evaluate_adjoint_(grid, eval=e_in, control_points=grad, **akw)
            t = float(lib.sg_profile_adjoint_main_ms())
            if t > 0:
                samples.append(t)
        lib.sg_profile_adjoint_main(0)
        if samples:
            ms_main = float(np.mean(samples))
    # roofline kernel = the forward march kernel (ONE launch == the whole evaluate! call, so its CUDA-event time is
    # the kernel's launch duration).  The adjoint is a chain of kernels; its op-level fraction is in "also"/"ops".
    fwd_roof = {"kernel": (f"sg_eval3d_march_kernel<double,3,2,4,4,TMA={'true' if var_fwd.endswith('tma') else 'false'}> "
                           if args.config == "C3" else "") + f"[{var_fwd}] (evaluate!, one launch per call)",
                "share_of_step": ms_fwd / (ms_fwd + ms_adj),
                "bound": "hbm", "achieved": ops["evaluate"]["achieved_GBs"], "peak": hbm, "unit": "GB/s",
                "frac": ops["evaluate"]["frac_of_hbm_roofline"], "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                "algorithmic_bytes_per_launch": nb_local, "traffic": args.traffic_bytes if args.config == "C3" else None,
                "traffic_source": "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                                  "N=1 full grid)" if args.traffic_bytes else None}
    if ms_main is not None:
        # algorithmic bytes of this launch: the sample array read once + the tables/indices (SURVEY 8d's adjoint
        # figure minus the control-point write, which the post kernel does)
        nb_main = nb_local - int(np.prod(w["n_cp"])) * w["nout"] * elem
        tr_main = None
        tp = ROOT / "profiles" / "r02_traffic.json"
        if tp.exists() and world == 1:
            k = json.loads(tp.read_text()).get("adjoint", {}).get("sg_adj_march2_tma_kernel")
            if k:
                tr_main = float(k["read"] + k["write"])
        roofline = {"kernel": "sg_adj_march2_tma_kernel<double,3,4,20,3> (evaluate_adjoint!: TMA-fed double march over the sample "
                              "array; dominant kernel of the step)",
                    "share_of_step": ms_main / (ms_fwd + ms_adj), "bound": "hbm",
                    "achieved": nb_main / (ms_main * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                    "frac": nb_main / (ms_main * 1e-3) / 1e9 / hbm, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                    "launch_ms": ms_main, "algorithmic_bytes_per_launch": nb_main, "traffic": tr_main,
                    "traffic_source": "profiles/r02_traffic.json (ncu --set full at HEAD, N=1 full grid: the sample array read once + the "
                                      "part of the 135 MB of evict-last partials that reaches HBM before the kernel ends)" if tr_main else None,
                    "timing": "CUDA events recorded by the library on the launch stream around this kernel",
                    "also": {"evaluate_adjoint_op_frac": ops["evaluate_adjoint"]["frac_of_hbm_roofline"],
                             "evaluate_kernel": fwd_roof}}
    else:
        roofline = dict(fwd_roof, also={"evaluate_adjoint_op_frac": ops["evaluate_adjoint"]["frac_of_hbm_roofline"]})

    # ---- end-to-end through the public API with HOST buffers -------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    flat = lambda t: t.permute(*reversed(range(t.dim()))).reshape(-1)       # contiguous view of a column-major array
    h_cp = torch.empty(cp.numel(), dtype=cp.dtype).pin_memory()
    h_cp.copy_(flat(cp).cpu())
    h_ein = torch.empty(e_in.numel(), dtype=e_in.dtype).pin_memory()
    h_ein.copy_(flat(e_in).cpu())
    h_eval = torch.empty(grid.eval.numel(), dtype=cp.dtype).pin_memory()
    h_grad = torch.empty(cp.numel(), dtype=cp.dtype).pin_memory()

    # Three streams: uploads, compute, downloads.  PCIe is full duplex, so the upload of the adjoint's input overlaps
    # the forward kernel and the download of its result; every byte still crosses the bus inside the timed region.
    s_up, s_down, s_main = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()

    def e2e_step():
        ev_cp, ev_ein, ev_fwd, ev_adj = (torch.cuda.Event() for _ in range(4))
        with torch.cuda.stream(s_up):
            flat(cp).copy_(h_cp, non_blocking=True)             # H2D control points
            ev_cp.record(s_up)
            flat(e_in).copy_(h_ein, non_blocking=True)          # H2D adjoint input
            ev_ein.record(s_up)
        s_main.wait_event(ev_cp)
        S.evaluate_(grid)
        ev_fwd.record(s_main)
        s_main.wait_event(ev_ein)
        sh.evaluate_adjoint_(eval=e_in, control_points=grad, **akw)    # (+ gradient exchange for N > 1)
        ev_adj.record(s_main)
        with torch.cuda.stream(s_down):
            s_down.wait_event(ev_fwd)
            h_eval.copy_(flat(grid.eval), non_blocking=True)    # D2H evaluated grid
            s_down.wait_event(ev_adj)
            h_grad.copy_(flat(grad), non_blocking=True)         # D2H gradient
        s_down.synchronize()
        s_main.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    dt = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": values_per_step / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": e2e_steps,
           "h2d_bytes_per_step": int((cp.numel() + e_in.numel()) * elem),
           "d2h_bytes_per_step": int((grid.eval.numel() + cp.numel()) * elem),
           "numa_node": numa_node, "note": "per rank; pinned host buffers (process bound to the GPU's NUMA node); uploads, kernels and downloads on three streams (full-duplex PCIe), all "
                   "copies + kernels + syncs inside the timed region"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpl = min(args.cpu_planes, w["n_samples"][-1])
        v, dts, cores, _ = run_cpu_reference(steps=8, warmup=1, planes=cpl, budget_s=15.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_step_sample": dts * 1e3,
               "sample": f"slab of {cpl}/{w['n_samples'][-1]} rows along the slowest axis of the same grid, evaluate!+adjoint, C/OpenMP restatement "
                         f"of the reference algorithm (Julia not installed, the reference package cannot run)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64" if elem == 8 else "f32", "data": "synthetic",
                "config": {"workload": w["name"], "step": "evaluate! + evaluate_adjoint! (+ NCCL all-reduce of the gradient for N>1)",
                           "sharding": f"sample grid in {world} slab(s) along axis 3, control points replicated",
                           "gradient_exchange": sh.exchange_kind, "exchange_check": exchange_check,
                           "exchange_mode": (None if world == 1 else ("nccl_allreduce" if sh.exchange is None else sh.exchange.mode)),
                           "exchange_note": (None if world == 1 or sh.exchange is None or sh.exchange.mode != "support" else
                                             "every rank ends the step with the summed gradient on the control planes its own slab "
                                             "reads (what the next evaluate! needs); --exchange replicated gives every rank the whole "
                                             "gradient (profiles/r02_bench_n{2,4,8}.json: 0.255 / 0.159 / 0.115 ms per step)"),
                           "nvls_multicast_push": bool(sh.exchange is not None and any(sh.exchange.mc_ptrs)),
                           "launch": ("CUDA graph of the step's kernels (2 steps per replay)" if cap is not None
                                      else "eager" + (f" (graph capture failed: {cap_err})" if cap_err else "")),
                           "l2": (f"inputs/outputs ({nb_local / 1e9:.2f} GB per op) exceed the 126 MB L2; no flush needed" if nb_local > 150e6
                                  else "working set fits the 126 MB L2: warm-cache numbers (launch-latency regime)")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "ops": ops,
                "cpu_baseline": cpu}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(WORKLOADS), help="BASELINE config (default C3, the one the metric is "
                    "quoted on; the others run at N = 1 only and give one line per config for profiles/)")
    ap.add_argument("--cpu-planes", type=int, default=32, help="slab thickness of the bounded CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-allreduce", action="store_true", help="N>1: use the NCCL all-reduce instead of the peer-memory exchange")
    ap.add_argument("--exchange", choices=["replicated", "support"], default=os.environ.get("SG_BENCH_EXCHANGE", "support"),
                    help="N>1 gradient exchange: 'replicated' = every rank ends with the whole summed gradient; 'support' = every "
                         "rank ends with the summed gradient on the control planes its slab reads (halo exchange with the "
                         "neighbouring ranks only)")
    ap.add_argument("--graph", action="store_true", help="N>1: also try to capture the step (with its gradient exchange) in a CUDA graph")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--check", action="store_true", help="(kept for compatibility: the N>1 exchange check always runs)")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.config != "C3":
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            ap.error("--config other than C3 runs at N = 1 only")
        global WORKLOAD
        WORKLOAD = WORKLOADS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.traffic_bytes is None:     # DRAM traffic of the roofline kernel from the committed ncu capture
        tp = ROOT / "profiles" / "r02_traffic.json"
        if tp.exists():
            args.traffic_bytes = float(json.loads(tp.read_text())["total"])
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
