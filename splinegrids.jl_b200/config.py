"""Run-time switches of the host mirror."""
import contextlib

import torch

# The reference calls `synchronize(backend)` after every kernel launch (20 call sites in src/).  The host
# mirror reproduces that by default; optimisation loops and the benchmark switch it off and
# synchronise once themselves.
_synchronous = True


def set_synchronous(flag: bool) -> None:
    global _synchronous
    _synchronous = bool(flag)


def is_synchronous() -> bool:
    return _synchronous


@contextlib.contextmanager
def asynchronous():
    """``with asynchronous(): ...`` -- launch without the per-call stream synchronisation."""
    global _synchronous
    old = _synchronous
    _synchronous = False
    try:
        yield
    finally:
        _synchronous = old


def after_launch(device=None) -> None:
    if _synchronous:
        torch.cuda.current_stream(device).synchronize()


# evaluate_adjoint! through an adjoint plan (sg_adjoint_plan_*): the inverse sample map is built once per grid and
# derivative order instead of on every call, and exactly the kernels that are needed are launched.  Switch off to
# exercise the plain sg_evaluate_adjoint entry point (all decisions on device, fallback kernels launched).
_adjoint_plans = True


def set_adjoint_plans(flag: bool) -> None:
    global _adjoint_plans
    _adjoint_plans = bool(flag)


def adjoint_plans() -> bool:
    return _adjoint_plans
