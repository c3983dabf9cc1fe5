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
