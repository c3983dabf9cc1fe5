"""CUDA-graph capture of a sequence of library calls (the ``CUDA.@captured`` idiom of a Julia host).

A fitting loop calls ``evaluate!`` and ``evaluate_adjoint!`` with the same arrays thousands of times
(ext/SplineGridsLinearMapsExt.jl:16-48).  Every C-ABI entry point only enqueues kernels and memset nodes on the
stream it is given -- no allocation, no host synchronisation, grid descriptions passed by value -- so a whole
iteration (including the multi-GPU gradient exchange) can be captured once and replayed with ONE launch.  That
removes the per-call host cost (argument marshalling, ~6 kernel launches per adjoint), which is what bounds a
slab-sharded step on 8 GPUs where the kernels take only tens of microseconds.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib
from .config import asynchronous


class CapturedCalls:
    """Capture ``fn()`` (any sequence of ``evaluate_`` / ``evaluate_adjoint_`` / exchange calls on fixed arrays)
    into a CUDA graph; ``replay()`` enqueues the whole sequence on the current stream.

    ``unroll``: how many consecutive ``fn()`` calls one replay stands for (the peer-memory gradient exchange
    alternates between two staging buffers, so its steps are captured in pairs).
    ``kernel_launches``: kernels of this library per replay (counted at capture time).
    """

    def __init__(self, fn: Callable[[], None], unroll: int = 1, warmup: int = 2, device: Optional[torch.device] = None):
        assert unroll >= 1
        self.unroll = int(unroll)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with asynchronous(), torch.cuda.stream(side):
            for _ in range(max(1, warmup) * self.unroll):     # workspaces allocated, attributes set, before the capture
                fn()
        side.synchronize()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with asynchronous(), torch.cuda.graph(self.graph, stream=side):
            for _ in range(self.unroll):
                fn()
        self.kernel_launches = _lib.launch_count() - before

    def replay(self) -> None:
        self.graph.replay()
