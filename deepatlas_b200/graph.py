"""Whole-step CUDA graph: one training step (zero the gradient bucket, forward, backward, gradient all-reduce, optimizer)
captured once and replayed, so that a step costs one graph launch instead of about a thousand kernel launches through
Python (autograd functions, ctypes calls, allocator).  At 160x192x160 the eager host side of a joint step is about as
long as its GPU side (47 ms), so without the graph any further kernel-side gain would be hidden behind Python.

Everything the step touches has a fixed address: the caller's inputs are copied into static input tensors (device to
device), all intermediate tensors and workspaces come from the graph's private memory pool, parameters / optimizer state /
BatchNorm buffers are updated in place, and the library's kernels take only raw pointers, extents and the capturing stream
(TMA tensor maps are kernel arguments, encoded at capture time from those fixed addresses).  Shapes are static: one
GraphedStep per input shape.  The optimizer must be capturable (``torch.optim.Adam(..., fused=True, capturable=True)``).
"""
from __future__ import annotations

import os
from typing import Callable, Sequence

import torch

from . import _lib


class GraphedStep:
    """``step_fn(*inputs) -> loss`` captured as one CUDA graph.

    ``step_fn`` must be a pure device-side step (no ``.item()``, no host-dependent control flow) that leaves its result in
    the returned tensor.  ``warmup`` eager runs come first: they create the lazily initialised state (optimizer moments,
    kernel attributes, cuDNN-free here, NCCL communicators) that must not be created during capture.  Call the object
    with the step's inputs (device tensors of the captured shapes); it returns the static loss tensor of the replay."""

    def __init__(self, step_fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 eager_tail: Callable | None = None):
        """``eager_tail``: an optional part of the step that runs eagerly after every replay (and after every warm-up
        run).  Multi-process training puts the NCCL gradient all-reduce and the optimizer there: a collective captured in
        a graph leaves torch's NCCL watchdog waiting on work it never sees complete (observed: the watchdog aborts the
        process at teardown, 480 s later)."""
        if not example_inputs or not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("deepatlas_b200: GraphedStep needs CUDA example inputs (no CPU path exists)")
        self.device = example_inputs[0].device
        self.static_inputs = [t.clone() for t in example_inputs]
        # DA_GRAPH_PRIORITY=1 captures the step on a HIGH-priority stream (the priority is recorded in the graph's kernel
        # nodes, so the chain the step was written on would get free SMs before the branches forked onto side streams).
        prio = os.environ.get("DA_GRAPH_PRIORITY", "0") == "1"   # measured: 41.84 vs 41.46 ms without -- off
        side = torch.cuda.Stream(device=self.device, priority=-1 if prio else 0)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                step_fn(*self.static_inputs)
                if eager_tail is not None:
                    eager_tail()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.size("da_launch_count")
        # thread_local: the NCCL watchdog thread polls events while the collective of the step is being captured
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self.static_loss = step_fn(*self.static_inputs)
        self.launches_per_step = _lib.size("da_launch_count") - l0   # library kernels inside one replay
        self.eager_tail = eager_tail
        self.replays = 0

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if len(inputs) != len(self.static_inputs):
            raise ValueError(f"GraphedStep: {len(inputs)} inputs for a step captured with {len(self.static_inputs)}")
        for dst, src in zip(self.static_inputs, inputs):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"GraphedStep: input {tuple(src.shape)} {src.dtype} does not match the captured "
                                 f"{tuple(dst.shape)} {dst.dtype} (one graph per shape)")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.eager_tail is not None:
            self.eager_tail()
        self.replays += 1
        return self.static_loss
