"""CUDA-graph replay of a SOM step.

A step of the hot path through the module API (``forward -> compute_weights -> som_loss -> backward``) is 5 kernel
launches and one memset with no host-side data dependence; issued eagerly it costs 250-300 us of Python / autograd /
ctypes time per step, more than the GPU needs.  ``StepGraph`` captures the caller's step function once (on static
input buffers) and replays it: the step then costs one graph launch.  This is host plumbing only - the captured
work is exactly the launches of libsom_b200.so that the eager call sequence makes.
"""
from __future__ import annotations

import torch


class StepGraph:
    """``g = StepGraph(step_fn)``; ``out = g.replay()``.

    ``step_fn()`` must read its inputs from tensors that stay allocated (copy new data INTO them before a replay)
    and return a tensor (or tuple of tensors); the returned objects are the static outputs of every replay.
    Gradients written by the step (``x.grad``, ``prototypes.grad``) live in graph-owned memory as well: read them
    after ``replay()`` on the same stream.

    Prototype updates between replays are honoured: under capture ``SOMLayer`` always puts the staging of the prototypes
    INTO the captured step (it re-reads the parameter on every replay), or reads the buffer that
    ``FusedPrototypeAdamW`` rewrites in place - a replay never computes with the prototypes of capture time.  (An
    optimizer that REPLACES ``layer.prototypes`` by a new tensor, instead of updating it in place, needs a re-capture,
    as with any CUDA graph.)"""

    def __init__(self, step_fn, warmup: int = 2, stream: torch.cuda.Stream | None = None):
        self.stream = stream if stream is not None else torch.cuda.current_stream()
        with torch.cuda.stream(self.stream):
            for _ in range(max(warmup, 1)):       # allocator warm-up and lazy initialisation outside the capture
                step_fn()
        self.stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.outputs = step_fn()

    def replay(self):
        self.graph.replay()
        return self.outputs
