"""CUDA-graph replay of the inference forward for the small, launch-bound configurations.

A 3840x2160 forward is ~400 kernel launches of tens to thousands of microseconds: the host is never the
limit.  At the LOL sizes (400x600, BASELINE configs[1]) the same ~400 launches are a few microseconds
each and the Python / launch overhead dominates the step.  ``GraphedForward`` captures one forward of a
fixed input shape into a ``torch.cuda.CUDAGraph`` and replays it: one launch per image, no Python between
the kernels.  Everything this package enqueues is capturable (kernel launches on the current stream;
workspaces and outputs come from torch's caching allocator, which gives a captured graph its own pool;
the TMA descriptors are kernel parameters, frozen at capture together with the buffer addresses they
point at).

    g = GraphedForward(net, (4, 3, 400, 600))
    y = g(x)          # x is copied into the graph's input buffer; y is the graph's output buffer

The output tensor is reused by the next call: clone it if it must survive.  Inference only
(``torch.no_grad``); parameters are read at replay time, so loading new weights into the same tensors
is picked up, except by the dense 3x3 convolutions, whose prepacked weights are cached per parameter
version -- call ``refresh()`` after changing weights.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops


class GraphedForward:
    def __init__(self, net: torch.nn.Module, shape: Sequence[int], warmup: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedForward needs a CUDA device")
        self.net = getattr(net, "restoration_network", net)
        self.device = next(self.net.parameters()).device
        self.shape = tuple(int(s) for s in shape)
        self.warmup = max(1, int(warmup))
        self._x = torch.zeros(self.shape, device=self.device, dtype=torch.float32)
        self._capture()

    @torch.no_grad()
    def _capture(self):
        # warm-up on a side stream (torch's capture protocol): first-call work that cannot be captured
        # (weight prepacking, lazily allocated device words, cudaFuncSetAttribute) happens here
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.net(self._x)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._y = self.net(self._x)

    def refresh(self):
        """Re-capture (after the parameters changed behind the prepack cache)."""
        ops.clear_pack_cache()
        self._capture()

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape) != self.shape:
            raise ValueError(f"GraphedForward was captured for {self.shape}, got {tuple(x.shape)}")
        self._x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self._y
