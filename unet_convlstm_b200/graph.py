"""One training step (forward + backward + optimizer update) captured into a CUDA graph.

The step of the UNet-ConvLSTM launches ~750 kernels, many of them microseconds long (BatchNorm
finalisation, weight re-packing, the per-step BPTT launches), so the eager host path -- Python, autograd
bookkeeping, ctypes -- leaves ~3 % of the step as launch gaps.  Shapes are static in training, so the whole
step is captured once and replayed: the GPU sees one graph launch per step.

Everything on the path is capturable: the C-ABI launches run on torch's current stream, take their
TMA descriptors by value and never synchronise or allocate; torch's caching allocator serves the
temporaries from the graph's private pool; the packed weight copies are re-derived inside the graph
every step (the parameters change under the optimizer), which is what the eager path does as well.
"""
from __future__ import annotations

import torch

from . import _lib, ops


class GraphedTrainStep:
    """loss = step(x, y): copies the batch into static device buffers and replays the captured step.

    model     : train.unet.TemporalUNetDualView (or any module built from this package) on the GPU
    optimizer : a capturable optimizer, e.g. torch.optim.AdamW(..., fused=True, capturable=True)
    loss_fn   : (list_of_frames, y) -> scalar loss tensor
    x, y      : example batch (device tensors) fixing the shapes
    """

    def __init__(self, model, optimizer, loss_fn, x, y, warmup: int = 3, with_optimizer: bool = True):
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.x = x.clone()
        self.y = y.clone()
        self.with_optimizer = with_optimizer
        saved = (ops.CELL_TIMER, _lib.TIMER)
        saved_bg, ops.WGRAD_STREAM = ops.WGRAD_STREAM, False  # one stream inside the capture
        ops.CELL_TIMER, _lib.TIMER = None, None  # CUDA events cannot be timed inside a capture
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            optimizer.zero_grad(set_to_none=True)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._eager_step(zero=False)
        finally:
            ops.CELL_TIMER, _lib.TIMER = saved
            ops.WGRAD_STREAM = saved_bg
        if _lib.lib().b200_device_error() != 0:
            raise RuntimeError("device watchdog flag set during graph capture")

    def _eager_step(self, zero: bool = True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        out, _ = self.model(self.x)
        loss = self.loss_fn(out, self.y)
        loss.backward()
        if self.with_optimizer:
            self.optimizer.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor | None = None, y: torch.Tensor | None = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y is not None and y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
