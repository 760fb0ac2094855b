"""Host -> device input pipeline of the training loop: the reference copies
every batch synchronously on the compute stream (main.py:87-91, `x = x.to(device)`); here the copy of
batch i+1 runs on a side stream from pinned host memory while batch i computes."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    """Double-buffered asynchronous host->device copies.

        pf = DevicePrefetcher(device)
        pf.start(x_host, y_host)            # pinned host tensors
        for ...:
            x, y = pf.get()                 # device tensors of the batch started last
            pf.start(next_x_host, next_y_host)
            step(x, y)
    """

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None

    def start(self, *host_tensors):
        for t in host_tensors:
            if not t.is_pinned():
                raise RuntimeError("DevicePrefetcher: host tensors must be in pinned memory (tensor.pin_memory())")
        with torch.cuda.stream(self.stream):
            dev = [t.to(self.device, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (dev, ev)

    def get(self):
        if self._pending is None:
            raise RuntimeError("DevicePrefetcher.get() without a started copy")
        dev, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            t.record_stream(cur)  # the caching allocator must not recycle the buffer while `cur` uses it
        return dev
