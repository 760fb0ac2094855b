"""Data-parallel training over the GPUs of one box: one process per GPU, the batch of sequences is
sharded across ranks, every rank holds a full model replica, and the only
collective is the gradient all-reduce (sum, then 1/N) over NCCL / NVLink.

The reducer launches the all-reduce of a bucket as soon as every gradient in it is final, so the
transfers overlap the rest of backward.  Buckets are filled in reverse parameter order, which is the
order gradients become final (decoder -> ConvLSTM cells -> encoder).
BatchNorm statistics stay local to each replica (the reference has no SyncBN).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


class GradReducer:
    def __init__(self, params, bucket_bytes: int = 64 << 20, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []  # list of (flat buffer, [(param, offset, numel)])
        self._bucket_of = {}
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self._add_bucket(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._add_bucket(cur)
        self._pending = [0] * len(self.buckets)
        self._works = []
        ops.enable_background_wgrad()  # _on_grad below orders its reads after the background stream
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        for p in self.params:
            p._b200_bg_aware = True  # _on_grad orders its reads after the background stream (ops.background)
        self.reset()

    def _add_bucket(self, plist):
        total = sum(p.numel() for p in plist)
        flat = torch.zeros(total, device=plist[0].device, dtype=plist[0].dtype)
        entries, off = [], 0
        for p in plist:
            entries.append((p, off, p.numel()))
            self._bucket_of[p] = len(self.buckets)
            off += p.numel()
        self.buckets.append((flat, entries))

    def reset(self):
        self._pending = [len(e) for _, e in self.buckets]
        self._works = []

    def _on_grad(self, p):
        if p.is_cuda and ops.WGRAD_STREAM:
            # Weight gradients may still be in flight on the background stream (ops.background).  The bucket copy and
            # the all-reduce are queued THERE -- after the gradient's producer, and after the current stream for the
            # gradients made on it -- so the current stream never waits for a weight gradient during backward.
            cur, side = torch.cuda.current_stream(), ops.background_stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._bucket_grad(p, side)
            return
        self._bucket_grad(p, None)

    def _bucket_grad(self, p, side):
        b = self._bucket_of[p]
        flat, entries = self.buckets[b]
        for q, off, n in entries:
            if q is p:
                view = flat[off:off + n].view_as(p)
                if p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                    if side is not None:
                        p.grad.record_stream(side)
                    p.grad = view  # the gradient lives in the bucket from now on
                break
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Waits for the outstanding all-reduces and turns sums into means.  Call after backward."""
        if ops.WGRAD_STREAM and self.params and self.params[0].is_cuda:
            ops.background_join()
        for w in self._works:
            w.wait()
        if self.world > 1:
            inv = 1.0 / self.world
            for flat, _ in self.buckets:
                flat.mul_(inv)
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
        for p in self.params:
            p._b200_bg_aware = False
