"""The reference's training loss as two fused kernels.

`compute_loss(y_pred, y, mask=None, use_mask=True)` has the signature and the arithmetic of the reference's
`compute_loss` (main.py:28-72) -- weighted L1 with weight 1 + 4|y|^3 plus 0.005 x the spatial-gradient loss,
masked means with the 1e-8 epsilon when a mask is given -- but runs as one reduction pass forward
(b200_wl1_grad_loss_fwd) and one pass backward (b200_wl1_grad_loss_bwd) over the prediction / target / mask
maps instead of ~25 + ~40 element-wise PyTorch launches.  CUDA tensors only (no CPU fallback).
"""
from __future__ import annotations

import torch

from . import _lib
from .ops import _p, _st


class _WeightedL1GradLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pred, y, mask):
        for t, name in ((y_pred, "y_pred"), (y, "y"), (mask, "mask")):
            if t is not None and not t.is_cuda:
                raise RuntimeError(f"compute_loss: {name} is on {t.device}; the B200 kernels have no CPU fallback")
        if y_pred.shape != y.shape or (mask is not None and mask.shape != y.shape):
            raise ValueError(f"compute_loss: shapes differ: {tuple(y_pred.shape)} / {tuple(y.shape)}"
                             + ("" if mask is None else f" / {tuple(mask.shape)}"))
        if y_pred.dim() < 2:
            raise ValueError("compute_loss: expected [..., H, W] maps")
        yp = y_pred.detach().float().contiguous()
        yt = y.detach().float().contiguous()
        mk = None if mask is None else mask.detach().float().contiguous()
        H, W = yp.shape[-2], yp.shape[-1]
        img = yp.numel() // (H * W)
        sums = torch.empty(6, device=yp.device, dtype=torch.float64)
        loss = torch.empty((), device=yp.device, dtype=torch.float32)
        _lib.call("b200_wl1_grad_loss_fwd", _p(yp), _p(yt), _p(mk), img, H, W, _p(sums), _p(loss), _st())
        ctx.save_for_backward(yp, yt, mk, sums) if mk is not None else ctx.save_for_backward(yp, yt, sums)
        ctx.has_mask = mk is not None
        ctx.in_dtype = y_pred.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.has_mask:
            yp, yt, mk, sums = ctx.saved_tensors
        else:
            (yp, yt, sums), mk = ctx.saved_tensors, None
        H, W = yp.shape[-2], yp.shape[-1]
        img = yp.numel() // (H * W)
        go = grad_out.detach().float().contiguous()
        d = torch.empty_like(yp)
        _lib.call("b200_wl1_grad_loss_bwd", _p(yp), _p(yt), _p(mk), img, H, W, _p(sums), _p(go), _p(d), _st())
        return d.to(ctx.in_dtype), None, None


def compute_loss(y_pred: torch.Tensor, y: torch.Tensor, mask: torch.Tensor | None = None, use_mask: bool = True):
    """Drop-in for the reference's compute_loss (main.py:28): returns a scalar fp32 loss tensor."""
    return _WeightedL1GradLoss.apply(y_pred, y, mask if (use_mask and mask is not None) else None)
