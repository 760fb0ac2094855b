"""Gradient clipping + AdamW of the reference's training step as multi-tensor kernels.

    clip_grad_norm_(parameters, max_norm)        main.py:106  torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    AdamW(params, lr=..., weight_decay=...)      main.py:275  torch.optim.AdamW(...)

Both keep the reference calls' meaning (L2 norm over all gradients, coefficient max_norm / (norm + 1e-6) clamped to
1; decoupled weight decay, bias-corrected moments, state keys `step` / `exp_avg` / `exp_avg_sq` so a state_dict moves
between this class and torch.optim.AdamW).  `AdamW.step(clip_max_norm=1.0)` does both at once: one pass over the
gradients for the norm and one pass over {param, grad, exp_avg, exp_avg_sq}, the clip coefficient never leaving the
device.  fp32 CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from . import functional as Fn
from .ops import _p, _st, bump_version

# AdamW.step() emits the packed GEMM-operand copies of the convolution weights from its update kernel (b200_adamw_pack)
# for every packed copy a functional.WeightCache currently holds; B200_ADAMW_PACK=0: the weights are re-packed at their
# first use in the next forward pass instead (b200_pack_weight).
ADAMW_PACK = os.environ.get("B200_ADAMW_PACK", "1") != "0"


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _numel_array(tensors):
    return (ctypes.c_longlong * len(tensors))(*[t.numel() for t in tensors])


MT_MAX = 48  # tensors per kernel launch (csrc/optimizer.cu): calls are split so that one call = one launch
AP_MAX = 16  # weights per launch of the update-and-pack tile kernel


def _chunks(seq):
    return [seq[i:i + MT_MAX] for i in range(0, len(seq), MT_MAX)]


def _check(tensors, what):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(f"{what}: tensor on {t.device}; the B200 kernels have no CPU fallback")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"{what}: expected contiguous fp32 tensors, got {t.dtype}, strides {t.stride()}")


def grad_sqnorm(grads, out=None):
    """Sum of squares of all gradients as one fp64 device scalar (no host synchronisation)."""
    _check(grads, "grad_sqnorm")
    if out is None:
        out = torch.empty((), device=grads[0].device, dtype=torch.float64)
    if len(grads) <= MT_MAX:
        _lib.call("b200_grad_sqnorm_multi", len(grads), _ptr_array(grads), _numel_array(grads), _p(out), _st())
        return out
    parts = torch.empty(len(_chunks(grads)), device=out.device, dtype=torch.float64)
    for i, gs in enumerate(_chunks(grads)):
        _lib.call("b200_grad_sqnorm_multi", len(gs), _ptr_array(gs), _numel_array(gs), _p(parts[i]), _st())
    torch.sum(parts, dim=0, out=out)
    return out


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """Drop-in for torch.nn.utils.clip_grad_norm_ (L2 only): scales the gradients in place and returns the total
    norm as a device tensor."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("clip_grad_norm_: only the L2 norm (the reference's call, main.py:106)")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros(())
    sq = grad_sqnorm(grads)
    for gs in _chunks(grads):
        _lib.call("b200_grad_clip_multi", len(gs), _ptr_array(gs), _numel_array(gs), _p(sq), float(max_norm), _st())
    bump_version(*grads)
    return sq.sqrt().float()


class AdamW(torch.optim.Optimizer):
    """torch.optim.AdamW arithmetic (amsgrad = False, maximize = False) in one launch per 48 parameter tensors."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("AdamW: amsgrad / maximize are not used by the reference and not implemented")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("AdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False))
        self._sq = None

    @torch.no_grad()
    def step(self, closure=None, clip_max_norm=None):
        """One update.  clip_max_norm: fold clip_grad_norm_(all parameters of this optimizer, clip_max_norm) into the
        update (the stored gradients are left unscaled).  Returns the closure's loss, like torch."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            gs = [p.grad for p in ps]
            _check(ps, "AdamW.step (params)"), _check(gs, "AdamW.step (grads)")
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            groups.append((group, ps, gs))
        if not groups:
            return loss
        packs, redo, entries = ({}, [], [])
        if ADAMW_PACK:
            packs, redo, entries = Fn.pack_refresh_plan({p.data_ptr(): p for _, ps, _ in groups for p in ps})
        sq = None
        if clip_max_norm is not None:
            allg = [g for _, _, gs in groups for g in gs]
            if self._sq is None or self._sq.device != allg[0].device:
                self._sq = torch.empty((), device=allg[0].device, dtype=torch.float64)
            sq = grad_sqnorm(allg, self._sq)
        for group, ps, gs in groups:
            # parameters of one group may be at different step counts (added later): one call per count
            by_step = {}
            for p, g in zip(ps, gs):
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault(int(st["step"]), []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            b1, b2 = group["betas"]
            hyper = (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
            for t, its in by_step.items():
                # weights with packed copies in a WeightCache: tile kernel that updates AND re-packs, 16 weights per launch
                rest, packed = [], []
                for it in its:
                    geoms = packs.get(it[0].data_ptr())
                    if geoms:
                        _check(it[2:], "AdamW.step (moments)")
                        packed.append((it, geoms))
                    else:
                        rest.append(it)
                for i in range(0, len(packed), AP_MAX):
                    self._update_and_pack(packed[i:i + AP_MAX], hyper, t, sq, clip_max_norm)
                for items in _chunks(rest):
                    P, G, M, V = ([it[k] for it in items] for k in range(4))
                    _check(M, "AdamW.step (exp_avg)"), _check(V, "AdamW.step (exp_avg_sq)")
                    _lib.call("b200_adamw_multi", len(P), _ptr_array(P), _ptr_array(G), _ptr_array(M), _ptr_array(V),
                              _numel_array(P), *hyper, t, _p(sq), float(clip_max_norm or 0.0), _st())
                    bump_version(*P, *M, *V)  # written through raw pointers: packed-weight caches key on the version
        for fn in redo:
            fn()
        Fn.pack_refresh_commit(entries)
        return loss

    @staticmethod
    def _update_and_pack(items, hyper, t, sq, clip_max_norm):
        """b200_adamw_pack_multi on up to AP_MAX weights: the update plus the first two packed copies of each; further
        copies (none in this package's models) through b200_pack_weight."""
        n = len(items)
        P, G, M, V = ([it[k] for it, _ in items] for k in range(4))
        LL = ctypes.c_longlong
        dims = (LL * (3 * n))(*[x for _, geoms in items for x in geoms[0][:3]])
        none = (None, 0, 0, 0, 0, 1, 0, 0)
        d0 = [geoms[0][3:] for _, geoms in items]
        d1 = [geoms[1][3:] if len(geoms) > 1 else none for _, geoms in items]
        ptrs = lambda ds: (ctypes.c_void_p * n)(*[None if d[0] is None else d[0].data_ptr() for d in ds])
        geom = lambda ds: (LL * (7 * n))(*[int(x) for d in ds for x in d[1:]])
        _lib.call("b200_adamw_pack_multi", n, _ptr_array(P), _ptr_array(G), _ptr_array(M), _ptr_array(V), dims, *hyper, t,
                  _p(sq), float(clip_max_norm or 0.0), ptrs(d0), geom(d0), ptrs(d1), geom(d1), _st())
        bump_version(*P, *M, *V)
        for (p, _, _, _), geoms in items:
            for g in geoms[2:]:
                _lib.call("b200_pack_weight", _p(p.detach()), *g[:3], _p(g[3]), *g[4:], _st())
