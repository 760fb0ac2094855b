"""TEST INFRASTRUCTURE ONLY -- numpy fp64 restatement of the optimizer part of the reference's training step:
`torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)` (main.py:106) followed by `torch.optim.AdamW(...).step()`
(main.py:275).  The arithmetic lives in PyTorch (requirements.txt: torch==2.4.1; same formulas in 2.11):
clip coefficient max_norm / (||g||_2 + 1e-6) clamped to 1; decoupled weight decay p *= 1 - lr*wd; first / second
moment EMAs; bias corrections 1 - beta^t; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  Pinned by tests/test_oracle.py
against torch's own CPU implementation.  Only tests/ may import this module.
"""
import numpy as np


def clip_grad_norm(grads, max_norm):
    """Returns (total_norm, clipped copies)."""
    total = float(np.sqrt(sum(float((np.asarray(g, dtype=np.float64) ** 2).sum()) for g in grads)))
    coef = min(1.0, max_norm / (total + 1e-6))
    return total, [np.asarray(g, dtype=np.float64) * coef for g in grads]


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
    """One update (step = 1-based count AFTER this update).  All lists of float64 arrays; returns new lists."""
    b1, b2 = betas
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    out_p, out_m, out_v = [], [], []
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        p = np.asarray(p, dtype=np.float64) * (1.0 - lr * weight_decay)
        m = m + (g - m) * (1.0 - b1)
        v = b2 * v + (1.0 - b2) * g * g
        p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
        out_p.append(p), out_m.append(m), out_v.append(v)
    return out_p, out_m, out_v
