"""TEST INFRASTRUCTURE ONLY -- numpy fp64 restatement of the reference's training loss and its gradient.

Follows `compute_loss` of the reference (main.py:28-72): weighted L1 (weight 1 + 4|y|^3, main.py:38-46) plus
0.005 x the spatial-gradient loss on the [H-1, W-1] crop (main.py:49-68), masked means with a 1e-8 epsilon when
a mask is used, plain means otherwise.  The backward is hand-derived (d|x|/dx = sign x with sign 0 = 0, as
torch's abs).  Pinned by tests/golden/loss_main_compute_loss.npz, which was produced by the reference function
itself (tests/golden/make_golden_loss.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product path never does.
"""
import numpy as np


def compute_loss(y_pred, y, mask=None, use_mask=True):
    """y_pred, y, mask: [..., H, W].  Returns (loss, dloss/dy_pred) in float64."""
    yp = np.asarray(y_pred, dtype=np.float64)
    yt = np.asarray(y, dtype=np.float64)
    masked = use_mask and mask is not None
    m = np.asarray(mask, dtype=np.float64) if masked else np.ones_like(yp)
    d = yp - yt
    w = 1.0 + 4.0 * np.abs(yt) ** 3                                   # main.py:38
    if masked:
        den1 = (m * w).sum() + 1e-8                                   # main.py:42
        l1 = (np.abs(d) * m * w).sum() / den1
        g = np.sign(d) * m * w / den1
    else:
        l1 = (np.abs(d) * w).mean()                                   # main.py:45
        g = np.sign(d) * w / d.size
    e = d
    dxe = e[..., :-1, 1:] - e[..., :-1, :-1]                          # main.py:49-61 on the common crop
    dye = e[..., 1:, :-1] - e[..., :-1, :-1]
    mc = m[..., :-1, :-1]
    if masked:
        den2 = mc.sum() + 1e-8                                        # main.py:65
    else:
        den2 = float(dxe.size)                                        # main.py:67
    gl = ((np.abs(dxe) + np.abs(dye)) * mc).sum() / den2
    sx, sy = np.sign(dxe) * mc / den2, np.sign(dye) * mc / den2
    gg = np.zeros_like(yp)
    gg[..., :-1, :-1] -= sx + sy
    gg[..., :-1, 1:] += sx
    gg[..., 1:, :-1] += sy
    return l1 + 0.005 * gl, g + 0.005 * gg                             # main.py:71
